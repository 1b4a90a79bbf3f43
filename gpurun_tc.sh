SB200_TRACE=1 python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>&1 | grep -E "sb200|bench" | head -14 | cut -c1-160
