SB200_TRACE=1 python bench.py --steps 3 --warmup 4 --no-cpu-baseline 2>&1 | tail -40
