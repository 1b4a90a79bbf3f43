timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2000
