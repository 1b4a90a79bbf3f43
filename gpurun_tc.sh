timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>&1 | tail -12 | cut -c1-600
