timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
python bench.py --steps 6 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'])"
