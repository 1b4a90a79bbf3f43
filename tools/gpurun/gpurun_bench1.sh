set -x
python __graft_entry__.py smoke
python bench.py --steps 5 --warmup 4 2>&1 | tail -5
