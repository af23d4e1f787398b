set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --worlds 4096 --steps 10 --warmup 3 --skip-world100k 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 2>&1 | tail -3
