python -m pytest tests -m gpu -x -q -k "batch or smoke" 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --skip-world100k --skip-cpu 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'])"
