for k in 1 2 4; do echo ctas_per_sm $k; SYLT_SOLVER_CTAS_PER_SM=$k timeout 120 python scratch/prof_world.py 100000 600 100; done
SYLT_SOLVER_CTAS_PER_SM=2 timeout 300 python -m pytest tests -m gpu -x -q -k "world_step or random_pile or config3" 2>&1 | tail -3
