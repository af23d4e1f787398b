python -m pytest tests -m gpu -x -q 2>&1 | tail -5
SYLT_SOLVER_SMEM_SLOTS=0 python -m pytest tests -m gpu -x -q -k "world_step or random_pile or flag" 2>&1 | tail -3
SYLT_SOLVER_SMEM_SLOTS=40 SYLT_SOLVER_SMEM_CONTACTS=50 python -m pytest tests -m gpu -x -q -k "world_step or random_pile" 2>&1 | tail -3
for sp in 1 2 4; do SYLT_COLOR_SPREAD=$sp python scratch/prof_world.py 100000 600 50; done
