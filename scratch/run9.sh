python -m pytest tests -m gpu -x -q 2>&1 | tail -5
SYLT_TILING=0 python -m pytest tests -m gpu -x -q -k "world_step or random_pile or config3" 2>&1 | tail -3
SYLT_SOLVER_SMEM_SLOTS=40 SYLT_SOLVER_SMEM_CONTACTS=50 python -m pytest tests -m gpu -x -q -k "world_step or random_pile" 2>&1 | tail -3
python scratch/prof_world.py 100000 600 100
SYLT_TILING=0 python scratch/prof_world.py 100000 600 100
