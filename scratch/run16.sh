for sl in 0 25 50 100 200; do echo slack $sl; SYLT_SOLVER_SLACK=$sl SYLT_SOLVER_TIMING=1 timeout 120 python scratch/prof_world.py 100000 600 100 2>&1 | tail -2 | cut -c1-160; done
