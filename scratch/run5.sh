python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SYLT_COLOR_SPREAD=1 python scratch/prof_world.py 100000 600 100
python scratch/prof_world.py 100000 600 100
W="python scratch/prof_world.py 100000 600 3"
ncu --metrics gpu__time_duration.sum --clock-control none -s 10830 -c 54 --csv --log-file gpurun_out/launches_world3.csv $W > gpurun_out/ncu_world.log 2>&1
