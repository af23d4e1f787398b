W="python scratch/prof_world.py 100000 600 3"
timeout 120 $W > gpurun_out/plain_world.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 10230 -c 60 --csv --log-file gpurun_out/launches_world5.csv $W > gpurun_out/ncu_world.log 2>&1
