W="python scratch/prof_world.py 100000 600 3"
$W > gpurun_out/plain_world.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 10830 -c 60 --csv --log-file gpurun_out/launches_world4.csv $W > gpurun_out/ncu_world.log 2>&1
$W > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve -s 601 -c 1 -o gpurun_out/prof_solve3 $W > gpurun_out/ncu_solve_full.log 2>&1
