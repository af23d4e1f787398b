for ns in 0 20 100; do echo sleep $ns; SYLT_SOLVER_SLEEP_NS=$ns timeout 120 python scratch/prof_world.py 100000 600 100; done
W="python scratch/prof_world.py 100000 600 3"
timeout 120 $W > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve -s 601 -c 1 -o gpurun_out/prof_solve4 $W > gpurun_out/ncu_solve_full.log 2>&1
