set -x
B="python bench.py --steps 2 --warmup 1 --skip-cpu --skip-world100k"
$B > gpurun_out/plain_batch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_batch.csv $B > gpurun_out/ncu_batch.log 2>&1
B2="python bench.py --worlds 4096 --steps 2 --warmup 1 --skip-cpu --skip-world100k"
$B2 > gpurun_out/plain_batch4k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_batch_step -s 2 -c 2 -o gpurun_out/prof_batch $B2 > gpurun_out/ncu_batch_full.log 2>&1
W="python scratch/prof_world.py 100000 600 3"
$W > gpurun_out/plain_world.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 10830 -c 54 --csv --log-file gpurun_out/launches_world.csv $W > gpurun_out/ncu_world.log 2>&1
tail -2 gpurun_out/plain_world.log
$W > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve -s 601 -c 1 -o gpurun_out/prof_solve $W > gpurun_out/ncu_solve_full.log 2>&1
ls -la gpurun_out
