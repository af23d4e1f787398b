set -x
B="python bench.py --steps 4 --warmup 3 --skip-cpu --skip-world100k"
$B > gpurun_out/plain_batch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_batch32k.csv $B > gpurun_out/ncu_batch.log 2>&1
B2="python bench.py --worlds 4096 --steps 4 --warmup 3 --skip-cpu --skip-world100k"
$B2 > gpurun_out/plain_batch4k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_batch_step -s 4 -c 1 -o gpurun_out/r02_prof_batch $B2 > gpurun_out/ncu_batch_full.log 2>&1
W="python tools/profiling/prof_world.py 100000 600 3"
timeout 120 $W > gpurun_out/plain_world.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 30 --csv --log-file gpurun_out/r02_launches_world100k.csv $W > gpurun_out/ncu_world.log 2>&1
timeout 120 $W > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve -s 601 -c 1 -o gpurun_out/r02_prof_solve $W > gpurun_out/ncu_solve_full.log 2>&1
python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json
tail -c 600 gpurun_out/bench_full.json
