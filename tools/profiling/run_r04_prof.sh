# Round-4 profiles (run on a GPU box from the repo root: bash tools/profiling/run_r04_prof.sh; outputs under gpurun_out/).
# Every ncu command runs only after the same command has exited 0 without ncu.
set -x
W="python tools/profiling/prof_world.py 100000 1100 3"   # 1100 steps: the state bench.py times (its 5 x 100 steps end there)
# launch list of the single 100k-body world: 3 steps after 600 settle steps (8 launches per step, 9 on a re-cut step)
timeout 120 $W > gpurun_out/r04_plain_world.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 8815 -c 48 --csv --log-file gpurun_out/r04_launches_world100k.csv $W > gpurun_out/r04_ncu_world.log 2>&1
# full set of one k_solve_regions launch
timeout 120 $W > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve_regions -s 1101 -c 1 -o gpurun_out/r04_prof_solve_regions $W > gpurun_out/r04_ncu_solve_full.log 2>&1
# the batch kernel: launch list of the bench command, full set on 4096 worlds x 4 steps, DRAM bytes of 1- and 8-step launches
B="python bench.py --steps 4 --warmup 3 --skip-cpu --skip-world100k"
$B > gpurun_out/r04_plain_batch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r04_launches_batch32k.csv $B > gpurun_out/r04_ncu_batch.log 2>&1
B2="python bench.py --worlds 4096 --steps 4 --warmup 3 --skip-cpu --skip-world100k"
$B2 > gpurun_out/r04_plain_batch4k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_batch_step -s 4 -c 1 -o gpurun_out/r04_prof_batch $B2 > gpurun_out/r04_ncu_batch_full.log 2>&1
# section times printed by the solver kernel itself (not under ncu), kernel times in the pipeline, sizes
SYLT_SOLVER_TIMING=1 python tools/profiling/prof_world.py 100000 1100 1 3 > gpurun_out/r04_sections.log 2>&1
SYLT_KERNEL_TIMING=1 python tools/profiling/prof_world.py 100000 1100 1 3 > gpurun_out/r04_kernel_times.log 2>&1
SYLT_REGIONS_PERIOD=2 SYLT_KERNEL_TIMING=1 SYLT_SOLVER_TIMING=1 python tools/profiling/prof_world.py 100000 600 1 2 > gpurun_out/r04_recut_step.log 2>&1
for n in 1000 4000 10000 20000 50000 100000; do for r in 0 1; do echo "bodies $n SYLT_REGIONS=$r"; SYLT_REGIONS=$r python tools/profiling/prof_world.py $n 600 100 3 2>&1 | grep ms/step | cut -c1-28; done; done > gpurun_out/r04_sizes.log 2>&1
python tools/profiling/soak_world.py > gpurun_out/r04_soak.log 2>&1
python bench.py > gpurun_out/r04_bench_n1.json 2> gpurun_out/r04_bench_n1.err
tail -c 300 gpurun_out/r04_bench_n1.json
