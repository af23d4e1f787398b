# Round-3 profiles of the single-world path (the batch kernel is unchanged: r02_* stay current for it).
# Run on a GPU box from the repo root: bash tools/profiling/run_r03_prof.sh   (outputs under gpurun_out/)
set -x
W="python tools/profiling/prof_world.py 100000 600 3"
# launch list: 3 steps after 600 settle steps (12 launches per step incl. the periodic k_resort / k_rebalance)
timeout 120 $W > gpurun_out/r03_plain_world.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 6010 -c 36 --csv --log-file gpurun_out/r03_launches_world100k.csv $W > gpurun_out/r03_ncu_world.log 2>&1
# full set of one k_solve_regions launch (cooperative; ncu replays it with all CTAs resident)
timeout 120 $W > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_solve_regions -s 601 -c 1 -o gpurun_out/r03_prof_solve_regions $W > gpurun_out/r03_ncu_solve_full.log 2>&1
# section times printed by the kernel itself (not under ncu)
SYLT_SOLVER_TIMING=1 python tools/profiling/prof_world.py 100000 600 100 3 > gpurun_out/r03_sections.log 2>&1
for n in 1000 4000 10000 20000 50000 100000; do for r in 0 1; do echo "bodies $n SYLT_REGIONS=$r"; SYLT_REGIONS=$r python tools/profiling/prof_world.py $n 600 100 3 2>&1 | grep ms/step | cut -c1-28; done; done > gpurun_out/r03_sizes.log 2>&1
python bench.py > gpurun_out/r03_bench.json 2> gpurun_out/r03_bench.err
tail -c 400 gpurun_out/r03_bench.json
