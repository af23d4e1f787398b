B2="python bench.py --worlds 4096 --steps 2 --warmup 1 --skip-cpu --skip-world100k"
$B2 > gpurun_out/plain_batch4k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_batch_step -s 2 -c 1 -o gpurun_out/prof_batch2 $B2 > gpurun_out/ncu_batch_full.log 2>&1
