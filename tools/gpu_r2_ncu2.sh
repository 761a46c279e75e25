#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 525 -c 526 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_bench_r2.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_ -s 259 -c 259 --csv --log-file gpurun_out/forward_dram_r2.csv python tools/one_forward.py > gpurun_out/ncu_dram_r2.log 2>&1
wc -l gpurun_out/launches_r2.csv gpurun_out/forward_dram_r2.csv
