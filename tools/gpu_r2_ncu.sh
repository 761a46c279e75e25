#!/bin/bash
# round-2 ncu evidence: launch list of the bench command, DRAM traffic of the GCFN launches and of the whole forward,
# full captures (with source) of the GCFN kernel and of the attention kernel
mkdir -p gpurun_out
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sepref -s 525 -c 526 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_bench_r2.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sepref -s 259 -c 259 --csv --log-file gpurun_out/forward_dram_r2.csv python tools/one_forward.py > gpurun_out/ncu_dram_r2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_gcfn -s 57 -c 2 -o gpurun_out/gcfn_r2_final -f python tools/one_forward.py > gpurun_out/ncu_full_r2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_attn_relpos -s 23 -c 1 -o gpurun_out/attn_r2 -f python tools/one_forward.py > gpurun_out/ncu_attn_r2.log 2>&1
tail -2 gpurun_out/ncu_full_r2.log; tail -2 gpurun_out/ncu_attn_r2.log; wc -l gpurun_out/launches_r2.csv gpurun_out/forward_dram_r2.csv; ls -la gpurun_out/*.ncu-rep | tail -3
