#!/bin/bash
# round-2 final evidence on one B200: GPU tests, smoke, bench lines (ours + reference arm), ncu launch list / traffic / captures
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -7 > gpurun_out/final_smoke.log
timeout 900 python bench.py 2> gpurun_out/final_bench_c2.err | tail -1 > gpurun_out/final_bench_c2.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_bench_ref.json
timeout 600 python bench.py --workload c1 --steps 30 --warmup 5 2>/dev/null | tail -1 > gpurun_out/final_bench_c1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 500 -c 500 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu --no-cuda-graph > gpurun_out/ncu_bench_r2.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_ -s 246 -c 246 --csv --log-file gpurun_out/forward_dram_r2.csv python tools/one_forward.py > gpurun_out/ncu_dram_r2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_gcfn -s 57 -c 2 -o gpurun_out/gcfn_r2_final -f python tools/one_forward.py > gpurun_out/ncu_full_r2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_attn_relpos -s 23 -c 1 -o gpurun_out/attn_r2 -f python tools/one_forward.py > gpurun_out/ncu_attn_r2.log 2>&1
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/final_profile.log 2>&1
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log; head -c 600 gpurun_out/final_bench_c2.json; echo; wc -l gpurun_out/launches_r2.csv gpurun_out/forward_dram_r2.csv
