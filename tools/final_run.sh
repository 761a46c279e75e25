set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 263 -c 526 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_gcfn -s 56 -c 56 --csv --log-file gpurun_out/gcfn_dram_r1.csv python tools/one_forward.py > gpurun_out/ncu_dram.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gcfn -s 57 -c 2 -o gpurun_out/gcfn_r1_final -f python tools/one_forward.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
head -c 400 gpurun_out/bench_final.json
