#!/bin/bash
# 2 GPUs: data_parallel drop-in test, N=2 bench; then the other workloads on one GPU
mkdir -p gpurun_out
python -m pytest tests/test_dropin_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/f_dp.log 2>&1; echo "dp rc=$?" >> gpurun_out/f_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err; echo "n2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err &
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c5.json 2> gpurun_out/f_bench_c5.err &
wait
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_c1.json 2> gpurun_out/f_bench_c1.err
tail -4 gpurun_out/f_dp.log
python - <<'PY'
import json
for n in ("n2", "c4", "c5", "c1"):
    try:
        d = json.loads(open(f"gpurun_out/f_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "frac", d["roofline"] and round(d["roofline"]["frac"], 3), "parity", d["parity"] and d["parity"]["rel_l2_vs_fp32_path"], "refgpu", d.get("reference_gpu") and {k: round(v["ms_per_step"], 1) for k, v in d["reference_gpu"].items() if isinstance(v, dict)}, "enq", round(d["host_enqueue_ms_b1"], 3))
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/f_bench_{n}.err").read()[-800:])
PY
