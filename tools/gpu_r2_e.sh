#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_c2.json 2> gpurun_out/e_bench_c2.err
echo "c2 rc=$?"; tail -3 gpurun_out/e_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err
echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/e_bench_c2.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "gpu_launches", "host_enqueue_ms_b1"): print(k, d[k])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "feature_boundary", d["e2e"]["feature_boundary"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "hbm_frac", "share_of_step")}, d["roofline"]["whole_step"]["tensor_frac"])
print("parity", d["parity"]); print("reference_gpu", d["reference_gpu"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"]); print("tf32", d["tf32"])
r = json.loads(open("gpurun_out/e_bench_ref.json").read().strip().splitlines()[-1]); print("ref arm", r["value"], r["cpu_baseline"])
PY
