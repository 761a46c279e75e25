#!/bin/bash
# final check after a k_tok change: full GPU tests, smoke, default bench, c1, per-kernel profile, cla_b timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
timeout 400 python bench.py --workload c1 --steps 30 --warmup 5 > gpurun_out/final_bench_c1.json 2> gpurun_out/final_bench_c1.err
timeout 200 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/final_profile.log 2>&1
timeout 200 python tools/tok_timeline.py cla_b 32 8000 > gpurun_out/tl_cla_b.txt 2>&1
python - <<'PY'
import json
for n in ("c2", "c1"):
    d = json.loads(open(f"gpurun_out/final_bench_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["parity"]["rel_l2_vs_fp32_path"], d["parity"].get("si_snri_delta_db"))
PY
head -8 gpurun_out/final_profile.log; tail -1 gpurun_out/tl_cla_b.txt | cut -c1-600
