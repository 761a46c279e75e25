#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "blocks or separator" > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/i_tests.log
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/i_profile.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 --no-reference-gpu > gpurun_out/i_bench_c2.json 2> gpurun_out/i_bench_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c1 --steps 30 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/i_bench_c1.json 2> gpurun_out/i_bench_c1.err; echo "c1 rc=$?"
tail -2 gpurun_out/i_tests.log; head -14 gpurun_out/i_profile.log
python - <<'PY'
import json
for n in ("c2", "c1"):
    try:
        d = json.loads(open(f"gpurun_out/i_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "feat e2e", round(d["e2e"]["feature_boundary"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"], "enq", round(d["host_enqueue_ms_b1"], 3), "launches", d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/i_bench_{n}.err").read()[-1500:])
PY
