#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_sizes.py tests/test_gpu_parity.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/k_bench.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "parity", d["parity"]["rel_l2_vs_fp32_path"], "launches", d["gpu_launches"])
print({k: round(v, 3) for k, v in d["kernel_ms"].items() if k.endswith("_ms")})
PY
