#!/bin/bash
# round-2 GPU call A: new parity tests, wait-loop variants, GCFN timeline, wide-tile check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
python -m pytest tests/test_gpu_parity_sizes.py tests/test_dropin_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/a_tests.log
timeout 900 python tools/variants.py time 32 > gpurun_out/a_variants.log 2>&1
timeout 300 python tools/gcfn_timeline.py SepReformer_Base_WSJ0 32 4000 2 2 > gpurun_out/a_timeline.log 2>&1
timeout 300 python tools/gcfn_wide_check.py > gpurun_out/a_wide.log 2>&1
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/a_profile.log 2>&1
tail -5 gpurun_out/a_tests.log; cat gpurun_out/a_variants.log
