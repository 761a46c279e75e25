#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_sizes.py -q -m gpu -x -p no:cacheprovider -k "not c1 and not pooled" > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.log
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/j_profile.log 2>&1
tail -4 gpurun_out/j_tests.log; head -14 gpurun_out/j_profile.log
