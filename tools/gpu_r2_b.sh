#!/bin/bash
# round-2 GPU call B: packed-FMA GCFN epilogue - parity (existing suite) + timing (default and wide tiles) + timeline
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_sizes.py tests/test_dropin_gpu.py -q -m gpu -x -p no:cacheprovider -k "not test_pooled and not c1 and not large_blocks" > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/b_tests.log
timeout 300 python tools/gcfn_wide_check.py > gpurun_out/b_wide.log 2>&1
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/b_profile.log 2>&1
timeout 300 python tools/gcfn_timeline.py SepReformer_Base_WSJ0 32 4000 2 2 > gpurun_out/b_timeline.log 2>&1
tail -3 gpurun_out/b_tests.log; cat gpurun_out/b_wide.log; head -8 gpurun_out/b_profile.log
