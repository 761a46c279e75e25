#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider > gpurun_out/d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/d_tests.log
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 2 > gpurun_out/d_profile.log 2>&1
timeout 300 python tools/profile_forward.py SepReformer_Base_WSJ0 32 2 1 > gpurun_out/d_profile_cl1.log 2>&1
timeout 300 python tools/gcfn_timeline.py SepReformer_Base_WSJ0 32 4000 2 2 > gpurun_out/d_timeline.log 2>&1
tail -3 gpurun_out/d_tests.log; head -6 gpurun_out/d_profile.log; head -4 gpurun_out/d_profile_cl1.log; sed -n 1,30p gpurun_out/d_timeline.log
