#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gcfn_pair_check.py > gpurun_out/c_pair.log 2>&1
echo "rc=$?" >> gpurun_out/c_pair.log
timeout 300 python tools/gcfn_pair_timeline.py > gpurun_out/c_timeline.log 2>&1
tail -8 gpurun_out/c_pair.log; sed -n 1,60p gpurun_out/c_timeline.log
