#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gcfn_pair_check.py trio > gpurun_out/g_trio.log 2>&1
echo "rc=$?" >> gpurun_out/g_trio.log
timeout 300 python tools/gcfn_pair_timeline.py 32 4000 trio > gpurun_out/g_timeline.log 2>&1
tail -12 gpurun_out/g_trio.log; sed -n 1,48p gpurun_out/g_timeline.log
