#!/bin/bash
# k_gcfn_tm: parity on the block shapes, forward time, timelines of both modes
timeout 500 python tools/gcfn_tm_check.py ${1:-1 2} full 2>&1 | grep -v "rows=1 \|rows=2 \|rows=3 " | tail -16
timeout 200 python tools/gcfn_tm_timeline.py 32 8000 1 > gpurun_out/tm_tl1.txt 2>&1; head -1 gpurun_out/tm_tl1.txt
