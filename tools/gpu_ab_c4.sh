#!/bin/bash
# A/B of two library builds on the Large workload (C4): per-kernel times of spk_proj
for lib in sepreformer_b200/variants/libsepref_old040.so sepreformer_b200/libsepref_b200.so; do
  SEPREF_LIB=$PWD/$lib timeout 300 python tools/profile_forward.py SepReformer_Large_DM_WHAMR 16 2 2 2>&1 | grep -i "sum of\|spproj\|qkv>\|gcfn" | head -5
done
