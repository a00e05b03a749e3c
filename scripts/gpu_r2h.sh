#!/bin/bash
# round 2, call H: ncu of the K1 conv1 kernels (sync replay) next to the TMA versions
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on --cache-control none -k regex:"conv_slab_tcgen05_kernel|conv_wgrad_tcgen05_kernel" -s 24 -c 12 -o $OUT/r2h_k1 -f python scripts/profile_step.py --updates 1 --replay sync > $OUT/r2h_ncu_k1.log 2>&1; echo "ncu k1 exit $?"
B2RL_K1=0 timeout 900 ncu --set full --clock-control none --import-source on --cache-control none -k regex:"conv_slab_tcgen05_kernel|conv_wgrad_tcgen05_kernel" -s 24 -c 12 -o $OUT/r2h_tma -f python scripts/profile_step.py --updates 1 --replay sync > $OUT/r2h_ncu_tma.log 2>&1; echo "ncu tma exit $?"
ls -la $OUT/r2h_*.ncu-rep
