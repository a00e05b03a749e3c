#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on --cache-control none -k regex:"conv_slab_tcgen05_kernel" -s 12 -c 1 -o $OUT/r2j_k1 -f python scripts/profile_step.py --updates 1 --replay sync > $OUT/r2j_ncu.log 2>&1; echo "ncu exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --cache-control none -k regex:"conv_wgrad_tcgen05_kernel" -s 11 -c 1 -o $OUT/r2j_k1w -f python scripts/profile_step.py --updates 1 --replay sync > $OUT/r2j_ncuw.log 2>&1; echo "ncu exit $?"
