#!/bin/bash
# round 2, call D: ncu --set full of the update's kernels (one eager update), for stall analysis
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on --cache-control none -k regex:"nature_|tcgen05|head_|dqn_loss|gather" -s 44 -c 22 -o $OUT/r2d_prof -f python scripts/profile_step.py --updates 1 > $OUT/r2d_ncu.log 2>&1; echo "ncu exit $?"; tail -3 $OUT/r2d_ncu.log
ls -la $OUT/r2d_prof.ncu-rep
