#!/bin/bash
# full ncu capture of the kernels matching a regex inside one eager update (after warm-up updates)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-p}; KREGEX=${2:-conv_wgrad}; SKIP=${3:-8}; CNT=${4:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $CNT -o $OUT/prof_${KREGEX}_$TAG -f \
    python scripts/profile_step.py --updates 1 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ls -la $OUT/*.ncu-rep | tail -3
