#!/bin/bash
# ncu launch list of eager updates (+ optional full capture of one kernel family)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-p}; KREGEX=${2:-gemm_tcgen05}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python scripts/profile_step.py --updates 2 > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 14 -c 14 -o $OUT/prof_${KREGEX}_$TAG -f \
    python scripts/profile_step.py --updates 1 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ls -la $OUT | tail -5
