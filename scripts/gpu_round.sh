#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list + full capture of the gather kernel.
# usage: scripts/gpu_round.sh [tests|notests] [tag]
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${2:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1
if [ "$1" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  grep -E "^(FAILED|ERROR)|passed|failed" $OUT/pytest_$TAG.log | tail -30
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $OUT/smoke_$TAG.log; tail -3 $OUT/smoke_$TAG.log
timeout 900 python bench.py --steps 300 --warmup 20 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
for w in ${WORKLOADS:-per c51 qr}; do
  timeout 600 python bench.py --steps 200 --warmup 10 --workload $w > $OUT/bench_${w}_$TAG.json 2> $OUT/bench_${w}_$TAG.err; echo "bench $w exit $?"; head -c 600 $OUT/bench_${w}_$TAG.json; echo
done
# launch list (cold-cache, serialised: shares only) of 2 eager-equivalent updates
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python scripts/profile_step.py --updates 3 > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather -c 4 -o $OUT/prof_gather_$TAG -f \
    python scripts/profile_step.py --updates 2 > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ls -la $OUT | tail -20
