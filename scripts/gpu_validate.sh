#!/bin/bash
# round-2 validation on one B200: full GPU test suite, smoke(), the driver's bench command, launch lists + ncu captures,
# compute-sanitizer racecheck of the sum-tree / gather tests.  Every step is time-boxed.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r02}
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > $OUT/pytest_full_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_full_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench $?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench reference $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_async_$TAG.csv python scripts/profile_step.py --updates 2 > $OUT/ncu_launch_async_$TAG.log 2>&1; echo "launch list async $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_sync_$TAG.csv python scripts/profile_step.py --updates 2 --replay sync > $OUT/ncu_launch_sync_$TAG.log 2>&1; echo "launch list sync (K1) $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_cvt -c 2 -o $OUT/prof_gather_$TAG -f python scripts/profile_step.py --updates 1 > $OUT/ncu_gather_$TAG.log 2>&1; echo "ncu gather $?"
timeout 400 ncu --set full --clock-control none --import-source on --cache-control none -k regex:"tcgen05|nature_" -s 44 -c 22 -o $OUT/prof_step_$TAG -f python scripts/profile_step.py --updates 1 > $OUT/ncu_step_$TAG.log 2>&1; echo "ncu step kernels $?"
timeout 120 python scripts/trace_step.py > $OUT/timeline_async_$TAG.txt 2>&1; timeout 120 python scripts/trace_step.py --replay sync > $OUT/timeline_sync_$TAG.txt 2>&1; echo "timelines done"
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=380 -k "sumtree_trace or sumtree_batched or uniform_replay_matches or prioritized_replay_matches" > $OUT/racecheck_$TAG.log 2>&1; echo "racecheck exit $?"; tail -4 $OUT/racecheck_$TAG.log
