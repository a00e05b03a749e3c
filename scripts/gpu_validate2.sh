#!/bin/bash
# round-2 final validation on one B200 (time-boxed, most important first): full GPU test suite, smoke(), the driver's bench
# commands, the launch list of one update, stall profile of the persistent PPO kernel, racecheck of the replay / sum-tree tests.
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r02b}
T0=$(date +%s); lap() { echo "[$(( $(date +%s) - T0 )) s] $1"; }
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > $OUT/pytest_full_$TAG.log 2>&1; lap "pytest exit $?"; tail -3 $OUT/pytest_full_$TAG.log
timeout 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; lap smoke
timeout 500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; lap "bench $?"
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; lap "bench reference $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_async_$TAG.csv python scripts/profile_step.py --updates 2 > $OUT/ncu_launch_async_$TAG.log 2>&1; lap "launch list async $?"
timeout 150 ncu --section WarpStateStats --section SchedulerStats --section InstructionStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --clock-control none -k regex:ppo_minibatch -s 1 -c 1 --csv --page raw --log-file $OUT/ncu_ppo_$TAG.csv python scripts/ppo_phase_clocks.py 256 > $OUT/ncu_ppo_$TAG.log 2>&1; lap "ncu ppo $?"
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=280 -k "sumtree_trace or sumtree_batched or uniform_replay_matches or prioritized_replay_matches" > $OUT/racecheck_$TAG.log 2>&1; lap "racecheck exit $?"; tail -4 $OUT/racecheck_$TAG.log
