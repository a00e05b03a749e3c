#!/bin/bash
# round-end validation on one B200: full GPU test suite, smoke(), bench for every workload, launch list + ncu captures
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-v}
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest_full_$TAG.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_full_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 300 --warmup 20 > $OUT/bench_dqn_$TAG.json 2> $OUT/bench_dqn_$TAG.err; echo "bench dqn $?"
for w in per c51 qr; do
  timeout 900 python bench.py --workload $w --steps 100 --warmup 10 > $OUT/bench_${w}_$TAG.json 2> $OUT/bench_${w}_$TAG.err; echo "bench $w $?"
done
timeout 300 python bench.py --workload ppo --steps 2 > $OUT/bench_ppo_$TAG.json 2> $OUT/bench_ppo_$TAG.err; echo "bench ppo $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench reference $?"
python - <<PY
import json
for w in ("dqn", "per", "c51", "qr", "ppo", "ref"):
    try:
        d = json.load(open("$OUT/bench_%s_$TAG.json" % w))
        print(w, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("other_replay_mode"), d.get("gpu_launches_per_step"))
    except Exception as e:
        print(w, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv python scripts/profile_step.py --updates 2 > $OUT/ncu_launch_$TAG.log 2>&1; echo "launch list $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather -c 3 -o $OUT/prof_gather_$TAG -f python scripts/profile_step.py --updates 1 > $OUT/ncu_gather_$TAG.log 2>&1; echo "ncu gather $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcgen05 -s 26 -c 13 -o $OUT/prof_tc_$TAG -f python scripts/profile_step.py --updates 1 > $OUT/ncu_tc_$TAG.log 2>&1; echo "ncu tcgen05 $?"
