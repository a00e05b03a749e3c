#!/bin/bash
# round 2, call B: fused tail (csrc/tail.cu), fused DQN head, fc4 split-K fix-up: tests, A/B bench, launch list
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tail.py -m gpu -q --timeout=300 -x > $OUT/r2b_pytest_tail.log 2>&1; echo "pytest tail exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2b_pytest_tail.log | head -30
timeout 900 python -m pytest tests/test_gpu_learner.py -m gpu -q --timeout=300 > $OUT/r2b_pytest_learner.log 2>&1; echo "pytest learner exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2b_pytest_learner.log | head -30
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 > $OUT/r2b_pytest_parity.log 2>&1; echo "pytest parity exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2b_pytest_parity.log | head -30
run() { echo "== $1"; env $1 timeout 300 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2b_bench.err | tee -a $OUT/r2b_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_TAIL=0"
run "B2RL_FUSED_HEAD=0"
run "B2RL_FC4_FIXUP=0"
run "B2RL_TAIL=0 B2RL_FC4_FIXUP=0"
run "B2RL_X=1" "--replay sync"
run "B2RL_X=1" "--workload per"
run "B2RL_X=1" "--workload c51"
run "B2RL_X=1" "--workload qr"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2b_launches.csv python scripts/profile_step.py --updates 2 > $OUT/r2b_ncu_launch.log 2>&1; echo "launch list $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2b_launches.csv")) if len(r) > 5 and r[0].isdigit()]
# keep the last update: find kernels after the last 'nature_fused_opt' but one
names = [(r[4], float(r[-1])) for r in rows]
idx = [i for i, (n, _) in enumerate(names) if "fused_opt" in n or "rmsprop" in n]
if len(idx) >= 2:
    seg = names[idx[-2] + 1: idx[-1] + 1]
    tot = sum(t for _, t in seg)
    print("last update: %d launches, %.1f us serialised" % (len(seg), tot / 1e3))
    for n, t in seg:
        print("%8.1f  %s" % (t / 1e3, n[:90]))
PY
