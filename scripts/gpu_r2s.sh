#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_learner.py tests/test_gpu_actor.py -m gpu -q --timeout=120 > $OUT/r2s_pytest.log 2>&1; echo "pytest learner+actor exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2s_pytest.log | head -20
run() { echo "== $1 $2"; env $1 timeout 120 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2s_bench.err | tee -a $OUT/r2s_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_PREFETCH_AT=end"
run "B2RL_X=1"
echo "=== trace async"; timeout 120 python scripts/trace_step.py 2>&1 | grep -v Warning | tail -26
