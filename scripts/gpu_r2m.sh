#!/bin/bash
# round 2, call M: late prefetch branch, batched tail loads; learner tests; traces
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tail.py tests/test_gpu_learner.py tests/test_gpu_step_vs_oracle.py -m gpu -q --timeout=120 -x > $OUT/r2m_pytest.log 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2m_pytest.log | head -30
run() { echo "== $1 $2"; env $1 timeout 120 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2m_bench.err | tee -a $OUT/r2m_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_PREFETCH_LATE=0"
run "B2RL_X=1" "--replay sync"
run "B2RL_K1=0" "--replay sync"
echo "=== trace async (late prefetch)"; timeout 120 python scripts/trace_step.py 2>&1 | grep -v Warning | tail -28
