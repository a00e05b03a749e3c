#!/bin/bash
# round 2, call N: wgrad M-stacking, split prefetch (select early / gather late)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=120 -k "gemm or conv or fused or agent" > $OUT/r2n_pytest_b.log 2>&1; echo "pytest parity subset exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2n_pytest_b.log | head -30
timeout 400 python -m pytest tests/test_gpu_k1.py tests/test_gpu_learner.py tests/test_gpu_step_vs_oracle.py -m gpu -q --timeout=120 > $OUT/r2n_pytest.log 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2n_pytest.log | head -30
run() { echo "== $1 $2"; env $1 timeout 120 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2n_bench.err | tee -a $OUT/r2n_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_WGRAD_STACK=0"
run "B2RL_X=1" "--replay sync"
run "B2RL_X=1" "--workload per"
echo "=== trace async"; timeout 120 python scripts/trace_step.py 2>&1 | grep -v Warning | tail -28
