#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tail.py tests/test_gpu_parity.py -m gpu -q --timeout=100 -k "dqn_head or qr_loss or loss_boundary or agent_update" > $OUT/r2q_pytest_head.log 2>&1; echo "pytest head+qr exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2q_pytest_head.log | head -20
run() { echo "== $1 $2"; env $1 timeout 120 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2q_bench.err | tee -a $OUT/r2q_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_FUSED_HEAD=0"
run "B2RL_X=1" "--workload per"
run "B2RL_FUSED_HEAD=0" "--workload per"
run "B2RL_X=1" "--workload qr"
