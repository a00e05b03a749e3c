#!/bin/bash
# round 2, call U: graphed Q actor + n-step DQN on the device, agent-API throughput
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 280 python -m pytest tests/test_gpu_q_actor.py tests/test_nstep_dqn.py tests/test_gpu_async_and_checkpoint.py -m gpu -q --timeout=120 -x > $OUT/r2u_pytest.log 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2u_pytest.log | head -30
timeout 150 python - <<'PY' 2>&1 | grep -v Warning | tail -5
import json, torch, bench
import deeprl_b200 as rl
rl.select_device(0); rl.Config.COMPUTE_DTYPE = torch.bfloat16
print(json.dumps(bench.agent_e2e(rl, steps=200)))
PY
