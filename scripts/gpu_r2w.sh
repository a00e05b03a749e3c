#!/bin/bash
# round 2, call W: phase clocks of the persistent PPO kernel, arena rollout + feed_many parity, agent-API throughput
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python scripts/ppo_phase_clocks.py 1024 2>&1 | grep -v Warning | tee $OUT/r2w_ppo_clocks.txt
timeout 250 python -m pytest tests/test_gpu_actor.py tests/test_gpu_q_actor.py tests/test_ppo_persistent.py -m gpu -q --timeout=100 > $OUT/r2w_pytest.log 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2w_pytest.log | head -30
timeout 200 python bench.py --workload ppo --steps 2 2> $OUT/r2w_ppo.err | tee $OUT/r2w_ppo.json | cut -c1-200; echo "ppo exit ${PIPESTATUS[0]}"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2w_ppo.json")); print(d["ms_per_step"], d["minibatch_phase"]["seconds_per_iteration"], d["env_steps_per_s"])
PY
timeout 150 python - <<'PY' 2>&1 | grep -v "Warning\|INFO" | tail -3
import json, torch, bench
import deeprl_b200 as rl
rl.select_device(0); rl.Config.COMPUTE_DTYPE = torch.bfloat16
print(json.dumps(bench.agent_e2e(rl, steps=200)))
PY
