#!/bin/bash
# round 2, call V: persistent PPO minibatch kernel -- parity on the device, PPO iteration timing
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_ppo_persistent.py "tests/test_gpu_learner.py::test_graphed_ppo_minibatches_match_the_eager_loop" tests/test_gpu_actor.py -m gpu -q --timeout=100 > $OUT/r2v_pytest.log 2>&1; echo "pytest exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2v_pytest.log | head -30
timeout 200 python bench.py --workload ppo --steps 2 2> $OUT/r2v_ppo.err | tee $OUT/r2v_ppo.json | cut -c1-1500; echo "ppo exit ${PIPESTATUS[0]}"; tail -3 $OUT/r2v_ppo.err
