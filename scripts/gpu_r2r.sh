#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_actor.py -m gpu -q --timeout=100 > $OUT/r2r_pytest_actor.log 2>&1; echo "pytest actor exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2r_pytest_actor.log | head -30
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=100 -k "ppo or a2c or gae" > $OUT/r2r_pytest_ppo.log 2>&1; echo "pytest ppo exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2r_pytest_ppo.log | head -10
timeout 200 python bench.py --workload ppo --steps 2 2> $OUT/r2r_ppo.err | tee $OUT/r2r_ppo.json | cut -c1-900
