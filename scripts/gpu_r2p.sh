#!/bin/bash
# round 2, call P: distributional heads on tcgen05, async / checkpoint tests
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_tail.py -m gpu -q --timeout=100 -k "dist_head" > $OUT/r2p_pytest_dh.log 2>&1; echo "pytest dist head exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2p_pytest_dh.log | head -20
timeout 300 python -m pytest tests/test_gpu_async_and_checkpoint.py -m gpu -q --timeout=120 > $OUT/r2p_pytest_ac.log 2>&1; echo "pytest async/ckpt exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2p_pytest_ac.log | head -20
timeout 400 python -m pytest tests/test_gpu_learner.py tests/test_gpu_step_vs_oracle.py -m gpu -q --timeout=120 > $OUT/r2p_pytest.log 2>&1; echo "pytest learner+oracle exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2p_pytest.log | head -30
run() { echo "== $1 $2"; env $1 timeout 120 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2p_bench.err | tee -a $OUT/r2p_bench.jsonl; }
run "B2RL_X=1" "--workload c51"
run "B2RL_DIST_HEAD=0" "--workload c51"
run "B2RL_X=1" "--workload qr"
run "B2RL_DIST_HEAD=0" "--workload qr"
