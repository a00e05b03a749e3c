#!/bin/bash
# round 2, call G: K1 with dedicated converter warps
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_k1.py -m gpu -q --timeout=300 > $OUT/r2g_pytest_k1.log 2>&1; echo "pytest k1 exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2g_pytest_k1.log | head -30
run() { echo "== $1 $2"; env $1 timeout 300 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2g_bench.err | tee -a $OUT/r2g_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_X=1" "--replay sync"
run "B2RL_K1=0" "--replay sync"
echo "=== trace sync (K1)"; timeout 300 python scripts/trace_step.py --replay sync 2>&1 | grep -v Warning | tail -30
echo "=== trace async"; timeout 300 python scripts/trace_step.py 2>&1 | grep -v Warning | tail -30
