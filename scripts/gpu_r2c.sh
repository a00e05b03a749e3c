#!/bin/bash
# round 2, call C: in-graph timeline of the update in several configurations
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tail.py -m gpu -q --timeout=300 > $OUT/r2c_pytest_tail.log 2>&1; echo "pytest tail exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2c_pytest_tail.log | head -30
tr() { echo "=== $1 $2"; env $1 timeout 300 python scripts/trace_step.py $2 2>&1 | grep -v Warning | tail -40; }
tr "B2RL_X=1" "" > $OUT/r2c_trace_default.txt
tr "B2RL_FUSED_HEAD=1" "" > $OUT/r2c_trace_fusedhead.txt
tr "B2RL_TAIL=0" "" > $OUT/r2c_trace_notail.txt
tr "B2RL_X=1" "--replay sync" > $OUT/r2c_trace_sync.txt
tr "B2RL_FC4_FIXUP=0" "" > $OUT/r2c_trace_nofixup.txt
cat $OUT/r2c_trace_default.txt
