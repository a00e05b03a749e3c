#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
run() { echo "== $1 $2"; env $1 timeout 100 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2o_bench.err | tee -a $OUT/r2o_bench.jsonl; }
run "B2RL_X=1"
run "B2RL_WGRAD_CTAS=74"
run "B2RL_WGRAD_CTAS=100"
run "B2RL_K1=0" "--replay sync"
run "B2RL_PREFETCH_LATE=0"
run "B2RL_FC4_SPLITS=2"
run "B2RL_FC4_SPLITS=8"
