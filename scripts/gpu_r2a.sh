#!/bin/bash
# round 2, call A: validate the merged fused-backward-epilogue branch (tests + A/B bench)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -2
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > $OUT/r2a_pytest_full.log 2>&1; echo "pytest full exit $?"; tail -5 $OUT/r2a_pytest_full.log
B2RL_FUSED_BWD=1 timeout 900 python -m pytest tests -m gpu -q --timeout=600 -k "learner or fused or agent or conv_grid or head" > $OUT/r2a_pytest_fused.log 2>&1; echo "pytest fused exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2a_pytest_fused.log | head -40
for f in 0 1; do
  B2RL_FUSED_BWD=$f timeout 300 python bench.py --quick --steps 300 --warmup 20 2> $OUT/r2a_bench_f$f.err | tee $OUT/r2a_bench_f$f.json
done
B2RL_FUSED_BWD=1 B2RL_GATHER_BULK=1 timeout 300 python -m pytest tests -m gpu -q --timeout=300 -k "gather or uniform_replay or fused_layers" 2>&1 | tail -5
B2RL_FUSED_BWD=1 B2RL_GATHER_BULK=1 timeout 300 python bench.py --quick --steps 300 --warmup 20 2> $OUT/r2a_bench_bulk.err | tee $OUT/r2a_bench_bulk.json
B2RL_FUSED_BWD=1 timeout 300 python bench.py --quick --replay sync --steps 300 --warmup 20 2> $OUT/r2a_bench_sync.err | tee $OUT/r2a_bench_sync.json
