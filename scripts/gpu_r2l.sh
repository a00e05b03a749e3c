#!/bin/bash
# round 2, call L (2 GPUs): NCCL bit-identical test, in-graph overlapped all-reduce vs the split form
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=250 > $OUT/r2l_pytest_multi.log 2>&1; echo "pytest multi exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|skipped|^E  " $OUT/r2l_pytest_multi.log | head -20
NG=$(nvidia-smi -L | wc -l)
run() { echo "== $1 N=$2 $3"; env $1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --quick --steps 300 --warmup 20 $3 2>> $OUT/r2l_bench.err | tee -a $OUT/r2l_bench.jsonl; }
run "B2RL_X=1" $NG ""
run "B2RL_NCCL_IN_GRAPH=0" $NG ""
run "B2RL_X=1" 1 ""
run "B2RL_X=1" $NG "--workload per"
tail -5 $OUT/r2l_bench.err
