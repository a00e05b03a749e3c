#!/bin/bash
# quick GPU check: selected tests + bench (1 GPU), optional N-GPU bench when launched with --gpus N
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-q}
SEL=${2:-"gemm or fused or optimizer or dqn_loss"}
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -k "$SEL" > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_$TAG.log | head -60
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 200 --warmup 10 > $OUT/bench_${TAG}_n$NG.json 2> $OUT/bench_${TAG}_n$NG.err
  echo "bench N=$NG exit $?"; tail -c 1500 $OUT/bench_${TAG}_n$NG.json; tail -5 $OUT/bench_${TAG}_n$NG.err
else
  timeout 900 python bench.py --steps 300 --warmup 20 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; tail -c 2200 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
fi
