#!/bin/bash
# round 2, call T (2 GPUs): NCCL bit-identical test + the full bench contract line at N=2 (clean multi-rank exit)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -3
NG=$(nvidia-smi -L | wc -l)
T0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=180 > $OUT/r2t_pytest_multi.log 2>&1; echo "pytest multi exit $? ($(( $(date +%s) - T0 )) s)"; grep -E "^(FAILED|ERROR)|passed|failed|skipped|^E  " $OUT/r2t_pytest_multi.log | head -20
T0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 300 --warmup 20 --no-extras 2> $OUT/r2t_bench.err | tee $OUT/r2t_bench_n2.json | cut -c1-600
echo "bench N=$NG exit ${PIPESTATUS[0]} ($(( $(date +%s) - T0 )) s)"
T0=$(date +%s)
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --quick --steps 300 --warmup 20 --workload c51 2>> $OUT/r2t_bench.err | tee -a $OUT/r2t_quick.jsonl
echo "quick c51 N=$NG exit ${PIPESTATUS[0]} ($(( $(date +%s) - T0 )) s)"
tail -4 $OUT/r2t_bench.err
