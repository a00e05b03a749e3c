#!/bin/bash
# round 2, call I: K1 with 3-D tensor loads (tight timeouts: a hang must not burn the budget)
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python -m pytest tests/test_gpu_k1.py -m gpu -q --timeout=60 -x > $OUT/r2i_pytest_k1.log 2>&1; rc=$?; echo "pytest k1 exit $rc"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2i_pytest_k1.log | head -30
if [ $rc -ne 0 ]; then exit 1; fi
run() { echo "== $1 $2"; env $1 timeout 120 python bench.py --quick --steps 300 --warmup 20 $2 2>> $OUT/r2i_bench.err | tee -a $OUT/r2i_bench.jsonl; }
run "B2RL_X=1" "--replay sync"
run "B2RL_K1=0" "--replay sync"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file $OUT/r2i_launches.csv -k regex:"conv_slab_tcgen05_kernel|conv_wgrad_tcgen05_kernel" -s 24 -c 12 python scripts/profile_step.py --updates 1 --replay sync > $OUT/r2i_ncu.log 2>&1
python - <<'PY'
import csv
for r in csv.reader(open("gpurun_out/r2i_launches.csv")):
    if len(r) > 5 and r[0].isdigit():
        print("%8.1f  %s" % (float(r[-1]) / 1e3, r[4][:60]))
PY
