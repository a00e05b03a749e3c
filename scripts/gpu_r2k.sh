#!/bin/bash
# round 2, call K: new tests (step vs oracle), smoke, full bench line; tight timeouts
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_step_vs_oracle.py -m gpu -q --timeout=200 > $OUT/r2k_pytest_oracle.log 2>&1; echo "pytest step-vs-oracle exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2k_pytest_oracle.log | head -40
timeout 300 python -m pytest tests/test_gpu_tail.py tests/test_gpu_learner.py tests/test_gpu_k1.py -m gpu -q --timeout=120 > $OUT/r2k_pytest_a.log 2>&1; echo "pytest tail+learner+k1 exit $?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/r2k_pytest_a.log | head -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r2k_bench.json 2> $OUT/r2k_bench.err; echo "bench exit $?"; tail -c 600 $OUT/r2k_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_bench.json"))
for k in ("value", "ms_per_step", "repeat_ms", "gpu_launches_per_step", "tensor_frac_of_sustained"):
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], "agent", d.get("e2e_agent"))
print("other", d["other_replay_mode"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "us_per_launch", "traffic")}, d["roofline"]["raw_u8_variant"])
print("tensor", {k: d["roofline_tensor"][k] for k in ("achieved", "frac", "us_per_launch")}, d["roofline_tensor"]["whole_step"])
print("cpu", {k: d["cpu_baseline"][k] for k in ("value", "cores", "one_thread", "capped_threads")})
for w, r in d["extra_workloads"].items():
    print(w, r.get("value"), r.get("ms_per_step"), (r.get("e2e") or {}).get("value"), r.get("gpu_launches_per_step"), r.get("minibatch_phase"), r.get("error"))
PY
