#!/bin/bash
# usage (under gpurun): bash dev/gpu_check.sh [pytest -k expr] -- runs the GPU parity tests and short benches
K="${1:-}"
if [ -n "$K" ]; then timeout 300 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -8; else timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; fi
for w in ${WORKLOADS:-cfg4 cfg5}; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo rc=$?
  tail -c 700 gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    e=j.get("e2e") or {}
    print("$w ms", round(j["ms_per_step"],3), "%.3e"%j["value"], "e2e", round(e.get("ms_per_step",0),3), j["config"]["mode"])
    print("   dev:", {k:round(v,3) for k,v in j["kernel_ms_per_step"].items()})
    print("   e2e:", {k:round(v,3) for k,v in (e.get("kernel_ms_per_step") or {}).items()})
except Exception as ex: print("ERR", ex)
PY
done
