#!/bin/bash
# usage (under gpurun): bash dev/ab_variants.sh "<variant names>" "<workloads>"  -- short benches of dev/variants/lib_<name>.so
for v in $1; do
  for w in $2; do
    L=""; [ "$v" != "base" ] && L="$PWD/dev/variants/lib_$v.so"
    CONCEPTHASH_B200_LIB=$L timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-others > gpurun_out/ab_${v}_$w.json 2> gpurun_out/ab_${v}_$w.err
    python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/ab_${v}_$w.json").read().strip().splitlines()[-1])
    print("$v $w ms", round(j["ms_per_step"],3), {k:round(x,3) for k,x in j["kernel_ms_per_step"].items() if "select" in k}, j["parity_check"]["ok"])
except Exception as ex: print("$v $w ERR", ex)
PY
  done
done
