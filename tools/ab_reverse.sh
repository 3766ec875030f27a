#!/bin/bash
# Developer tool (GPU box): A/B of the last-to-first tile order (slab and conv+pool kernels) inside the whole
# step, alternating in one box.  Y2_SLAB_REVERSE=0 disables it, unset = the planner's choice.
for r in 0 default 0 default; do
  if [ $r = default ]; then unset Y2_SLAB_REVERSE; else export Y2_SLAB_REVERSE=$r; fi
  python bench.py --no-cpu-baseline --profile-out gpurun_out/prof_rev$r.json > gpurun_out/bench_rev$r.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_rev$r.json")); p=json.load(open("gpurun_out/prof_rev$r.json"))
print("$r", d["value"], d["ms_per_step"], [(x["layer"], x["ms"]) for x in p["layers"] if x["layer"] in (2,4,5,6,8,9)])
PY
done
