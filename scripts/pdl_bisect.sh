#!/bin/bash
# Which source-file group's kernels tolerate programmatic dependent launch?  (GRAPES_PDL bit mask, see include/grapes_b200.h)
for m in 0 1 2 4 8 16 31; do
  GRAPES_PDL=$m timeout 300 python bench.py --workload ${1:-arxiv} --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/pdl_m$m.json 2> gpurun_out/pdl_m$m.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/pdl_m$m.json")); print("mask $m rc $rc ms/step %.4f" % d["ms_per_step"])
except Exception as e:
    print("mask $m rc $rc FAILED")
PY
done
