#!/bin/bash
# quick kernel-only timing of the fused index build on cfg2 and cfg3 (1 GiB each); prints kernel_ms / frac
# usage (on the GPU box): bash tools/kbench.sh [tag]
tag=${1:-k}
for wl in cfg2_unquoted cfg3_quoted; do
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/kb_${tag}_${wl}.json 2> gpurun_out/kb_${tag}_${wl}.err || tail -5 gpurun_out/kb_${tag}_${wl}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/kb_${tag}_${wl}.json"))
r=d["roofline"]
print("${tag} ${wl}: kernel_ms=%.4f frac=%.3f step_ms=%.4f value=%.0f e2e=%.1f" % (r["kernel_ms"], r["frac"], d["ms_per_step"], d["value"], d["e2e"]["value"]))
PY
done
