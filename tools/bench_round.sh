#!/bin/bash
# Runs on the GPU box (under gpurun): the bench line of every workload (device-resident value, e2e, cpu_baseline) and
# the reference arm of the default workload.  Outputs -> gpurun_out/${TAG}_bench_*.json
set -u
TAG=${1:-r01d}
for w in cfg3 cfg1 cfg2 cfg4 n1; do
  timeout 300 python bench.py --workload $w > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err || echo "bench $w failed"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_cfg3.json 2> gpurun_out/${TAG}_bench_reference.err
for w in cfg3 cfg1 cfg2 cfg4 n1; do python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_$w.json"))
r = d["roofline"]
print("$w", "ms/step %.4f" % d["ms_per_step"], "value %.3e" % d["value"], "roofline %.1f %s frac %.3f" % (r["achieved"], r["unit"], r["frac"]),
      "e2e %.3e" % d["e2e"]["value"], d["clocks"])
PY
done
