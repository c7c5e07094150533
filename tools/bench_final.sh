for w in cfg3 cfg4 cfg1; do timeout 200 python bench.py --workload $w > gpurun_out/r01g_bench_$w.json 2> gpurun_out/r01g_bench_$w.err; done
python - <<'PY'
import json
for w in ("cfg3","cfg4","cfg1"):
    d=json.load(open("gpurun_out/r01g_bench_%s.json"%w)); r=d["roofline"]
    print(w, round(d["ms_per_step"],4), "%.3e"%d["value"], round(r["achieved"],1), round(r["frac"],3), "e2e %.3e"%d["e2e"]["value"], d["clocks"])
PY
