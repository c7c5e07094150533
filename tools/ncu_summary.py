#!/usr/bin/env python
"""Extract the judged numbers from an .ncu-rep (run here, no GPU needed) into profiles/<name>.json + .txt.

    python tools/ncu_summary.py gpurun_out/k2_cfg3.ncu-rep profiles/r01_k2_cfg3 [--workload cfg3]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg",
    "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_uniform_realtime.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    workload = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for i, name in enumerate(hdr):
            short = name.split(".", 2)[-1] if name.split(".")[0] in ("TPC", "SM_A", "SM_B", "SM_C", "LTS", "FBPA", "GPC") else name
            for k in KEYS + ["Kernel Name", "ID"]:
                if short == k or name == k:
                    d[k] = (r[i] + (" " + units[i] if units[i] else "")).strip()
        res.append(d)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for d in res:
            f.write(f"== {d.get('Kernel Name', '?')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"  {k:90s} {d[k]}\n")
    print(open(out + ".txt").read())
    if workload and res:
        def num(s):
            v, u = s.split()[0], (s.split() + [""])[1]
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)
        k2 = [d for d in res if "filter_mma" in d.get("Kernel Name", "")]
        if k2:
            t = num(k2[0]["dram__bytes_read.sum"]) + num(k2[0]["dram__bytes_write.sum"])
            import os
            p = os.path.join(os.path.dirname(out), "k2_traffic.json")
            cur = json.load(open(p)) if os.path.exists(p) else {}
            cur[workload] = t
            json.dump(cur, open(p, "w"), indent=1)
            print("traffic", workload, t)


if __name__ == "__main__":
    main()
