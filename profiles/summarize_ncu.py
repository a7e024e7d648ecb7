#!/usr/bin/env python
"""Summarise the `ncu --set full` raw pages profiles/ncu_all.sh leaves in gpurun_out/ into tracked files:
   python profiles/summarize_ncu.py <tag> [outdir]     ->  <outdir>/ncu_<tag>.md, <outdir>/ncu_<tag>.json,
                                                           profiles/ncu_traffic.json (what bench.py's `traffic` reads)
Only numbers the report itself holds; nothing is measured here."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ["z_inv", "y_inv", "x_fused", "y_fwd", "z_fwd", "assemble_stage"]
PICK = {
    "time_ms": "gpu__time_duration.sum",
    "dram_read_gb": "dram__bytes_read.sum",
    "dram_write_gb": "dram__bytes_write.sum",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_pipe_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "inst_executed": "smsp__inst_executed.sum",
    "local_ld_st": "smsp__inst_executed_op_local_ld.sum",
    "occ_limit_regs": "launch__occupancy_limit_registers",
    "occ_limit_smem": "launch__occupancy_limit_shared_mem",
}
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "math_pipe_throttle", "barrier", "mio_throttle", "not_selected", "lg_throttle",
          "dispatch_stall", "branch_resolving", "no_instruction", "drain"]
UNIT_SCALE = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6, "Tbyte": 1e3}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(d, key):
    if key not in d:
        return None
    u, v = d[key]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * UNIT_SCALE.get(u, 1.0)


def main(tag, outdir):
    src = os.path.join(ROOT, "gpurun_out")
    out = {}
    for k in KERNELS:
        p = os.path.join(src, "raw_%s_%s.csv" % (tag, k))
        if not os.path.exists(p):
            continue
        d = read(p)
        e = {"kernel": d["Kernel Name"][1]}
        for name, key in PICK.items():
            e[name] = num(d, key)
        tot = 0.0
        st = {}
        for s in STALLS:
            v = num(d, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s)
            st[s] = v or 0.0
            tot += v or 0.0
        tot += num(d, "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio") or 0.0
        e["stall_pct"] = {s: round(100 * v / tot, 1) for s, v in st.items() if v / tot > 0.02}
        e["dram_gb"] = (e["dram_read_gb"] or 0) + (e["dram_write_gb"] or 0)
        e["dram_tbs"] = e["dram_gb"] / e["time_ms"]
        out[k] = e
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "ncu_%s.json" % tag), "w") as f:
        json.dump(out, f, indent=1)
    lines = ["# ncu --set full, one launch per kernel, 3-D MHD 512^3 RK4 on one B200 (`profiles/ncu_all.sh %s`)" % tag, "",
             "Times are cold-cache, serialised launches under the profiler (not bench values); traffic = dram__bytes_read + write.", "",
             "| kernel | time ms | DRAM read + write GB | TB/s | FP64 pipe % | issue % | smem pipe % | warps active % | regs | top stalls (% of warp samples) |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    for k, e in out.items():
        st = ", ".join("%s %.0f" % kv for kv in sorted(e["stall_pct"].items(), key=lambda kv: -kv[1])[:5])
        lines.append("| `%s` %s | %.3f | %.2f + %.2f = %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %d | %s |" % (
            k, e["kernel"].replace("|", "/")[:60], e["time_ms"], e["dram_read_gb"], e["dram_write_gb"], e["dram_gb"], e["dram_tbs"],
            e["fp64_pipe_pct"] or 0, e["issue_active_pct"] or 0, e["smem_pipe_pct"] or 0, e["warps_active_pct"] or 0, int(e["regs"] or 0), st))
    with open(os.path.join(outdir, "ncu_%s.md" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    traffic = {"512": {k: {"dram_bytes": e["dram_gb"] * 1e9, "time_ms_under_ncu": e["time_ms"],
                           "source": "%s/ncu_%s.json (raw page of gpurun_out/prof_%s_%s.ncu-rep)" % (os.path.relpath(outdir, ROOT), tag, tag, k)}
                       for k, e in out.items()},
               "workload": "3-D MHD 512^3 RK4, one B200, one launch per kernel"}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2"))
