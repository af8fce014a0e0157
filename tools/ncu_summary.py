#!/usr/bin/env python
"""Turns an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text/JSON
summaries committed under profiles/.  Usage:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_solve_v3 [--batch 4096]"""
import csv
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__shared_mem_per_block_dynamic",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    hdr = rows[hi]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = {}
    for r in rows[hi + 1:]:
        if r[0].strip().isdigit():
            try:
                k = (int(r[0]), r[1].strip()[:96])
                a = agg.setdefault(k, [0, 0])
                a[0] += int(r[ci]); a[1] += int(r[cs])
            except (ValueError, IndexError):
                pass
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    lines = ["%5.1f%% inst %5.1f%% stall-samples  L%-4d %s" % (100 * a[0] / ti, 100 * a[1] / ts, k[0], k[1])
             for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]]
    return ti, lines


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 4096
    hdr, units, launches = raw(rep)
    r = launches[0]
    vals = {h: (v, u) for h, u, v in zip(hdr, units, r)}
    name = vals.get("Kernel Name", ("?", ""))[0]
    summary = {"kernel": name, "report": rep, "batch": batch}
    txt = ["ncu --set full summary of %s (first captured launch of %s)" % (rep, name), ""]
    for k in WANT:
        if k in vals:
            txt.append("%-64s %s %s" % (k, vals[k][0], vals[k][1]))
            try:
                summary[k] = float(vals[k][0].replace(",", ""))
            except ValueError:
                summary[k] = vals[k][0]
    stalls = sorted(((float(v), h) for h, v in zip(hdr, r)
                     if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")), reverse=True)
    txt += ["", "warp stall reasons (warps stalled per issue-active cycle):"]
    for v, h in stalls[:8]:
        txt.append("  %6.3f %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    def to_bytes(key):
        v, u = vals.get(key, ("0", "byte"))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return float(v.replace(",", "")) * mult
    dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    summary["dram_bytes_per_launch"] = dram
    summary["dram_bytes_per_stack"] = dram / batch
    summary["inst_per_stack"] = summary.get("smsp__inst_executed.sum", 0) / batch
    txt += ["", "dram bytes per launch %.0f (%.0f per stack); warp instructions per stack %.0f"
            % (dram, dram / batch, summary["inst_per_stack"])]
    ti, lines = source(rep, 40)
    txt += ["", "hottest source lines (opensot_b200/csrc/kernel source and headers it inlines):"] + lines
    with open(outbase + ".txt", "w") as f:
        f.write("\n".join(txt) + "\n")
    with open(outbase + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    if "--index-key" in sys.argv:
        # profiles/ncu_summary.json: captures by "shape:dtype:batch"; bench.py takes `roofline.traffic` from the entry
        # that matches the kernel, shape, precision and batch it measured (null otherwise)
        import os
        key = sys.argv[sys.argv.index("--index-key") + 1]
        idx_path = os.path.join(os.path.dirname(outbase) or ".", "ncu_summary.json")
        idx = {}
        if os.path.exists(idx_path):
            with open(idx_path) as f:
                idx = json.load(f)
        if "captures" not in idx:
            idx = {"captures": {}}
        idx["captures"][key] = {"kernel": name, "dram_bytes_per_launch": dram, "gpu_time_us": summary.get("gpu__time_duration.sum"),
                                "source": "%s.json (ncu --set full --clock-control none, one launch)" % outbase}
        with open(idx_path, "w") as f:
            json.dump(idx, f, indent=1)
    print("\n".join(txt[:40]))


if __name__ == "__main__":
    main()
