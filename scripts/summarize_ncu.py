#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep + launches_*.csv into the small text summaries committed under profiles/.
Usage: python scripts/summarize_ncu.py <tag> [batch_frames=32]     (reads gpurun_out/, writes profiles/<tag>_*.md and profiles/latest_traffic.json)"""
import csv
import io
import os
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return [r for r in csv.reader(io.StringIO(out))]


def summarize_rep(rep, fh):
    rows = ncu_csv(rep, "raw")
    hdr = rows[0]; units = rows[1]
    for r in rows[2:]:
        fh.write(f"### {r[hdr.index('Kernel Name')][:120]}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in hdr:
                fh.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
        stalls = [(float(r[i] or 0), k) for i, k in enumerate(hdr) if k.startswith("smsp__average_warp_latency_issue_stalled") or
                  (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"))]
        stalls.sort(reverse=True)
        if stalls:
            fh.write("\nTop stall reasons (warps per issue-active cycle): " + ", ".join(f"{k.split('stalled_')[1].split('_per')[0]}={v:.2f}" for v, k in stalls[:6]) + "\n")
        fh.write("\n")
    src = ncu_csv(rep, "source", ("--print-source", "sass"))
    if len(src) > 2:
        hdr = src[1]
        try:
            ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
        except ValueError:
            return
        c = Counter()
        for r in src[2:]:
            if len(r) != len(hdr) or not r[ia].isdigit():
                continue
            op = r[isrc].split()
            o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
            c[o] += int(r[ia])
        tot = sum(c.values())
        if tot:
            fh.write("SASS opcode mix (executed warp instructions, all captured launches): " +
                     ", ".join(f"{o} {100 * n / tot:.1f}%" for o, n in c.most_common(14)) + "\n\n")
            if c.get("MUFU"):
                fh.write(f"warp instructions per MUFU.RCP (= per evaluated voxel-frame in k_integrate*): {tot / c['MUFU']:.1f}\n\n")


def summarize_launches(path, fh):
    per = defaultdict(list)
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            per[name].append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in per.values())
    fh.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        fh.write(f"| {k[:90]} | {len(v)} | {sum(v) / 1e3:.1f} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f}% |\n")
    fh.write("\n(ncu serialises launches with cold caches: compare SHARES, not absolutes.)\n\n")


def main():
    tag = sys.argv[1]
    g = os.path.join(ROOT, "gpurun_out"); p = os.path.join(ROOT, "profiles"); os.makedirs(p, exist_ok=True)
    with open(os.path.join(p, f"{tag}_ncu_summary.md"), "w") as fh:
        fh.write(f"# ncu summary {tag}\n\nCommand: see scripts/gpu_r02b.sh (bench.py --steps 2 --warmup 1 --scene-frames 96 --frames-per-step 96 --no-cpu --no-seg under ncu).\n\n")
        lp = os.path.join(g, f"launches_{tag}.csv")
        if os.path.exists(lp):
            fh.write("## launch list (gpu__time_duration.sum, --clock-control none)\n\n"); summarize_launches(lp, fh)
        for f in sorted(os.listdir(g)):
            if f.endswith(f"_{tag}.ncu-rep"):
                fh.write(f"## {f} (ncu --set full)\n\n"); summarize_rep(os.path.join(g, f), fh)
    # per-kernel numbers bench.py quotes in its roofline block (profiles/latest_traffic.json)
    import json
    tr = {}; issue = {}; inst = {}; mufu = {}
    for f in sorted(os.listdir(g)):
        if f.endswith(f"_{tag}.ncu-rep"):
            rep = os.path.join(g, f)
            rows = ncu_csv(rep, "raw"); hdr = rows[0]
            for r in rows[2:]:
                name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "").split("<")[0]
                def val(k):
                    v = float(r[hdr.index(k)].replace(",", "")); u = rows[1][hdr.index(k)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                tr.setdefault(name, []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
                if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr:
                    issue.setdefault(name, []).append(val("smsp__issue_active.avg.pct_of_peak_sustained_active"))
                if "smsp__inst_executed.sum" in hdr:
                    inst.setdefault(name, []).append(val("smsp__inst_executed.sum"))
            # executed MUFU.RCP per launch = evaluated voxel-frames / 32 in the integrate kernels (one rcp per voxel-frame)
            src = ncu_csv(rep, "source", ("--print-source", "sass"))
            if len(src) > 2 and "k_integrate" in f.replace("prof_integrate", "k_integrate"):
                hdr2 = src[1]
                try:
                    ia = hdr2.index("Instructions Executed"); isrc = hdr2.index("Source")
                    c = Counter()
                    for r in src[2:]:
                        if len(r) == len(hdr2) and r[ia].isdigit():
                            op = r[isrc].split(); o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]; c[o] += int(r[ia])
                    if c.get("MUFU"):
                        for name in list(inst):
                            if name.startswith("k_integrate"):
                                mufu[name] = sum(c.values()) / c["MUFU"]
                except ValueError:
                    pass
    if tr:
        kint = next((k for k in tr if k.startswith("k_integrate")), None)
        with open(os.path.join(p, "latest_traffic.json"), "w") as fh:
            json.dump({"tag": tag, "source": "ncu --set full --clock-control none, bench.py --steps 2 --warmup 1 --scene-frames 96 --frames-per-step 96 (default batch)",
                       "batch": int(sys.argv[2]) if len(sys.argv) > 2 else 32, "integrate_kernel": kint,
                       "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in tr.items()},
                       "issue_active": {k: sum(v) / len(v) for k, v in issue.items()},
                       "warp_inst_per_launch": {k: sum(v) / len(v) for k, v in inst.items()},
                       "thread_inst_per_voxel_frame": mufu}, fh, indent=1)
    for f in (f"bench_{tag}.json", f"bench_ref_{tag}.json", f"gpu_{tag}.txt", f"pytest_gpu_{tag}.log", f"smoke_{tag}.log"):
        s = os.path.join(g, f)
        if os.path.exists(s):
            with open(s) as a, open(os.path.join(p, f), "w") as b:
                b.write(a.read())
    print("wrote", os.path.join(p, f"{tag}_ncu_summary.md"))


if __name__ == "__main__":
    main()
