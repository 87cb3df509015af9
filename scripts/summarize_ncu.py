#!/usr/bin/env python3
"""Summaries for profiles/: (1) launch list csv -> markdown table, (2) one --set full report -> metric list + ncu_traffic.json.
usage: summarize_ncu.py launches <launches.csv> <out.md>
       summarize_ncu.py full <report.ncu-rep> <out.txt> [frames_per_launch]
       summarize_ncu.py gemm <report.ncu-rep> <out.txt>      (tensor-pipe utilisation of the captured GEMM launches)"""
import csv, io, json, subprocess, sys, collections, os

def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = r[ix["Kernel Name"]].split("(")[0][:120]
        agg[k][0] += 1; agg[k][1] += us
    tot = sum(v[1] for v in agg.values()); n = sum(v[0] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src}); {n} launches, {tot/1e3:.3f} ms total (cold-cache, serialised)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {c} | {t:.1f} | {t/c:.1f} | {100*t/tot:.1f}% |\n")
    print(open(dst).read()[:1500])

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_lsu.sum"]

def full(rep, dst, fpl):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        for li, r in enumerate(rows[2:]):
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            f.write(f"## launch {li}: {d.get('Kernel Name')}\n")
            for k in hdr:
                if k in KEEP or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                    f.write(f"{k:90s} {d[k]:>18s} {u[k]}\n")
            if li == 0:
                def val(k):
                    v = float(d[k].replace(",", "")); un = u[k]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(un, 1.0)
                sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
                from bench import kernel_source_hash
                tr = {"kernel": d.get("Kernel Name"), "frames_per_launch": fpl, "kernel_hash": kernel_source_hash(),
                      "dram_bytes_read": val("dram__bytes_read.sum"),
                      "dram_bytes_write": val("dram__bytes_write.sum"), "launch_ms_under_ncu": val("gpu__time_duration.sum"),
                      "source": f"{dst} (ncu --set full, one {fpl}-frame launch, 0.6B, prompt 39 rows)"}
                json.dump(tr, open(os.path.join(os.path.dirname(dst) or ".", "ncu_traffic.json"), "w"), indent=1)
                print(tr)
    print(open(dst).read()[:2500])

def gemm(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    want = [k for k in hdr if k in KEEP or k.startswith("sm__pipe_tensor") or k.startswith("sm__inst_executed_pipe_tensor")
            or k in ("launch__grid_size", "sm__cycles_elapsed.avg", "smsp__inst_executed_pipe_uniform.sum")]
    with open(dst, "w") as f:
        f.write(f"# {rep}: tensor-pipe view of the captured launches (ncu --set full)\n")
        for li, r in enumerate(rows[2:]):
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            f.write(f"## launch {li}: {d.get('Kernel Name')}  grid {d.get('launch__grid_size')}\n")
            for k in want:
                f.write(f"{k:90s} {d[k]:>18s} {u[k]}\n")
    print(open(dst).read()[:3000])

if sys.argv[1] == "launches": launches(sys.argv[2], sys.argv[3])
elif sys.argv[1] == "gemm": gemm(sys.argv[2], sys.argv[3])
else: full(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 8)
