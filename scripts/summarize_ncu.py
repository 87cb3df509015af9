#!/usr/bin/env python3
"""Summarise ncu outputs into small text files for profiles/.

  launches <launches.csv> <out.md>      : per-kernel launch count / total / share from the
                                          `--metrics gpu__time_duration.sum` launch list
  full <report.ncu-rep> <out.txt>       : key metrics of the first kernel in an `ncu --set full` report
"""
import csv, subprocess, sys, collections, re


def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, v * scale))
    agg = collections.OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path}); {len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {c} | {us:.1f} | {us/c:.1f} | {100*us/tot:.1f}% |\n")
    print(open(out).read())


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for k, vals in enumerate(rows[2:]):
            d = dict(zip(hdr, zip(units, vals)))
            f.write(f"## launch {k}: {d.get('Kernel Name', ('', '?'))[1]}\n")
            for key in KEYS:
                if key in d:
                    f.write(f"{key:72s} {d[key][1]:>18s} {d[key][0]}\n")
            for h in hdr:
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    f.write(f"{h:72s} {d[h][1]:>18s}\n")
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
