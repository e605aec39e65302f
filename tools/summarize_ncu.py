#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md>      # per-kernel time shares of ONE bench step
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md>         # key metrics of every captured launch
  python tools/summarize_ncu.py traffic <report.ncu-rep> <out.json>    # mean DRAM bytes per launch, per kernel
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("lb::", "").replace("void ", "")
        seq.append((name, v))
    # one align() = one prep_source_kernel launch (the stream-ordered execution has no single align kernel)
    idx = [i for i, s in enumerate(seq) if "prep_source" in s[0]]
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): one bench step\n\n")
        f.write("Source: `%s` (%d launches captured); the step between the last two align() calls (prep_source_kernel launches).\n" % (path, len(seq)))
        f.write("Durations under ncu are serialised (and cold-cache unless --cache-control none): compare SHARES.\n\n")
        if len(idx) >= 2:
            step = seq[idx[-2] + 1: idx[-1] + 1]
            agg = collections.OrderedDict()
            for n, v in step:
                a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
            tot = sum(a[1] for a in agg.values())
            f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
            for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, n, t, 100 * t / tot))
            f.write("| **sum** | %d | %.1f | |\n" % (sum(a[0] for a in agg.values()), tot))
    print("wrote", out)


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % rep)
        for r in rows[2:]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[ix["Kernel Name"]])
            for k in KEYS:
                if k in ix:
                    f.write("| %s | %s | %s |\n" % (k, r[ix[k]], units[ix[k]]))
            f.write("\n")
    print("wrote", out)


def traffic(rep, out):
    """profiles/traffic.json: what bench.py reports as roofline.traffic (dram read + write bytes per launch)."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("lb::", "").replace("void ", "")
        name = re.sub(r"<.*", "", name)
        rd = float(r[ix["dram__bytes_read.sum"]].replace(",", ""))
        wr = float(r[ix["dram__bytes_write.sum"]].replace(",", ""))
        dur = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
        a = acc.setdefault(name, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += rd; a[2] += wr; a[3] += dur
    res = {k: {"launches_captured": n, "dram_bytes_read_per_launch": rd / n, "dram_bytes_write_per_launch": wr / n,
               "dram_bytes_per_launch": (rd + wr) / n, "ncu_duration_ns_per_launch": d / n}
           for k, (n, rd, wr, d) in acc.items()}
    res["_source"] = "ncu --set full --clock-control none capture %s" % rep
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
