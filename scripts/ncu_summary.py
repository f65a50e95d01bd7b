#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  launches CSV (--metrics gpu__time_duration.sum)  -> per-kernel totals and shares of the step
  .ncu-rep (--set full)                             -> the handful of metrics the roofline argument needs
Usage: python scripts/ncu_summary.py launches <csv> <out.txt> | rep <file.ncu-rep> <out.txt>"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        u = r[iu]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = re.sub(r"\(.*", "", r[ik]).strip()
        tot[name][0] += 1
        tot[name][1] += ms
    total = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel totals from {path} (ncu gpu__time_duration.sum; cold-cache, serialised launches: compare SHARES)\n")
        f.write(f"# total {total:.3f} ms over {sum(v[0] for v in tot.values())} launches\n")
        f.write(f"{'kernel':60s} {'launches':>9s} {'ms':>12s} {'share':>8s} {'avg_ms':>10s}\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:60]:60s} {n:9d} {ms:12.3f} {ms / total:8.4f} {ms / n:10.4f}\n")


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# selected metrics from {path} (ncu --set full --clock-control none)\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')][:120]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:90s} {r[i]:>22s} {units[i]}\n")


def ratios(path, out):
    """DRAM bytes per algorithmic byte of the launches captured from scripts/ncu_targets.py (cfg4: N = 100000,
    bootstrap sample 0 of seed 17000, m cells) -> profiles/ncu_traffic_ratios.json, read by bench.py for `traffic`."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import raceid3_stemid2_package_b200 as rid
    N = 100000
    m = len(rid.bootstrap_samples(N, 1, 17000)[0])
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    ik, ir, iw, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    res = {"_source": os.path.basename(path), "_workload": f"scripts/ncu_targets.py: N={N}, G=3000, subset m={m}"}
    alg = {"k_scan": float(m) * m * 4, "k_compact_sym": (0.5 * m * (m + 1) + float(m) * m) * 4}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ik])
        base = re.match(r"[A-Za-z0-9_]+", name).group(0)
        try:
            dram = float(r[ir]) * scale[rows[1][ir]] + float(r[iw]) * scale[rows[1][iw]]
        except (ValueError, KeyError):
            continue
        if dram != dram or base in res:
            continue
        if base in alg:
            if base == "k_scan" and abs(dram / alg[base] - 1.0) > 0.2:
                continue  # an incremental pass (fewer rows): keep the full-size launches only
            res[base] = {"dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": alg[base],
                         "dram_per_algorithmic": dram / alg[base], "kernel": name[:80]}
        elif base.startswith("k_gram_i8"):
            res[base] = {"dram_bytes_per_launch_cfg4": dram, "kernel": name[:80]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rawcsv(path, out):
    """The wide CSV of `ncu -i x.ncu-rep --page raw --csv` (made on the GPU box when the report itself is too large to bring
    back) -> the metrics the roofline argument needs, per captured launch; launches under 0.05 ms are left out."""
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
    seen = defaultdict(int)
    with open(out, "w") as f:
        f.write(f"# selected metrics from {path} (ncu --set full --clock-control none --kernel-name-base demangled)\n")
        for r in rows[2:]:
            if len(r) <= ik:
                continue
            name = re.sub(r"\(.*", "", r[ik]).strip()
            it = hdr.index("gpu__time_duration.sum")
            try:
                ms = float(r[it].replace(",", "")) * tscale.get(units[it], 1e-6)
            except ValueError:
                ms = 0.0
            seen[name] += 1
            if ms < 0.05:
                continue
            f.write(f"== {name}  (launch #{seen[name]})\n")
            tot = 0.0
            for k in RAW_KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"   {k:100s} {r[i]:>14s} {units[i]}\n")
                    if k.startswith("dram__bytes"):
                        try:
                            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
                        except ValueError:
                            pass
            if ms > 0 and tot > 0:
                f.write(f"   {'-> DRAM bytes / duration':100s} {tot / 1e9 / ms:14.2f} TB/s (GB per ms)\n")


if __name__ == "__main__":
    {"launches": launches, "rep": rep, "ratios": ratios, "rawcsv": rawcsv}[sys.argv[1]](sys.argv[2], sys.argv[3])
