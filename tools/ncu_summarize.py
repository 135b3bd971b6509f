#!/usr/bin/env python3
"""Turn ncu CSV output into the short tables kept under profiles/.

  launches <launches.csv> [last_n_launches]   per-kernel totals of `ncu --metrics gpu__time_duration.sum --csv --log-file`
  raw <raw.csv> <kernel-regex>                selected metrics of `ncu -i prof.ncu-rep --page raw --csv`, one column per launch
"""
import csv, re, sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def short(name):
    m = re.search(r"zkb_(?:block_|phased_)?kernel<(?:zkb::)?(\w+)", name)
    base = m.group(1) if m else re.sub(r"\(.*", "", name)
    if re.search(r"msm_\w+<Fp2T?<", name) and base.startswith("k_msm"):
        base += "<G2>"
    return base[:40]


def launches(path, last_n=None):
    rs = [r for r in rows(path) if r["Metric Name"] == "gpu__time_duration.sum"]
    if last_n:
        rs = rs[-int(last_n):]
    tot = OrderedDict()
    for r in rs:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("nsecond", "ns") else v * 1e3 if unit in ("msecond", "ms") else v
        k = short(r["Kernel Name"])
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + us)
    total = sum(t for _, t in tot.values())
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<40} n={n:4d} {t:10.1f} us {100 * t / total:5.1f}%")
    print(f"total {total:.0f} us over {len(rs)} launches")


def raw(path, pattern):
    rs = rows(path)
    cols = [r for r in rs if re.search(pattern, r["Kernel Name"])]
    print("| metric | unit | " + " | ".join(f"{short(r['Kernel Name'])} #{r['ID']}" for r in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    units = rs[0] if rs and all(v == "" or not re.match(r"^[\d.,]+$", v) for v in list(rs[0].values())[11:]) else {}
    for m in METRICS:
        if m in (cols[0] if cols else {}):
            print(f"| {m} | {units.get(m, '')} | " + " | ".join(r[m] for r in cols) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:4])
    else:
        raw(sys.argv[2], sys.argv[3])
