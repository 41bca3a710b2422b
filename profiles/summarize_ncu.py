#!/usr/bin/env python
"""Summarise ncu outputs brought back from gpurun into small text files that are committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_X.csv  > profiles/r01_X_launches.txt
    python profiles/summarize_ncu.py full     gpurun_out/prof_X.ncu-rep  > profiles/r01_X_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def short(name, n=110):
    name = name.replace("void ", "").replace("mr::", "")
    return name if len(name) <= n else name[: n - 3] + "..."


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        k = short(r["Kernel Name"])
        a = per.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes")
    print(f"{'share':>7} {'launches':>8} {'total_us':>12} {'avg_us':>10}  kernel")
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * us / total:6.2f}% {n:8d} {us:12.1f} {us / n:10.1f}  {k}")
    print(f"total {total:.1f} us over {sum(n for n, _ in per.values())} launches")


def full(path):
    if path.endswith(".csv"):   # already exported on the GPU box (tools/ncu_capture.sh)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none ({path}); one block per profiled launch")
    for r in rows[2:]:
        print(f"kernel: {short(r[hdr.index('Kernel Name')], 160)}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]:>16s} {units[i]}")
        if "dram__bytes_read.sum" in hdr:
            def val(k):
                i = hdr.index(k)
                v = float(r[i].replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[units[i]]
            tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            i = hdr.index("gpu__time_duration.sum")
            t = float(r[i].replace(",", "")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(units[i].replace("econd", ""), 1e-6)
            print(f"  {'dram traffic (read+write) bytes':70s} {tr:16.0f} byte   -> {tr / t / 1e9:.1f} GB/s under ncu")
        print()


def source(path, top=25):
    """Hottest SASS instructions (warp-stall samples) of a `--page source --csv` export (.csv or .csv.gz)."""
    import gzip
    text = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rows = list(csv.reader(io.StringIO(text)))
    hdr = rows[1]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [(int(r[isamp]), int(r[iex]), r[isrc].strip()) for r in rows[2:] if len(r) > iex]
    tot, totex = sum(d[0] for d in data), sum(d[1] for d in data)
    print(f"# ncu source page ({path}): {rows[0][1][:120]}")
    print(f"# {len(data)} SASS instructions, {totex} warp-instructions executed, {tot} stall samples")
    print(f"{'samples':>8} {'share':>7} {'executed':>12}  instruction")
    for smp, ex, src in sorted(data, reverse=True)[:top]:
        print(f"{smp:8d} {100.0 * smp / max(tot, 1):6.2f}% {ex:12d}  {src[:100]}")
    ops = collections.Counter()
    for smp, ex, src in data:
        op = src.split()[1] if src.startswith("@") and len(src.split()) > 1 else src.split()[0]
        ops[op.split(".")[0]] += ex
    print("# executed warp-instructions by opcode: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


if __name__ == "__main__":
    {"launches": launches, "full": full, "source": source}[sys.argv[1]](sys.argv[2])
