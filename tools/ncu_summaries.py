#!/usr/bin/env python
"""CPU-side readers for the ncu artefacts of tools/gpu_r02_evidence.sh (no GPU needed, only the ncu binary):

    python tools/ncu_summaries.py launches gpurun_out/r02_launches.csv "<title>"     # per-kernel time shares of the last frame
    python tools/ncu_summaries.py full "<title>" report.ncu-rep [kernel-regex]       # headline metrics of every captured launch
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe inst %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU data-pipe wavefronts % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("smsp__inst_executed.sum", "warp instructions"),
])


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")[:70]


def launches(path, title):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[start + 1:] if len(r) == len(hdr) and r[ix["Metric Name"]] == "gpu__time_duration.sum"]
    # the frames are identical eager forwards: keep the second half of the launches (the last frame)
    data = data[len(data) // 2:]
    agg = OrderedDict()
    for r in data:
        k = short(r[ix["Kernel Name"]])
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + v)
    total = sum(t for _, t in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches): %s" % title)
    print("# kernel, launches, us, share of the frame")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %4d  %8.1f  %4.1f %%" % (k, n, t, 100.0 * t / total))
    print("%-70s %4d  %8.1f" % ("total", sum(n for n, _ in agg.values()), total))


def full(title, report, pattern=None):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("== %s" % title)
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        if pattern and not re.search(pattern, name):
            continue
        print("  %s" % short(name))
        for m, label in METRICS.items():
            if m in ix:
                print("    %-42s %s %s" % (label, r[ix[m]], units[ix[m]]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
