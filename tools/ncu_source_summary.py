#!/usr/bin/env python
"""Stall-sample summary of one kernel from an ncu report that was captured with --import-source on (runs anywhere ncu is
installed, no GPU needed):

    python tools/ncu_source_summary.py gpurun_out/prof_conv_32_occ2.ncu-rep [--top 40] [--regions 640,1600,2000,3920]

Prints (1) total samples and the stall mix of the whole kernel, (2) the same per address region - `--regions` takes
instruction indices that split the SASS listing (e.g. the role boundaries of a warp-specialised kernel: find them with
--markers, which lists barrier / MMA / TMEM / TMA / EXIT instructions with their indices), (3) the hottest instructions.
"""
import argparse, csv, io, subprocess, sys

MARKERS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UBLKCP", "UTMALDG", "LDTM", "STTM", "BAR.SYNC", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "ELECT",
           "UTCATOMSWS", "EXIT", "WARPSYNC", "REDUX", "RED.", "ATOM")


def load(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
    kernel = next((r[1] for r in rows[:start] if r and r[0] == "Kernel Name"), "?")
    return kernel, hdr, data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--regions", default="")
    ap.add_argument("--markers", action="store_true")
    a = ap.parse_args()
    kernel, hdr, data = load(a.report)
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    S = lambda r: int(r[ix["# Samples"]])
    total = sum(S(r) for r in data)
    print("kernel:", kernel[:100])
    print("instructions: %d, samples: %d" % (len(data), total))

    def mix(rows_):
        agg = sorted(((sum(int(r[ix[h]]) for r in rows_), h[6:]) for h in stalls), reverse=True)
        n = max(1, sum(v for v, _ in agg))
        return ", ".join("%s %.0f%%" % (h, 100.0 * v / n) for v, h in agg[:5] if v)

    print("whole kernel:", mix(data))
    if a.markers:
        for i, r in enumerate(data):
            if any(m in r[ix["Source"]] for m in MARKERS):
                print("  [%5d] %-70s samples %-6d executed %s" % (i, r[ix["Source"]].strip()[:70], S(r), r[ix["Instructions Executed"]]))
    if a.regions:
        cuts = [0] + [int(v) for v in a.regions.split(",")] + [len(data)]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            seg = data[lo:hi]
            n = sum(S(r) for r in seg)
            ins = sum(int(r[ix["Instructions Executed"]]) for r in seg)
            print("region [%5d,%5d): samples %6d (%4.1f%%)  warp-instructions %10d  | %s" % (lo, hi, n, 100.0 * n / max(1, total), ins, mix(seg)))
    print("hottest instructions:")
    for i in sorted(sorted(range(len(data)), key=lambda i: -S(data[i]))[:a.top]):
        r = data[i]
        top2 = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("  [%5d] %-64s %6d  %s" % (i, r[ix["Source"]].strip()[:64], S(r), ", ".join("%s %d" % (h, v) for v, h in top2 if v)))


if __name__ == "__main__":
    sys.exit(main())
