#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.

    python tools/launch_summary.py gpurun_out/launches.csv [--skip N] [--take M]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    take = int(sys.argv[sys.argv.index("--take") + 1]) if "--take" in sys.argv else None
    hdr, rows = None, []
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                rows.append((d["Kernel Name"], float(d["Metric Value"].replace(",", ""))))
            except ValueError:
                pass
    rows = rows[skip:skip + take] if take else rows[skip:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = k.split("(")[0].replace("void ", "").replace("sagan::", "")[:60]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.1f} us total (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.1f} us {v[0]:5d}x {v[1] / tot * 100:5.1f}%  {k}")


if __name__ == "__main__":
    main()
