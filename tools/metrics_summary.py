#!/usr/bin/env python
"""Pivot an `ncu --metrics ... --csv` log: one line per kernel launch with the requested metrics.

    python tools/metrics_summary.py gpurun_out/r1_metrics_attn.csv
"""
import collections
import csv
import sys


def load(path):
    hdr, out = None, collections.OrderedDict()
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            k = (d["ID"], d["Kernel Name"].split("(")[0].replace("void ", "").replace("sagan::", "")[:48], d["Grid Size"])
            try:
                out.setdefault(k, {})[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])
            except ValueError:
                pass
    return out


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rows = load(sys.argv[1])
    for (i, name, grid), m in rows.items():
        dr = to_bytes(*m.get("dram__bytes_read.sum", (0, "byte")))
        dw = to_bytes(*m.get("dram__bytes_write.sum", (0, "byte")))
        l2 = to_bytes(*m.get("lts__t_bytes.sum", (0, "byte")))
        t, tu = m.get("gpu__time_duration.sum", (0, "ns"))
        t_us = t * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(tu, 1e-3)
        tp = m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
        xu = m.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
        iss = m.get("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
        print(f"{name:48s} grid {grid:14s} {t_us:9.1f} us  dram R {dr / 1e6:8.1f} MB W {dw / 1e6:8.1f} MB  "
              f"L2 {l2 / 1e6:9.1f} MB  tensor {tp:5.1f}%  xu {xu:5.1f}%  issue {iss:5.1f}%")


if __name__ == "__main__":
    main()
