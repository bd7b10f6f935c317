#!/usr/bin/env python
"""Text summary of one `ncu --set full --import-source on` capture: headline metrics, pipe utilisation, stall
reasons, opcode mix and the hottest SASS lines.

    python tools/ncu_full_summary.py gpurun_out/r1_attn_bwd_full.ncu-rep [scores]  > profiles/r1_ncu_attn_bwd_full.txt
`scores` (optional) = number of attention scores the launch processes, to print instructions per score.
"""
import collections, csv, io, re, subprocess, sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    scores = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = page(rep, "raw")
    hdr, vals = raw[0], raw[2] if len(raw) > 2 else raw[1]
    d = dict(zip(hdr, vals))
    print(d.get("Kernel Name", "")[:160])
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for k in keys:
        for h in hdr:
            if h == k or h.endswith("." + k):
                print(f"  {k:68s} {d[h]}")
                break
    print("  stall reasons (warps per issue-active cycle):")
    st = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), num(d[h])) for h in hdr
          if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for n, v in sorted(st, key=lambda t: -t[1])[:10]:
        print(f"    {n:28s} {v:6.2f}")
    src = page(rep, "source")
    hi = next(i for i, r in enumerate(src[:10]) if "Source" in r)
    sh = {h: i for i, h in enumerate(src[hi])}
    c_src, c_ex, c_smp = sh["Source"], sh["Instructions Executed"], sh["# Samples"]
    ops, tot, lines = collections.Counter(), 0, []
    for r in src[hi + 1:]:
        if len(r) <= max(c_src, c_ex, c_smp):
            continue
        n = int(num(r[c_ex]))
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", r[c_src])
        if not m:
            continue
        op = m.group(2)
        op = ".".join(op.split(".")[:2]) if op.startswith(("SYNCS", "MUFU", "LDTM", "STTM", "F2FP", "IMAD")) else op.split(".")[0]
        ops[op] += n
        tot += n
        lines.append((int(num(r[c_smp])), r[c_src].strip()[:90]))
    print(f"  opcode mix ({tot} warp instructions incl. predicated-off" + (f", {tot * 32 / scores:.1f} per score" if scores else "") + "):")
    for op, n in ops.most_common(16):
        extra = f"  {n * 32 / scores:5.2f} / score" if scores else ""
        print(f"    {op:18s} {100 * n / tot:5.1f}%{extra}")
    print("  hottest SASS lines (stall samples):")
    for n, l in sorted(lines, key=lambda t: -t[0])[:14]:
        print(f"    {n:7d}  {l}")


if __name__ == "__main__":
    main()
