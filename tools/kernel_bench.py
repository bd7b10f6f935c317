#!/usr/bin/env python
"""Stand-alone kernel runs for profiling (ncu-friendly: few launches) and for the config-5 sweep.

    python tools/kernel_bench.py attn --B 64 --N 4096 --C 16 --math bf16_tc --iters 3 [--bwd]
    python tools/kernel_bench.py sn --rows 4096 --cols 4096 --iters 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, iters, flush):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.mean(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["attn", "sn", "snmodel"])
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--C", type=int, default=16)
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--cols", type=int, default=4096)
    ap.add_argument("--math", default="bf16_tc")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    import sagan_b200.functional as F
    flush = None if a.no_flush else torch.zeros(64 * 1024 * 1024, device="cuda")
    out = {"op": a.op}
    if a.op == "attn":
        mode = F.MATH_BF16_TC if a.math == "bf16_tc" else F.MATH_FP32_STRICT
        B, N, C = a.B, a.N, a.C
        d, dv = C // 8, C // 2
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, N, C, device="cuda", generator=g, requires_grad=True)
        mk = lambda *s: (torch.randn(*s, device="cuda", generator=g) / np.sqrt(s[0])).requires_grad_(True)
        w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C),
             torch.tensor(0.5, device="cuda", requires_grad=True)]
        with torch.no_grad():      # un-scaled logits (layers.py:108) of a few units, std 1.5, whatever d is
            w[0] *= 1.5 ** 0.5 / d ** 0.25
            w[2] *= 1.5 ** 0.5 / d ** 0.25
        dy = torch.randn(B, N, C, device="cuda", generator=g)
        with torch.no_grad():
            t_f = timed(lambda: F.attention(x, *w, mode), a.iters, flush)
        fl_f = 2 * B * N * C * (2 * d + dv) + 2 * B * N * N * (d + dv) + 2 * B * N * dv * C
        out.update(shape=[B, N, C], math=a.math, fwd_ms=t_f[0] * 1e3, fwd_min_ms=t_f[1] * 1e3,
                   fwd_tflops=fl_f / t_f[0] / 1e12, fwd_exps_per_s=B * N * N / t_f[0])
        if a.bwd:
            y = F.attention(x, *w, mode)
            t_b = timed(lambda: torch.autograd.grad(y, [x] + w, dy, retain_graph=True), a.iters, flush)
            fl_b = 2 * B * N * N * (3 * d + 2 * dv) + 2 * (2 * B * N * C * (2 * d + dv) + 2 * B * N * dv * C)
            out.update(bwd_ms=t_b[0] * 1e3, bwd_tflops=fl_b / t_b[0] / 1e12, bwd_exps_per_s=B * N * N / t_b[0])
    elif a.op == "snmodel":
        # the 13 spectrally-normalised kernels of the church64 generator in ONE launch
        shapes = [(4096, 128), (256, 2048), (128, 1024), (64, 512), (32, 256), (4, 32), (4, 32), (16, 32), (32, 16),
                  (2, 16), (2, 16), (8, 16), (16, 8)]
        Ws = [torch.randn(K, R, device="cuda") * 0.02 for R, K in shapes]
        us = [torch.randn(1, R, device="cuda") for R, K in shapes]
        grp = F.SpectralNormGroup(Ws, [u / u.norm() for u in us], 1)
        t = timed(grp.run, a.iters, flush)
        out.update(ms=t[0] * 1e3, min_ms=t[1] * 1e3, algorithmic_bytes=grp.algorithmic_bytes,
                   gbs=grp.algorithmic_bytes / t[0] / 1e9, phases_ms=grp.phase_times_ms())
    else:
        R, K = a.rows, a.cols
        W = torch.randn(K, R, device="cuda") * 0.02
        u = torch.randn(1, R, device="cuda")
        grp = F.SpectralNormGroup([W], [u / u.norm()], 1)
        t = timed(grp.run, a.iters, flush)
        out.update(shape=[R, K], ms=t[0] * 1e3, min_ms=t[1] * 1e3, algorithmic_bytes=grp.algorithmic_bytes,
                   gbs=grp.algorithmic_bytes / t[0] / 1e9, gbs_best=grp.algorithmic_bytes / t[1] / 1e9,
                   traffic_3r1w_gbs=16.0 * R * K / t[0] / 1e9, phases_ms=grp.phase_times_ms())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
