#!/usr/bin/env python
"""SURVEY.md §8d config 5 / BASELINE.json configs[4]: stand-alone SelfAttention + SpectralNorm kernel sweep.

Attention block forward and backward over N in {256..4096} x C in {64..512} (B chosen so that B * N >= 65 536; X ~ N(0,1),
1x1 weights ~ N(0, 1/C) with the logits kept at a few units, gamma = 0.5), BF16_TC mode, L2 flushed between timed
launches, CUDA events; algorithmic FLOPs of SURVEY.md §8d.  Spectral norm over the in-model shapes and the four
roofline shapes with the plan's phase times.  Writes one JSON document (stdout or --out)."""
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

import bench  # noqa: E402  (time_cuda, attn_flops, peaks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    import sagan_b200.functional as F
    from sagan_b200 import MATH_BF16_TC
    pk = bench.peaks()
    flush = torch.zeros(128 * 1024 * 1024, device="cuda")
    out = {"peaks": pk, "attention": [], "spectral_norm": []}
    Ns = [256, 1024, 4096] if a.quick else [256, 512, 1024, 2048, 4096]
    Cs = [64, 512] if a.quick else [64, 128, 256, 512]
    g = torch.Generator(device="cuda").manual_seed(0)
    for C in Cs:
        for N in Ns:
            B = max(1, 65536 // N)
            d, dv = C // 8, C // 2
            x = torch.randn(B, N, C, device="cuda", generator=g, requires_grad=True)
            mk = lambda *sh: (torch.randn(*sh, device="cuda", generator=g) / np.sqrt(sh[0])).requires_grad_(True)
            w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C),
                 torch.tensor(0.5, device="cuda", requires_grad=True)]
            with torch.no_grad():
                w[0] *= 1.5 ** 0.5 / d ** 0.25
                w[2] *= 1.5 ** 0.5 / d ** 0.25
            dy = torch.randn(B, N, C, device="cuda", generator=g)
            try:
                with torch.no_grad():
                    t_f = bench.time_cuda(lambda: F.attention(x, *w, MATH_BF16_TC), a.iters, flush)
                y = F.attention(x, *w, MATH_BF16_TC)
                t_b = bench.time_cuda(lambda: torch.autograd.grad(y, [x] + w, dy, retain_graph=True), max(2, a.iters // 2), flush)
            except Exception as e:      # noqa: BLE001
                out["attention"].append(dict(B=B, N=N, C=C, error=str(e)[:200]))
                continue
            ff, fb = bench.attn_flops(B, N, C, False), bench.attn_flops(B, N, C, True)
            out["attention"].append(dict(B=B, N=N, C=C, fwd_us=t_f * 1e6, bwd_us=t_b * 1e6, fwd_tflops=ff / t_f / 1e12,
                                         bwd_tflops=fb / t_b / 1e12, fwd_frac_of_bf16_peak=ff / t_f / 1e12 / pk["tc_burst"],
                                         bwd_frac_of_bf16_peak=fb / t_b / 1e12 / pk["tc_burst"]))
            print(out["attention"][-1], file=sys.stderr, flush=True)
            del x, w, dy, y
    shapes = [("512x4608", [(512, 4608)]), ("4096x4096", [(4096, 4096)]), ("4096x16384", [(4096, 16384)])]
    if not a.quick:
        shapes.append(("8192x32768", [(8192, 32768)]))
    shapes.append(("church64_G (13 matrices, one launch)", [(4096, 128), (256, 2048), (128, 1024), (64, 512), (32, 256), (4, 32),
                                                            (4, 32), (16, 32), (32, 16), (2, 16), (2, 16), (8, 16), (16, 8)]))
    shapes.append(("church64_D (8 matrices, one launch)", [(16, 48), (32, 256), (64, 512), (128, 1024), (2, 16), (2, 16), (8, 16),
                                                           (16, 8)]))
    for name, shp in shapes:
        Ws = [torch.randn(K, R, device="cuda") * 0.02 for R, K in shp]
        us = [torch.randn(1, R, device="cuda") for R, K in shp]
        grp = F.SpectralNormGroup(Ws, [u / u.norm() for u in us], 1)
        t = bench.time_cuda(grp.run, a.iters, flush)
        ph = grp.phase_times_ms()
        out["spectral_norm"].append(dict(shape=name, us=t * 1e6, algorithmic_bytes=grp.algorithmic_bytes,
                                         GBps=grp.algorithmic_bytes / t / 1e9, frac_of_hbm_peak=grp.algorithmic_bytes / t / 1e9 / pk["hbm"],
                                         phase_us=[p * 1e3 for p in ph]))
        print(out["spectral_norm"][-1], file=sys.stderr, flush=True)
        del grp, Ws, us
    txt = json.dumps(out, indent=1)
    if a.out:
        open(a.out, "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()
