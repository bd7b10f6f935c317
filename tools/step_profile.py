#!/usr/bin/env python
"""Warm per-kernel times of one church64 training step (eager, torch.profiler / CUPTI): cheaper than an ncu launch
list and without its cold-cache serialisation.  Groups launches by (kernel, grid).

    python tools/step_profile.py [--steps 3] [--top 50]
"""
import argparse, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=50)
    a = ap.parse_args()
    import bench
    from sagan_b200 import MATH_BF16_TC, _lib, nn as snn
    from sagan_b200.trainer import Trainer
    _lib.load()
    snn.set_default_math_mode(MATH_BF16_TC)
    cfg = dict(bench.CHURCH64)
    B = cfg["batch_size"]
    tr = Trainer(cfg, global_batch_size=B, steps_per_epoch=126227 // B, seed=0)
    x = torch.rand(B, 64, 64, 3, device="cuda") * 2 - 1
    for _ in range(3):
        tr.train_step(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            tr.train_step(x)
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    total = 0.0
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.device_time_total > 0:
            name = ev.name.replace("sagan::", "").replace("void ", "")[:60]
            k = name
            c = agg.setdefault(k, [0, 0.0])
            c[0] += 1
            c[1] += ev.device_time_total
            total += ev.device_time_total
    print(f"{total / a.steps:9.1f} us of kernel time per step (warm, eager; {a.steps} steps)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: a.top]:
        print(f"{t / a.steps:9.1f} us/step {n / a.steps:6.1f}x {t / n:8.1f} us each {100 * t / total:5.1f}%  {k}")


if __name__ == "__main__":
    main()
