#!/usr/bin/env python
"""bench.py -- SAGAN (church64_attn) training throughput on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: the full training step of
/root/reference/sagan/main.py:171-211 (D update + G update, `update_ratio` = 1) at the
example_configs/church64_attn.py model (z 128, gf 16, df 16, 64x64, attention at 32x32 and 64x64),
per-GPU batch 64, synthetic uniform[-1,1) images and N(0,1) noise.

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM (CUDA-graph replay of
the whole step), `e2e` = images/s through the public Trainer API with the batch copied every step from
pinned host memory as the reference's raw uint8 records (sagan/dataset.py:27-40; decoded on the device by
the first node of the step graph) and the two loss scalars read back.  `roofline` is the dominant kernel
(self-attention at N = 4096) timed alone with CUDA events; `kernels` holds the other kernel lines
(down-sampled attention, the large-C block forward and backward, spectral norm); `cpu_baseline` is the CPU
oracle (torch-CPU restatement of the reference's TF graph; TensorFlow is not installable here) on the same
step and batch, a bounded number of steps.  `--impl reference` times that CPU oracle only.
`--config cond128 | res128` run the 128x128 class-conditional configs (not the headline).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CHURCH64 = dict(  # /root/reference/example_configs/church64_attn.py:14-27 (model / training keys)
    model="vanilla", z_dim=128, gf_dim=16, df_dim=16, lr_g=2e-4, lr_d=7e-4, decay_rate=0.99, use_attention=True,
    attn_dim_G=[32, 64], attn_dim_D=[8, 4], use_label=False, batch_size=64, loss="hinge_loss", update_ratio=1,
    img_size=64, num_classes=1)
# BASELINE.json configs[3]: 128x128 class-conditional SAGAN (vanilla builders with img_size 128, use_label, 1000 ImageNet
# classes; attention lands at the 32x32 and 64x64 maps of both networks).  `--config cond128`; not the headline workload.
COND128 = dict(CHURCH64, img_size=128, use_label=True, num_classes=1000)
CONFIGS = {
    "church64": (CHURCH64, "sagan_train_images_per_sec_64x64",
                 "church64_attn train step (D update + G update), per-GPU batch 64, 64x64x3, synthetic"),
    "cond128": (COND128, "sagan_train_images_per_sec_128x128_cond",
                "128x128 class-conditional SAGAN train step (1000 classes, attention at 32x32 and 64x64), per-GPU batch "
                "64, 128x128x3, synthetic"),
}
# the reference's legacy program (/root/reference/main.py + models/): 128x128 class-conditional RESIDUAL SAGAN, attention at 32x32
RES128 = dict(COND128, model="resnet", attn_dim_G=[32])
CONFIGS["res128"] = (RES128, "sagan_train_images_per_sec_128x128_resnet",
                     "128x128 class-conditional residual SAGAN train step (models/generator.py:23, models/discriminator.py:40; "
                     "1000 classes, attention at 32x32), per-GPU batch 64, synthetic")
UNIT = "images/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU oracle leg
def cpu_oracle_images_per_sec(cfg, steps, warmup, threads=None):
    """Times the CPU oracle (oracle.train.OracleTrainer: un-fused torch-CPU fp32 restatement of the TF graph) on the
    SAME workload as the GPU arm: full train steps at the full per-GPU batch; only the NUMBER of steps is bounded."""
    from oracle import train as otrain
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads or cores)
    B, S = cfg["batch_size"], cfg["img_size"]
    tr = otrain.OracleTrainer(cfg, torch.float32, seed=0, attn_sigma=0.0, global_batch_size=B)
    rng = np.random.Generator(np.random.PCG64(1234))
    img = torch.tensor(rng.uniform(-1, 1, (B, S, S, 3)).astype(np.float32))
    lab = torch.tensor(rng.integers(0, cfg["num_classes"], B)) if cfg.get("use_label") else None
    times = []
    for s in range(warmup + steps):
        nd = torch.tensor(rng.standard_normal((B, cfg["z_dim"])).astype(np.float32))
        ng = torch.tensor(rng.standard_normal((B, cfg["z_dim"])).astype(np.float32))
        fl = torch.tensor(rng.integers(0, cfg["num_classes"], B)) if cfg.get("use_label") else None
        t0 = time.perf_counter()
        tr.train_step(img, [nd], ng, lab, None if fl is None else [fl], fl)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return dict(value=B / dt, unit=UNIT, cores=torch.get_num_threads(), kind="port",
                sample=f"{steps} timed + {warmup} warm-up full train steps at the full batch of {B} images (fp32), "
                       f"oracle.train.OracleTrainer on torch-CPU"), dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg, metric, workload = CONFIGS[args.config]
    if args.config == "res128":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU oracle trainer restates the vanilla train step only"}))
        return
    # one full-batch CPU step takes seconds: the step COUNT is bounded so the run ends within minutes, the batch is not
    steps = max(1, min(args.steps, 8 if args.config == "church64" else 2))
    base, dt = cpu_oracle_images_per_sec(cfg, steps, 1)
    out = {"impl": "reference", "metric": metric, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload, "global_batch": cfg["batch_size"],
                      "note": f"CPU oracle (torch-CPU port of the reference TF2 graph; TensorFlow is not installable in "
                              f"this image); same step and batch as the GPU arm, {steps} timed steps"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- kernel rooflines
def time_cuda(fn, iters, flush=None):
    """Average device time of fn() over `iters` launches (CUDA events on the launching stream)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)         # 512 MB write: evicts the 126 MB L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e-3


def attn_flops(B, N, C, bwd=False):
    d, dv = C // 8, C // 2
    fwd = 2 * B * N * C * (2 * d + dv) + 2 * B * N * N * d + 2 * B * N * N * dv + 2 * B * N * dv * C
    if not bwd:
        return fwd
    return 2 * B * N * N * (3 * d + 2 * dv) + 2 * (2 * B * N * C * (2 * d + dv) + 2 * B * N * dv * C)


TRAFFIC_FILES = ("r2_traffic.json", "r1_traffic.json")


def _traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this kernel at this
    shape (profiles/r*_traffic.json: NOT measured in this run -- a number taken under a profiler), or None."""
    for f in TRAFFIC_FILES:
        p = os.path.join(ROOT, "profiles", f)
        if os.path.exists(p):
            v = json.load(open(p)).get(name)
            if v is not None:
                return v
    return None


def _traffic_source():
    for f in TRAFFIC_FILES:
        if os.path.exists(os.path.join(ROOT, "profiles", f)):
            return f"profiles/{f} (ncu dram__bytes_read.sum + dram__bytes_write.sum of the same kernel and shape; not measured in this run)"
    return None


def kernel_rooflines(pk, math_mode):
    """The dominant kernels, each timed alone with CUDA events (L2 flushed between iterations):
      * self-attention at G's 64x64 map (B=64, N=4096, C=16), forward and backward  -- the step's top kernels
      * self-attention forward in the sweep regime (B=16, N=4096, C=512)           -- where the block is tensor-bound
      * spectral norm: 4096x16384 (256 MB, 16 B/element rule), 4096x4096 (64 MB, 8 B/element rule), and the 13
        matrices of the church64 generator in one launch."""
    import sagan_b200.functional as F
    from sagan_b200 import MATH_BF16_TC
    flush = torch.zeros(128 * 1024 * 1024, device="cuda")
    mufu_peak = 148 * 16 * 1.965e9       # ex2 / s: 16 per clock per SM at the maximum SM clock
    out = {}

    def attn_case(B, N, C, mode, bwd, iters, pool_grid=None):
        d, dv = C // 8, C // 2
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, N, C, device="cuda", generator=g, requires_grad=True)
        mk = lambda *sh: (torch.randn(*sh, device="cuda", generator=g) / np.sqrt(sh[0])).requires_grad_(True)
        w = [mk(C, d), mk(d), mk(C, d), mk(d), mk(C, dv), mk(dv), mk(dv, C), mk(C),
             torch.tensor(0.5, device="cuda", requires_grad=True)]
        with torch.no_grad():      # un-scaled logits (layers.py:108) of a few units (std 1.5) whatever d is
            w[0] *= 1.5 ** 0.5 / d ** 0.25
            w[2] *= 1.5 ** 0.5 / d ** 0.25
        dy = torch.randn(B, N, C, device="cuda", generator=g)
        with torch.no_grad():
            t_f = time_cuda(lambda: F.attention(x, *w, mode, pool_grid), iters, flush)
        t_b = None
        if bwd:
            y = F.attention(x, *w, mode, pool_grid)
            t_b = time_cuda(lambda: torch.autograd.grad(y, [x] + w, dy, retain_graph=True), max(3, iters // 2), flush)
        return t_f, t_b

    def attn_entry(name, B, N, C, t, bwd, Nk=None):
        Nk = Nk or N
        d, dv = C // 8, C // 2
        fl = attn_flops(B, N, C, bwd)
        if Nk != N:      # down-sampled keys: the score-shaped terms scale with N * Nk, the projections stay
            proj = 2 * B * N * C * (2 * d + dv) + 2 * B * N * dv * C
            fl = (2 * B * N * Nk * (3 * d + 2 * dv) + 2 * proj) if bwd else (proj + 2 * B * N * Nk * (d + dv))
        return dict(bound="tensor", achieved=fl / t / 1e12, peak=pk["tc_burst"], unit="TFLOP/s",
                    frac=fl / t / 1e12 / pk["tc_burst"], traffic=_traffic(name), seconds=t,
                    shape=f"B={B} N={N} keys={Nk} C={C} (d={d}, dv={dv})", exps_per_s=B * N * Nk / t,
                    mufu_frac=B * N * Nk / t / mufu_peak)

    B, N, C = 64, 4096, 16
    t_f, t_b = attn_case(B, N, C, math_mode, True, 10)
    out["attn_fwd"] = attn_entry("attn_fwd", B, N, C, t_f, False)
    out["attn_bwd"] = attn_entry("attn_bwd", B, N, C, t_b, True)
    # the same block with down-sampled keys / values (2x2 / stride-2 max-pooled phi and g: 1024 keys), SURVEY.md §8f-2
    t_f, t_b = attn_case(B, N, C, math_mode, True, 10, pool_grid=(64, 64))
    out["attn_fwd_pooled"] = attn_entry("attn_fwd_pooled", B, N, C, t_f, False, Nk=N // 4)
    out["attn_bwd_pooled"] = attn_entry("attn_bwd_pooled", B, N, C, t_b, True, Nk=N // 4)
    if math_mode == MATH_BF16_TC:
        B, N, C = 16, 4096, 512
        t_f, t_b = attn_case(B, N, C, math_mode, True, 10)
        out["attn_fwd_C512"] = attn_entry("attn_fwd_C512", B, N, C, t_f, False)
        out["attn_fwd_C512"]["note"] = ("sweep regime (BASELINE.json configs[4]): projection GEMM + flash forward + "
                                        "output-conv GEMM, whole block; forward only")
        out["attn_bwd_C512"] = attn_entry("attn_bwd_C512", B, N, C, t_b, True)
        out["attn_bwd_C512"]["note"] = ("whole backward of the block: projection / dA GEMMs, two fused flash launches "
                                        "(dK, dV | dQ; no [N, N] tensor in HBM), dX residual GEMM, weight-gradient GEMMs")

    def sn_entry(name, shapes, rule):
        Ws = [torch.randn(K, R, device="cuda") * 0.02 for R, K in shapes]
        us = [torch.randn(1, R, device="cuda") for R, K in shapes]
        grp = F.SpectralNormGroup(Ws, [u / u.norm() for u in us], 1)
        t = time_cuda(grp.run, 10, flush)
        out[name] = dict(bound="hbm", achieved=grp.algorithmic_bytes / t / 1e9, peak=pk["hbm"], unit="GB/s",
                         frac=grp.algorithmic_bytes / t / 1e9 / pk["hbm"], traffic=_traffic(name), seconds=t, rule=rule)
        del grp, Ws, us

    sn_entry("sn_4096x16384", [(4096, 16384)], "16 B/element (256 MB matrix: three reads of W + one write of W_bar)")
    sn_entry("sn_4096x4096", [(4096, 4096)], "8 B/element (64 MB matrix: read W once + write W_bar, passes 2-3 from L2)")
    sn_entry("sn_church64_G", [(4096, 128), (256, 2048), (128, 1024), (64, 512), (32, 256), (4, 32), (4, 32), (16, 32),
                               (32, 16), (2, 16), (2, 16), (8, 16), (16, 8)],
             "8 B/element; the 13 spectrally-normalised kernels of the generator in ONE launch (4.9 MB: latency-bound)")
    return out


# ----------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, world, local_rank):
    from sagan_b200 import MATH_BF16_TC, MATH_FP32_STRICT, _lib
    from sagan_b200 import nn as snn
    from sagan_b200.trainer import Trainer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    _lib.load()
    math_mode = MATH_BF16_TC if args.math == "bf16_tc" else MATH_FP32_STRICT
    snn.set_default_math_mode(math_mode)
    cfg, metric, workload = CONFIGS[args.config]
    cfg = dict(cfg, batch_size=args.batch or cfg["batch_size"], attn_downsample=args.attn_downsample)
    B, S = cfg["batch_size"], cfg["img_size"]
    n_records = 126227 if args.config == "church64" else 1281167           # LSUN church / ImageNet train set sizes
    tr = Trainer(cfg, global_batch_size=B * world, steps_per_epoch=n_records // (B * world), seed=0,
                 dp_mode=args.dp, overlap_streams=args.overlap)
    rng = np.random.Generator(np.random.PCG64(1234 + rank))
    # the reference's input contract (sagan/dataset.py:27-40): raw uint8 HWC records, decoded as x * (2. / 255) - 1. --
    # here on the device, as the first node of the step graph
    host_batches = [torch.tensor(rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)).pin_memory() for _ in range(4)]
    dev_batches = [b.to(dev) for b in host_batches]
    use_label = bool(cfg.get("use_label"))
    host_labels = [torch.tensor(rng.integers(0, cfg["num_classes"], B)).pin_memory() if use_label else None for _ in range(4)]
    dev_labels = [None if l is None else l.to(dev) for l in host_labels]

    n0 = _lib.launch_count()
    tr.capture(warmup=max(3, args.warmup), uint8_input=True)
    # kernels launched while capturing == kernel nodes of this library in one replay of the step graph
    n_cap0 = _lib.launch_count()
    launches_per_step = None
    # count by re-running one eager step (same launch sequence as the captured one)
    c0 = _lib.launch_count()
    tr.train_step(dev_batches[0], dev_labels[0])
    launches_per_step = _lib.launch_count() - c0
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) * 1e-3

    # ---- value: inputs resident in HBM, whole-step CUDA graph
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # sampled from the warm-up on: the GPU is under the same load throughout
    for s in range(args.warmup):
        tr.graph_step(dev_batches[s % 4], dev_labels[s % 4])
    sec = timed(lambda s: tr.graph_step(dev_batches[s % 4], dev_labels[s % 4]), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * args.steps / sec

    # ---- e2e: pinned host batch -> device each step, losses read back each step
    def e2e_step(s):
        tr.graph_step(host_batches[s % 4], host_labels[s % 4])
        tr.losses()                       # global loss sums (all-reduced at N > 1) read back to the host (synchronises)
    for s in range(2):
        e2e_step(s)
    sec_e2e = timed(e2e_step, args.steps)
    e2e = B * world * args.steps / sec_e2e
    losses = tr.losses()
    dp_mode = tr.dp_mode
    # replica consistency after all the steps above: every replica must hold bit-identical weights (the fused exchange
    # sums in a fixed order and writes the same values everywhere); a replica that timed out at a barrier raises
    dp_status, replica_diff = "single replica", 0.0
    if world > 1:
        try:
            tr.check_exchange()
            dp_status = "ok"
        except Exception as e:      # noqa: BLE001
            dp_status = f"FAILED: {e}"
        worst = torch.zeros(1, device=dev)
        for flat in (tr.G.flat_params, tr.D.flat_params):
            ref = flat.clone()
            torch.distributed.broadcast(ref, 0)
            worst = torch.maximum(worst, (flat - ref).abs().max().reshape(1))
        torch.distributed.all_reduce(worst, op=torch.distributed.ReduceOp.MAX)
        replica_diff = float(worst)

    # ---- the same step with the down-sampled attention layers.py:96,100,113 reaches for (keys / values max-pooled 2x2 /
    #      stride 2): NOT the headline workload (whose attention is the oracle's un-pooled reading) -- reported beside it
    variant = None
    if world == 1 and not args.attn_downsample and cfg.get("use_attention") and args.variant:
        cfg_p = dict(cfg, attn_downsample=True)
        tr_p = Trainer(cfg_p, global_batch_size=B, steps_per_epoch=n_records // B, seed=0, overlap_streams=args.overlap)
        tr_p.capture(warmup=3, uint8_input=True)
        for s in range(3):
            tr_p.graph_step(dev_batches[s % 4], dev_labels[s % 4])
        sec_p = timed(lambda s: tr_p.graph_step(dev_batches[s % 4], dev_labels[s % 4]), args.steps)
        variant = {"config": "attn_downsample=True (SURVEY.md 8f-2)", "value": B * args.steps / sec_p, "unit": UNIT,
                   "ms_per_step": sec_p / args.steps * 1e3, "final_losses": tr_p.losses()}
        tr_p.graph = None
        del tr_p

    # every collective is behind us: tear the communicator down on ALL ranks together (a rank that exits while
    # another still holds captured NCCL work can hang in the teardown), then rank 0 alone finishes the report
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        tr.graph = None
        del tr
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    pk = peaks()
    roofs = kernel_rooflines(pk, math_mode)
    dom = max(("attn_fwd", "attn_bwd"), key=lambda k: roofs[k]["seconds"])
    roofline = dict(roofs[dom], kernel=dom, peak_source=pk["src"],
                    note="dominant kernel of the step.  BASELINE.json names the tensor pipe as this block's roof, but at "
                         "the in-model dims (d=2, dv=8: 20 useful FLOPs per score against one exp2) the MUFU pipe is the "
                         "binding unit: mufu_frac = exps_per_s / (148 SMs x 16 ex2/clk x 1.965 GHz).  The tensor-bound "
                         "regime is kernels.attn_fwd_C512.")
    if args.cpu_baseline and args.config != "res128":      # the CPU oracle trainer restates the vanilla step only
        cpu, _ = cpu_oracle_images_per_sec(dict(CONFIGS[args.config][0]), 4 if args.config == "church64" else 1, 1)
    else:
        cpu = None
    act_mb = 4 * B * (S * S * 16 * 12 + (S // 2) ** 2 * 32 * 10) / 1e6
    roofline["traffic_source"] = _traffic_source()
    out = {
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if math_mode == MATH_FP32_STRICT else "bf16", "data": "synthetic (uniform uint8 records)",
        "config": {"workload": (workload if B == CONFIGS[args.config][0]["batch_size"] else workload + f" [per-GPU batch {B}]")
                               + (" [attn_downsample: keys / values max-pooled 2x2 / stride 2]" if args.attn_downsample else ""),
                   "global_batch": B * world, "parallelism": f"dp{world}",
                   "math_mode": args.math, "conv_precision": "split-bf16 (3 MMAs per K step, fp32-grade)", "cuda_graph": True,
                   "step_graph": ("two branches: generator forwards on a side stream" if args.overlap
                                  else "single stream (--no-overlap)"),
                   "dp_exchange": {"p2p": "fused NVLink peer-memory gradient sum + Adam kernel (csrc/dp.cu)",
                                   "nccl": "NCCL all-reduce + Adam", "none": "single replica"}[dp_mode],
                   "l2": f"no explicit flush: a step touches ~{act_mb:.0f} MB of saved activations (> 126 MB L2)"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": sec_e2e / args.steps * 1e3,
                "h2d_bytes_per_step": int(host_batches[0].numel() + (B * 8 if use_label else 0)),
                "input": "raw uint8 records (sagan/dataset.py:27-40), decoded on the device inside the step graph",
                "d2h_bytes_per_step": 8},
        "dp_status": dp_status, "replica_max_abs_diff": replica_diff,
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "roofline": roofline,
        "kernels": {k: {kk: vv for kk, vv in v.items()} for k, v in roofs.items()},
        "cpu_baseline": cpu,
        "attn_downsample_variant": variant,
        "final_losses": losses,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="church64", choices=sorted(CONFIGS),
                    help="church64 = BASELINE.json configs[1] (headline); cond128 = configs[3], 128x128 class-conditional; "
                         "res128 = the legacy residual 128x128 program")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's 64)")
    ap.add_argument("--attn-downsample", action="store_true",
                    help="down-sampled attention (keys / values max-pooled 2x2 / stride 2, layers.py:100,113); not the "
                         "headline workload, whose attention is the oracle's un-pooled reading")
    ap.add_argument("--math", default="bf16_tc", choices=["fp32_strict", "bf16_tc"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-variant", dest="variant", action="store_false",
                    help="skip the extra timing of the step with down-sampled attention (N = 1 only)")
    ap.add_argument("--dp", default="p2p", choices=["p2p", "nccl"], help="data-parallel gradient exchange (N > 1)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="single-stream step (default: G(z) and D(real) of the D phase run as two graph branches)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: spawn one process per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    args.warmup = max(args.warmup, 3)
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner at N > 1) are sent to
    # stderr while the run is in progress; the report itself goes to the saved descriptor
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(saved_stdout, "w", buffering=1)
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the CPU baseline leg is an N = 1 measurement
        args.cpu_baseline = False
    run_ours(args, rank, world, local_rank)
    if world > 1:
        sys.stdout.flush()
        os._exit(0)      # skip interpreter teardown of NCCL / CUDA-graph state (it has hung at exit on some boxes)


if __name__ == "__main__":
    main()
