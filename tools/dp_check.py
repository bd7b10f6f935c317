#!/usr/bin/env python
"""Multi-GPU check of the fused peer-memory gradient exchange (run under torchrun, N >= 2 GPUs of one box):
the weights after a few training steps with dp_mode="p2p" (csrc/dp.cu) must match dp_mode="nccl" (all-reduce + Adam)
and be identical on every replica; prints the per-update time of both."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from sagan_b200.trainer import Trainer
    cfg = dict(z_dim=128, gf_dim=16, df_dim=16, img_size=64, use_attention=True, attn_dim_G=[32, 64], attn_dim_D=[8, 4],
               use_label=False, batch_size=8, lr_g=2e-4, lr_d=7e-4, decay_rate=0.99, update_ratio=1)
    out = {}
    weights = {}
    for mode in ("nccl", "p2p"):
        tr = Trainer(cfg, seed=0, dp_mode=mode)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for s in range(3):
            img = torch.rand(8, 64, 64, 3, device=dev, generator=g) * 2 - 1
            nd = torch.randn(8, 128, device=dev, generator=g)
            ng = torch.randn(8, 128, device=dev, generator=g)
            tr.train_step(img, None, [nd], ng)
        torch.cuda.synchronize()
        tr.check_exchange()
        # the reported losses are the GLOBAL sums / global batch (main.py:216-220): identical on every replica
        ls = tr.losses()
        lt = torch.tensor([ls["D_loss"], ls["G_loss"]], device=dev, dtype=torch.float64)
        l0 = lt.clone(); dist.broadcast(l0, 0)
        out[f"{mode}_loss_replica_max_abs_diff"] = float((lt - l0).abs().max())
        weights[mode] = (tr.G.flat_params.clone(), tr.D.flat_params.clone())
        # identical on every replica?
        for name, w in zip("GD", weights[mode]):
            ref = w.clone()
            dist.broadcast(ref, 0)
            out[f"{mode}_{name}_replica_max_abs_diff"] = float((w - ref).abs().max())
        # time one exchange + update in isolation
        net, opt = tr.G, tr.opt_G
        def upd():
            if tr.peer_G is not None:
                opt.schedule(); tr.peer_G.step()
            else:
                tr._allreduce(net.flat_grads); opt.apply()
        for _ in range(5):
            upd()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            upd()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{mode}_G_update_us"] = float(t) * 1e3
        if tr.peer_G is not None:
            tr.peer_G.check()
        del tr
    for i, name in enumerate("GD"):
        a, b = weights["nccl"][i], weights["p2p"][i]
        out[f"p2p_vs_nccl_{name}_rel_l2_after_3_steps"] = float((a - b).norm() / a.norm())
    # exactness of the exchange itself: identical parameters, gradients and Adam state through both paths, ONE update
    res = {}
    for mode in ("nccl", "p2p"):
        tr = Trainer(cfg, seed=0, dp_mode=mode)
        gg = torch.Generator(device=dev).manual_seed(7 + rank)
        tr.G.flat_grads.copy_(torch.randn(tr.G.flat_grads.numel(), device=dev, generator=gg) * 1e-3)
        p0 = tr.G.flat_params.clone()
        if tr.peer_G is not None:
            tr.opt_G.schedule(); tr.peer_G.step()
        else:
            tr._allreduce(tr.G.flat_grads); tr.opt_G.apply()
        torch.cuda.synchronize()
        res[mode] = (tr.G.flat_params - p0).clone()
        del tr
    # every replica draws its own noise (main.py:176,194 run per replica): the device generators must differ
    tr = Trainer(cfg, seed=0, dp_mode="nccl")
    z = torch.randn(8, 128, device=dev)
    z0 = z.clone(); dist.broadcast(z0, 0)
    differs = torch.tensor([float((z - z0).abs().max() > 0)], device=dev)
    dist.all_reduce(differs)
    out["replicas_with_own_noise"] = int(differs.item()) + 1        # rank 0 trivially equals itself
    del tr
    out["one_update_delta_rel_l2"] = float((res["nccl"] - res["p2p"]).norm() / res["nccl"].norm())
    out["one_update_delta_max_abs"] = float((res["nccl"] - res["p2p"]).abs().max())
    if rank == 0:
        out["world"] = world
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
