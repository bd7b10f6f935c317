"""N > 1 host logic on CPU: world_size-2 gloo run of the data-parallel gradient exchange.

Each rank differentiates the oracle discriminator on ITS shard of a global batch with the reference's loss scaling
(mean over elements / global batch, sagan/main.py:184), the flat gradient buckets are SUM-all-reduced by
sagan_b200.parallel.ReplicaGradientSum, and the result must equal the single-process gradient of the whole batch
(exact for D, which has no BatchNorm: SURVEY.md §8e)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "self-attention-gan_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

CFG = dict(z_dim=16, gf_dim=8, df_dim=8, img_size=16, use_attention=True, attn_dim_G=[8], attn_dim_D=[8],
           use_label=False, batch_size=2, lr_g=2e-4, lr_d=7e-4, decay_rate=0.99, update_ratio=1)


def _d_grads(images, fake, global_batch):
    from oracle import nets as onets
    from oracle import train as otrain
    spec = onets.discriminator_spec(CFG)
    D = onets.init_params(spec, 100, torch.float64, 0.3, 0.05)
    sn = onets.init_sn_state(spec, 101, torch.float64)
    for v in D.values():
        v.requires_grad_(True)
    d_real = onets.discriminator_forward(D, dict(sn), images, CFG, None, True)
    d_fake = onets.discriminator_forward(D, dict(sn), fake, CFG, None, True)
    le = otrain.hinge_loss_d(d_real, d_fake)
    scalar = le.mean() * (1.0 / global_batch)                     # sagan/main.py:184
    names = list(D.keys())
    gs = torch.autograd.grad(scalar, [D[k] for k in names], allow_unused=True)
    flat = torch.cat([(g if g is not None else torch.zeros_like(D[k])).reshape(-1) for k, g in zip(names, gs)])
    return flat, float(le.sum())


def _data(global_batch):
    rng = np.random.Generator(np.random.PCG64(11))
    img = torch.tensor(rng.uniform(-1, 1, (global_batch, 16, 16, 3)))
    fake = torch.tensor(rng.uniform(-1, 1, (global_batch, 16, 16, 3)))
    return img, fake


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sagan_b200.parallel import ReplicaGradientSum, shard_range
    dp = ReplicaGradientSum()
    gb = dp.global_batch(CFG["batch_size"])
    img, fake = _data(gb)
    a, b = shard_range(gb, dp.rank, dp.world)
    flat, loss_sum = _d_grads(img[a:b], fake[a:b], gb)
    state = torch.full((4,), float(rank))
    dp.broadcast_(state)
    dp.sum_(flat)
    losses = dp.sum_losses_(torch.tensor([loss_sum], dtype=torch.float64))
    if rank == 0:
        torch.save(dict(flat=flat, loss=losses, state=state, gb=gb, world=dp.world), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_replicas_sum_equals_single_replica_big_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["world"] == 2 and res["gb"] == 4
    assert torch.equal(res["state"], torch.zeros(4))               # broadcast from rank 0
    img, fake = _data(4)
    ref, ref_loss = _d_grads(img, fake, 4)
    # The reference scales by BOTH the per-replica element mean and the global batch (sagan/main.py:184), so the summed
    # replica gradients equal the single-replica big-batch gradient times the replica count (a quirk kept for parity:
    # SURVEY.md Appendix A.8; Adam is invariant to it).
    err = float((res["flat"] - 2.0 * ref).norm() / (2.0 * ref).norm())
    assert err < 1e-12, err
    assert abs(float(res["loss"][0]) - ref_loss) < 1e-9


def test_shard_range_drops_remainder():
    from sagan_b200.parallel import shard_range
    assert shard_range(8, 0, 2) == (0, 4) and shard_range(8, 1, 2) == (4, 8)
    assert shard_range(9, 1, 2) == (4, 8)                           # sagan/dataset.py:39 drop_remainder
    assert shard_range(64, 7, 8) == (56, 64)
