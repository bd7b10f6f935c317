"""Generates the golden fixtures in this directory from the CPU oracle (fp64 unless noted).

The reference ships no golden vectors for this path (its tests assert shapes only), and TensorFlow is
not installable here, so these vectors are the oracle's own outputs frozen at commit time: they pin
the oracle against regressions and let the GPU tests compare against fp64 results without re-running
the fp64 oracle at size.  Inputs are regenerated from numpy PCG64 seeds (bit-reproducible across
platforms) by the helpers below, which the tests import too.  The oracle itself is held to the reference's own
code by the OTHER fixture, reference_layers.npz (make_reference_vectors.py).

    python tests/golden/make_golden.py            # everything except the 100-step trajectory
    python tests/golden/make_golden.py --traj     # also the 100-step loss trajectory (~minutes of CPU)
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import attention, nets, sn, train  # noqa: E402

# every spectrally-normalised matrix shape of church64_attn G and D + two odd ones
SN_SHAPES = [(4096, 128), (256, 2048), (128, 1024), (64, 512), (32, 256), (4, 32), (16, 32), (32, 16),
             (2, 16), (8, 16), (16, 8), (16, 48), (7, 13), (1, 130)]
ATTN_CASES = [(2, 64, 16), (2, 256, 16), (2, 64, 32), (2, 256, 32), (2, 64, 64), (2, 200, 16)]
# in-model token counts (N in {64, 256, 1024}): the 32x32 maps of church64 / 128x128-conditional G and D
ATTN_CASES_N1024 = [(1, 1024, 16), (1, 1024, 32), (1, 1024, 64)]

TEST_CFG = dict(z_dim=128, gf_dim=16, df_dim=16, img_size=64, use_attention=True, attn_dim_G=[32, 64],
                attn_dim_D=[8, 4], use_label=False, batch_size=4, lr_g=2e-4, lr_d=7e-4, decay_rate=0.99,
                update_ratio=1, loss="hinge_loss", model="vanilla")


def sn_inputs(R, K, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    # a Dense-style kernel [K, R]: its LAST axis is R, and layers.py:56 raw-reshapes it to [R, K]
    W = (rng.standard_normal((K, R)) * 0.02).astype(np.float32)
    u = rng.standard_normal((1, R)).astype(np.float32)
    u = (u / (np.linalg.norm(u) + 1e-12)).astype(np.float32)
    dW = rng.standard_normal((K, R)).astype(np.float32)
    return W, u, dW


def attn_inputs(B, N, C, seed):
    X, dY, w = attention.make_inputs(B, N, C, seed=seed, gamma=0.37, dtype=np.float32)
    return X, dY, w


def step_inputs(cfg, step, seed=1234):
    """Synthetic batch of SURVEY.md §8d: uniform[-1,1) images, N(0,1) noise, PCG64(seed + step)."""
    rng = np.random.Generator(np.random.PCG64(seed + step))
    B, S = cfg["batch_size"], cfg["img_size"]
    img = rng.uniform(-1.0, 1.0, (B, S, S, 3)).astype(np.float32)
    nd = rng.standard_normal((B, cfg["z_dim"])).astype(np.float32)
    ng = rng.standard_normal((B, cfg["z_dim"])).astype(np.float32)
    return img, nd, ng


def summarize(a):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    return np.concatenate([[np.sqrt(np.sum(a * a)), a.sum()], a[:16] if a.size >= 16 else np.pad(a, (0, 16 - a.size))])


def make_sn():
    out = {}
    for i, (R, K) in enumerate(SN_SHAPES):
        W, u, dW = sn_inputs(R, K, 100 + i)
        for Ip, factor in ((1, None), (2, 1.5)):
            u2, v2, sig, Wb = sn.power_iteration(W.astype(np.float64), u.astype(np.float64), Ip, factor)
            g = sn.backward(dW.astype(np.float64), Wb, u2, v2, sig, factor)
            tag = f"{R}x{K}_Ip{Ip}"
            out[tag + "_u"] = u2
            out[tag + "_v"] = v2
            out[tag + "_sigma"] = np.float64(sig)
            out[tag + "_Wbar_sum"] = summarize(Wb)
            out[tag + "_dW_sum"] = summarize(g)
    np.savez_compressed(os.path.join(HERE, "sn.npz"), **out)


def make_attn(cases=ATTN_CASES, seed0=200, fname="attention.npz"):
    out = {}
    for i, (B, N, C) in enumerate(cases):
        X, dY, w = attn_inputs(B, N, C, seed0 + i)
        w64 = {k: np.asarray(v, dtype=np.float64) for k, v in w.items()}
        Y = attention.forward(X.astype(np.float64), **w64)
        g = attention.backward(dY.astype(np.float64), X.astype(np.float64), **w64)
        tag = f"B{B}_N{N}_C{C}"
        out[tag + "_Y"] = Y.astype(np.float32)
        out[tag + "_dX"] = g["dX"].astype(np.float32)
        for k in attention.WEIGHT_NAMES:
            out[tag + "_d" + k] = np.asarray(g["d" + k], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, fname), **out)


def make_nets():
    """G / D forward + gradients at the example_configs/test.py model, B = 4, fp64, sigma = 0.37."""
    cfg = TEST_CFG
    tr = train.OracleTrainer(cfg, torch.float64, seed=0, attn_sigma=0.37, bias_scale=0.05, global_batch_size=4)
    img, nd, ng = step_inputs(cfg, 0)
    img, nd, ng = (torch.tensor(a, dtype=torch.float64) for a in (img, nd, ng))
    out = {}
    dgr, dl = tr.d_grads(img, nd)
    out["D_loss_elems"] = dl.numpy()
    for k, v in dgr.items():
        out["Dgrad." + k] = summarize(v.numpy())
    ggr, gl = tr.g_grads(ng)
    out["G_loss_elems"] = gl.numpy()
    for k, v in ggr.items():
        out["Ggrad." + k] = summarize(v.numpy())
    np.savez_compressed(os.path.join(HERE, "nets.npz"), **out)


def make_traj(steps=100):
    cfg = TEST_CFG
    tr = train.OracleTrainer(cfg, torch.float32, seed=0, attn_sigma=0.0, bias_scale=0.0, global_batch_size=4,
                             steps_per_epoch=40)
    G, D = [], []
    for s in range(steps):
        img, nd, ng = (torch.tensor(a) for a in step_inputs(cfg, s))
        r = tr.train_step(img, [nd], ng)
        G.append(r["G_loss"])
        D.append(r["D_loss"])
        if s % 10 == 0:
            print(s, r, flush=True)
    np.savez_compressed(os.path.join(HERE, "trajectory.npz"), G_loss=np.array(G), D_loss=np.array(D))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj", action="store_true")
    ap.add_argument("--only-traj", action="store_true")
    ap.add_argument("--only-n1024", action="store_true", help="only attention_n1024.npz (added in round 2)")
    a = ap.parse_args()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if a.only_n1024:
        make_attn(ATTN_CASES_N1024, 300, "attention_n1024.npz")
    elif not a.only_traj:
        make_sn()
        make_attn()
        make_attn(ATTN_CASES_N1024, 300, "attention_n1024.npz")
        make_nets()
    if a.traj or a.only_traj:
        make_traj()
    print("golden fixtures written to", HERE)
