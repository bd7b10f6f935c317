"""One replica of the SAGAN training step, B200-native.

Mirrors Trainer.train_step / distributed_train_step of /root/reference/sagan/main.py:171-236:
  D phase x update_ratio:  G forward (no tape) -> D(real), D(fake) -> hinge -> grads -> all-reduce -> Adam
  G phase:                 G forward -> D forward -> -D(G(z)) -> grads wrt G -> all-reduce -> Adam
with the loss scaling of main.py:184,201 (mean over elements / global batch, replica gradients SUMMED).

One process per GPU.  The replica gradient sum that MirroredStrategy performs inside
`apply_gradients` (main.py:190,205) is an NCCL all-reduce over each network's single flat gradient
bucket, issued through torch.distributed (plumbing); Adam then runs over the flat bucket.  The whole
step can be captured into one CUDA graph (`capture()`), which removes the launch latency of the
~250 small kernels a step consists of.
"""
import math

import torch

from . import functional as F
from . import nets
from ._lib import MATH_FP32_STRICT  # noqa: F401
from .parallel import PeerAdam, ReplicaGradientSum, symmetric_allocator

ADAM_B1, ADAM_B2, ADAM_EPS = 0.0, 0.999, 1e-7      # main.py:119-120 (Keras defaults, beta_1 = 0)


class FlatAdam:
    """Keras Adam + ExponentialDecay(staircase) over a network's flat bucket (main.py:111-120)."""

    def __init__(self, net, lr0, decay_steps, decay_rate):
        self.net = net
        self.lr0, self.decay_steps, self.decay_rate = float(lr0), int(decay_steps), float(decay_rate)
        self.iterations = 0
        self.v = torch.zeros_like(net.flat_params)
        self.hyper = torch.zeros(4, device=net.flat_params.device)
        self._host = torch.zeros(4).pin_memory()

    def lr_t(self):
        lr = self.lr0 * self.decay_rate ** (self.iterations // self.decay_steps)
        t = self.iterations + 1
        return lr * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)

    def stage_hyper(self):
        """Host -> device copy of this step's [lr_t, b1, b2, eps]; stays outside a captured graph."""
        self._host[0], self._host[1], self._host[2], self._host[3] = self.lr_t(), ADAM_B1, ADAM_B2, ADAM_EPS
        self.hyper.copy_(self._host, non_blocking=True)

    def apply(self):
        F.adam_step(self.net.flat_params, self.net.flat_grads, self.v, self.hyper)


class Trainer:
    def __init__(self, config, global_batch_size=None, steps_per_epoch=1000, process_group=None, seed=0, dp_mode="p2p",
                 overlap_streams=True):
        """config: the reference's dict (example_configs/*.py).  process_group: torch.distributed group for the
        data-parallel gradient sum (None = single replica / default group).  dp_mode: "p2p" = the fused NVLink
        peer-memory exchange + Adam kernel (parallel.PeerAdam), "nccl" = NCCL all-reduce followed by Adam."""
        self.config = dict(config)
        cfg = self.config
        self.B = cfg["batch_size"]
        self.dp = ReplicaGradientSum(process_group)
        self.world = self.dp.world
        # main.py:358: global_batch_size = batch_size * len(gpu)
        self.global_batch = global_batch_size or self.dp.global_batch(self.B)
        self.device = torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(seed)
        self.dp_mode = dp_mode if self.world > 1 else "none"
        if self.dp_mode == "p2p":
            nets.set_flat_allocator(symmetric_allocator())
        self.G = nets.get_generator(cfg)
        self.D = nets.get_discriminator(cfg)
        with torch.no_grad():   # build pass (Keras `model.build`, main.py:134-135)
            z = torch.zeros(self.B, cfg["z_dim"], device=self.device)
            lab = torch.zeros(self.B, dtype=torch.int64, device=self.device) if cfg.get("use_label") else None
            img = self.G([z, lab])
            self.D([img, lab])
        # identical initial weights and spectral-norm state on every replica (MirroredStrategy variables)
        self.dp.broadcast_(self.G.flat_params, self.D.flat_params, self.G.sn_group.out, self.D.sn_group.out)
        ur = cfg.get("update_ratio", 1)
        self.opt_G = FlatAdam(self.G, cfg["lr_g"], steps_per_epoch, cfg["decay_rate"])          # main.py:111-114
        self.opt_D = FlatAdam(self.D, cfg["lr_d"], steps_per_epoch * ur, cfg["decay_rate"])     # main.py:115-118
        nets.set_flat_allocator(None)
        self.peer_G = self.peer_D = None
        if self.dp_mode == "p2p":
            self.peer_G = PeerAdam(self.G, self.opt_G, self.dp)
            self.peer_D = PeerAdam(self.D, self.opt_D, self.dp)
        self.loss_sums = torch.zeros(2, device=self.device)     # [sum L_D (over update_ratio), sum L_G]
        self.overlap_streams = overlap_streams
        self._side = torch.cuda.Stream(device=self.device) if overlap_streams else None
        if overlap_streams and hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # intentional: see _d_phase
        self.graph = None
        self._static = {}

    # ------------------------------------------------------------------------------------------
    def _allreduce(self, flat):
        self.dp.sum_(flat)

    def _d_phase(self, images, labels, noise, fake_labels):
        G, D = self.G, self.D
        D.zero_grad_flat()
        if self.overlap_streams:
            # G(z) (no tape) and D(real) are independent: two branches of the step graph.  Most of their conv / BN /
            # spectral-norm launches fill a fraction of the 148 SMs, so the branches overlap.  The GENERATOR forwards go
            # to the side stream: every D call stays on the main stream, because the two D calls of this phase
            # accumulate their spectral-norm weight gradients in place into the same flat bucket and must not run
            # their backward nodes concurrently (autograd replays a node on the stream of its forward).
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side), torch.no_grad():                # main.py:178 (outside the tape)
                fake = G([noise, fake_labels], training=True)
            d_real = D([images, labels], training=True)                     # main.py:181
            main.wait_stream(self._side)
            fake.record_stream(main)
        else:
            with torch.no_grad():                                           # main.py:178 (outside the tape)
                fake = G([noise, fake_labels], training=True)
            d_real = D([images, labels], training=True)                     # main.py:181
        d_fake = D([fake, fake_labels], training=True)                      # main.py:182
        g_real, g_fake = F.hinge_d_grads(d_real, d_fake, self.global_batch, self.loss_sums[0:1])   # main.py:183-184
        torch.autograd.backward([d_real, d_fake], [g_real, g_fake])         # main.py:188-189
        if self.peer_D is not None:
            self.peer_D.step()                                              # main.py:190: replica SUM + Adam, one kernel
        else:
            self._allreduce(D.flat_grads)                                   # main.py:190 (replica SUM)
            self.opt_D.apply()

    def _g_phase(self, noise, fake_labels):
        G, D = self.G, self.D
        G.zero_grad_flat()
        for p in D.parameters():
            p.requires_grad_(False)                                         # main.py:203-204: grads wrt G only
        try:
            if self.overlap_streams:
                # G(z') of the G phase depends on nothing the D phase still has to do (D's update only matters for
                # D(G(z')) below): it follows the D phase's own G(z) on the side stream (spectral-norm u and the
                # BatchNorm moving statistics persist from call to call, so the two forwards stay ordered) and runs under
                # D(fake), the D backward and D's Adam / gradient exchange.  Its backward nodes run on the side stream
                # too; the engine orders them after the D backward-data nodes that feed them.
                main = torch.cuda.current_stream()
                with torch.cuda.stream(self._side):
                    fake = G([noise, fake_labels], training=True)           # main.py:198
                main.wait_stream(self._side)
                fake.record_stream(main)
            else:
                fake = G([noise, fake_labels], training=True)               # main.py:198
            d_fake = D([fake, fake_labels], training=True)                  # main.py:199
            g = F.hinge_g_grads(d_fake, self.global_batch, self.loss_sums[1:2])   # main.py:200-201
            d_fake.backward(g)
        finally:
            for p in D.parameters():
                p.requires_grad_(True)
        if self.peer_G is not None:
            self.peer_G.step()                                              # main.py:205
        else:
            self._allreduce(G.flat_grads)                                   # main.py:205 (replica SUM)
            self.opt_G.apply()

    def _step_body(self, images, labels, noises_d, noise_g, fake_labels_d, fake_labels_g):
        cfg = self.config
        ur = cfg.get("update_ratio", 1)
        self.loss_sums.zero_()
        # all noise / label draws first, in the reference's order (main.py:176-177,194-195): the G phase's forward runs
        # as a branch that starts before the D phase is over, so its inputs must exist by then
        draws = []
        for i in range(ur):                                                 # main.py:175
            nz = noises_d[i] if noises_d is not None else torch.randn(self.B, cfg["z_dim"], device=self.device)
            fl = fake_labels_d[i] if fake_labels_d is not None else self._rand_labels()
            draws.append((nz, fl))
        nz_g = noise_g if noise_g is not None else torch.randn(self.B, cfg["z_dim"], device=self.device)
        fl_g = fake_labels_g if fake_labels_g is not None else self._rand_labels()
        for nz, fl in draws:
            self._d_phase(images, labels, nz, fl)
        self._g_phase(nz_g, fl_g)

    def _rand_labels(self):
        cfg = self.config
        if not cfg.get("use_label"):
            return None
        return torch.randint(0, cfg["num_classes"], (self.B,), device=self.device)   # main.py:177,195

    def _stage(self):
        self.opt_D.stage_hyper()
        self.opt_G.stage_hyper()

    def _advance(self):
        self.opt_D.iterations += self.config.get("update_ratio", 1)
        self.opt_G.iterations += 1

    # ------------------------------------------------------------------------------------------
    def train_step(self, images, labels=None, noises_d=None, noise_g=None, fake_labels_d=None, fake_labels_g=None):
        """Eager step.  images: device NHWC float32 in [-1,1] (sagan/dataset.py:34).  Noise may be injected
        (parity tests); otherwise it is drawn on the device (main.py:176,194).  Returns the device tensor
        [sum L_D, sum L_G]; see `losses()` for the reported means."""
        if self.config.get("update_ratio", 1) != 1 and noises_d is None:
            pass
        self._stage()
        self._step_body(images, labels, noises_d, noise_g, fake_labels_d, fake_labels_g)
        self._advance()
        return self.loss_sums

    def losses(self):
        """Reported losses (main.py:216-229): sum over the batch / global batch, mean over the rest.
        Synchronises (device -> host read of two floats)."""
        s = self.loss_sums.tolist()
        ur = self.config.get("update_ratio", 1)
        n_elem = self._logit_elems()
        return dict(D_loss=s[0] / ur / (self.global_batch * n_elem), G_loss=s[1] / (self.global_batch * n_elem))

    def _logit_elems(self):
        return 1 if self.config.get("use_label") else 16     # [B,4,4,1] patch logits (discriminator.py:35)

    # ------------------------------------------------------------------------------------------
    def capture(self, warmup=3):
        """Capture the whole step (both phases, all-reduces, Adam) into one CUDA graph.  Images (and labels)
        are read from static device buffers that `graph_step` refills."""
        cfg = self.config
        self._static["images"] = torch.zeros(self.B, cfg["img_size"], cfg["img_size"], 3, device=self.device)
        self._static["labels"] = (torch.zeros(self.B, dtype=torch.int64, device=self.device)
                                  if cfg.get("use_label") else None)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.train_step(self._static["images"], self._static["labels"])
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self._stage()
        with torch.cuda.graph(self.graph):
            self._step_body(self._static["images"], self._static["labels"], None, None, None, None)
        return self.graph

    def graph_step(self, images=None, labels=None):
        """Replay the captured step.  `images` may be a pinned host tensor (copied asynchronously) or a
        device tensor; None keeps the buffer contents."""
        if self.graph is None:
            raise RuntimeError("call capture() first")
        if images is not None:
            self._static["images"].copy_(images, non_blocking=True)
        if labels is not None and self._static["labels"] is not None:
            self._static["labels"].copy_(labels, non_blocking=True)
        self._stage()
        self.graph.replay()
        self._advance()
        return self.loss_sums
