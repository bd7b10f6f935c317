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
    """Keras Adam + ExponentialDecay(staircase) over a network's flat bucket (main.py:111-120).

    `optimizer.iterations` lives in DEVICE memory and the schedule (staircase decay, bias correction) is evaluated by
    a one-thread kernel right before each application (`sagan_adam_schedule`), which also bumps the counter -- once
    per `apply_gradients`, as Keras does (so the `update_ratio` D updates of one step see t, t+1, ...).  Nothing is
    staged from the host per step: a replayed CUDA graph follows the schedule by itself, and the CPU running many
    replays ahead of the GPU cannot hand a step the learning rate of a later one."""

    def __init__(self, net, lr0, decay_steps, decay_rate):
        self.net = net
        self.lr0, self.decay_steps, self.decay_rate = float(lr0), int(decay_steps), float(decay_rate)
        dev = net.flat_params.device
        self._iters = torch.zeros(1, dtype=torch.int64, device=dev)
        self.v = torch.zeros_like(net.flat_params)
        self.hyper = torch.zeros(4, device=dev)

    @property
    def iterations(self):
        """optimizer.iterations (synchronises: device -> host read)."""
        return int(self._iters.item())

    @iterations.setter
    def iterations(self, value):
        self._iters.fill_(int(value))

    def lr_t(self, iterations=None):
        """Host restatement of the schedule kernel (tests / logging)."""
        it = self.iterations if iterations is None else int(iterations)
        lr = self.lr0 * self.decay_rate ** (it // self.decay_steps)
        t = it + 1
        return lr * math.sqrt(1.0 - ADAM_B2 ** t) / (1.0 - ADAM_B1 ** t)

    def schedule(self):
        """hyper <- [lr_t, b1, b2, eps] for the current `iterations`; iterations += 1 (on the stream)."""
        F.adam_schedule(self.hyper, self._iters, self.lr0, self.decay_rate, self.decay_steps, ADAM_B1, ADAM_B2, ADAM_EPS)

    def apply(self):
        self.schedule()
        F.adam_step(self.net.flat_params, self.net.flat_grads, self.v, self.hyper)


class Trainer:
    def __init__(self, config, global_batch_size=None, steps_per_epoch=1000, process_group=None, seed=0, dp_mode="p2p",
                 overlap_streams=True):
        """config: the reference's dict (example_configs/*.py).  process_group: torch.distributed group for the
        data-parallel gradient sum (None = single replica / default group).  dp_mode: "p2p" = the fused NVLink
        peer-memory exchange + Adam kernel (parallel.PeerAdam), "nccl" = NCCL all-reduce followed by Adam."""
        self.config = dict(config)
        cfg = self.config
        if cfg.get("loss") == "cross_entropy":               # main.py:122-125 swaps in the cross-entropy pair
            raise ValueError("config['loss'] = 'cross_entropy': only the hinge losses (sagan/main.py:21-27, what every "
                             "shipped config uses) are built")
        self.B = cfg["batch_size"]
        self.dp = ReplicaGradientSum(process_group)
        self.world = self.dp.world
        # main.py:358: global_batch_size = batch_size * len(gpu)
        self.global_batch = global_batch_size or self.dp.global_batch(self.B)
        self.device = torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(seed)            # initial weights (replica 0's are broadcast below)
        self.dp_mode = dp_mode if self.world > 1 else "none"
        if self.dp_mode == "p2p":
            nets.set_flat_allocator(symmetric_allocator())
        if cfg.get("model", "vanilla") == "resnet":      # sagan/main.py:101-107 (the branch the reference left disabled)
            if not cfg.get("use_label"):
                raise ValueError("the residual topologies are class-conditional (models/generator.py:25): set use_label")
            self.G, self.D = nets.get_res_generator(cfg), nets.get_res_discriminator(cfg)
        else:
            self.G, self.D = nets.get_generator(cfg), nets.get_discriminator(cfg)
        with torch.no_grad():   # build pass (Keras `model.build`, main.py:134-135)
            z = torch.zeros(self.B, cfg["z_dim"], device=self.device)
            lab = torch.zeros(self.B, dtype=torch.int64, device=self.device) if cfg.get("use_label") else None
            img = self.G([z, lab])
            self.D([img, lab])
            # Keras `model.build` creates variables without running data: the build pass above must not leave its
            # all-zero batch in the BatchNormalization moving statistics (they feed the training=False forward)
            for m in self.G.modules():
                if hasattr(m, "moving_mean") and hasattr(m, "moving_var"):
                    m.moving_mean.zero_()
                    m.moving_var.fill_(1.0)
        # identical initial weights and spectral-norm state on every replica (MirroredStrategy variables)
        self.dp.broadcast_(self.G.flat_params, self.D.flat_params, self.G.sn_group.out, self.D.sn_group.out)
        # every replica draws its OWN noise / fake labels (MirroredStrategy runs main.py:176-177,194-195 per replica):
        # the device generator is re-seeded per replica once the weights are identical everywhere
        torch.cuda.manual_seed(seed * 9973 + self.dp.rank)
        ur = cfg.get("update_ratio", 1)
        self.opt_G = FlatAdam(self.G, cfg["lr_g"], steps_per_epoch, cfg["decay_rate"])          # main.py:111-114
        self.opt_D = FlatAdam(self.D, cfg["lr_d"], steps_per_epoch * ur, cfg["decay_rate"])     # main.py:115-118
        nets.set_flat_allocator(None)
        self.peer_G = self.peer_D = None
        self.loss_sums = torch.zeros(2, device=self.device)     # [sum L_D (over update_ratio), sum L_G]
        if self.dp_mode == "p2p":
            # the loss sums live in symmetric memory: G's exchange kernel (the last kernel of a step, when both sums are
            # complete on every replica) also reduces them over the replicas (main.py:216-220)
            self.loss_sums = symmetric_allocator()(2, self.device)
            self.peer_G = PeerAdam(self.G, self.opt_G, self.dp, loss_sums=self.loss_sums)
            self.peer_D = PeerAdam(self.D, self.opt_D, self.dp)
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
            self.opt_D.schedule()
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
            self.opt_G.schedule()
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

    # ------------------------------------------------------------------------------------------
    def train_step(self, images, labels=None, noises_d=None, noise_g=None, fake_labels_d=None, fake_labels_g=None):
        """Eager step.  images: device NHWC float32 in [-1,1] (sagan/dataset.py:34).  Noise may be injected
        (parity tests); otherwise it is drawn on the device (main.py:176,194).  Returns the device tensor
        [sum L_D, sum L_G]; see `losses()` for the reported means."""
        if images.dtype == torch.uint8:          # raw records (sagan/dataset.py:31-34): decoded on the device
            images = F.decode_records(images.contiguous())
        self._step_body(images, labels, noises_d, noise_g, fake_labels_d, fake_labels_g)
        return self.loss_sums

    def sample(self, noise, labels=None):
        """sagan/main.py:333: generator([fixed_vector, fixed_label], training=False) -- BatchNormalization on its moving
        statistics, spectral normalisation with the W / sigma of the latest training forward (no power iteration).
        Returns images NHWC in (-1, 1); main.py:334 maps them to uint8 with x * 127.5 + 128."""
        with torch.no_grad():
            return self.G([noise, labels], training=False)

    def losses(self, reduce=True):
        """Reported losses (main.py:216-229): sum over the GLOBAL batch (all replicas, `strategy.reduce(SUM)`) / global
        batch, mean over the rest.  Synchronises (device -> host read of two floats) and raises if the peer-memory
        exchange reported a lost replica.  reduce=False keeps the replica-local sums (divided by the per-replica
        batch)."""
        sums = self.loss_sums
        denom = self.global_batch
        if self.world > 1:
            if not reduce:
                denom = self.B
            elif self.peer_G is not None:
                sums = self.peer_G.loss_global          # summed over the replicas by the step's last exchange kernel
            else:
                sums = self.dp.sum_losses_(sums.clone())
        s = sums.tolist()
        self.check_exchange()
        ur = self.config.get("update_ratio", 1)
        n_elem = self._logit_elems()
        return dict(D_loss=s[0] / ur / (denom * n_elem), G_loss=s[1] / (denom * n_elem))

    def check_exchange(self):
        """Raises if a replica timed out at an exchange barrier (the replicas have diverged).  Synchronises."""
        for peer in (self.peer_D, self.peer_G):
            if peer is not None:
                peer.check()

    def _logit_elems(self):
        return 1 if self.config.get("use_label") else 16     # [B,4,4,1] patch logits (discriminator.py:35)

    # ------------------------------------------------------------------------------------------
    def _state_tensors(self):
        """Everything a training step mutates: weights, Adam moments and counters, spectral-norm u / v / W_bar,
        BatchNorm moving statistics, loss sums."""
        ts = [self.G.flat_params, self.D.flat_params, self.G.sn_group.out, self.D.sn_group.out,
              self.opt_G.v, self.opt_D.v, self.opt_G._iters, self.opt_D._iters, self.loss_sums]
        for peer in (self.peer_G, self.peer_D):
            if peer is not None:
                ts.append(peer.v_shard)
        for net in (self.G, self.D):
            for m in net.modules():
                if hasattr(m, "moving_mean"):
                    ts += [m.moving_mean, m.moving_var]
        return ts

    def capture(self, warmup=3, static_noise=False, uint8_input=False):
        """Capture the whole step (both phases, gradient exchange, LR schedule, Adam) into one CUDA graph.  Images (and
        labels) are read from static device buffers that `graph_step` refills.  The warm-up steps that CUDA-graph
        capture needs run on throw-away state: weights, optimiser state and counters, spectral-norm vectors, BatchNorm
        statistics and the device RNG are restored afterwards, so capturing does not train.
        static_noise=True additionally routes the latent noise (and fake labels) through static buffers
        (`graph_step(noises_d=..., noise_g=...)`) instead of drawing them inside the graph.
        uint8_input=True: the graph takes the batch as the reference's raw uint8 records (sagan/dataset.py:27-40) and
        starts with the decode `x * (2. / 255) - 1.`; `graph_step` then expects uint8 images (4x fewer bytes to copy)."""
        cfg = self.config
        ur = cfg.get("update_ratio", 1)
        self._static["images"] = torch.zeros(self.B, cfg["img_size"], cfg["img_size"], 3, device=self.device)
        self._static["labels"] = (torch.zeros(self.B, dtype=torch.int64, device=self.device)
                                  if cfg.get("use_label") else None)
        self._static["images_u8"] = (torch.full((self.B, cfg["img_size"], cfg["img_size"], 3), 128, dtype=torch.uint8,
                                                device=self.device) if uint8_input else None)
        nz_d = nz_g = fl_d = fl_g = None
        if static_noise:
            nz_d = [torch.zeros(self.B, cfg["z_dim"], device=self.device) for _ in range(ur)]
            nz_g = torch.zeros(self.B, cfg["z_dim"], device=self.device)
            if cfg.get("use_label"):
                fl_d = [torch.zeros(self.B, dtype=torch.int64, device=self.device) for _ in range(ur)]
                fl_g = torch.zeros(self.B, dtype=torch.int64, device=self.device)
        self._static.update(noises_d=nz_d, noise_g=nz_g, fake_labels_d=fl_d, fake_labels_g=fl_g)
        torch.cuda.synchronize()
        saved = [t.clone() for t in self._state_tensors()]
        rng = torch.cuda.get_rng_state(self.device)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_body(self._static["images"], self._static["labels"], nz_d, nz_g, fl_d, fl_g)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.dp.barrier()              # no replica restores its buffers while a peer is still writing into them
        for t, t0 in zip(self._state_tensors(), saved):
            t.copy_(t0)
        torch.cuda.set_rng_state(rng, self.device)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if uint8_input:
                F.decode_records(self._static["images_u8"], out=self._static["images"])
            self._step_body(self._static["images"], self._static["labels"], nz_d, nz_g, fl_d, fl_g)
        return self.graph

    def graph_step(self, images=None, labels=None, noises_d=None, noise_g=None, fake_labels_d=None, fake_labels_g=None):
        """Replay the captured step.  `images` may be a pinned host tensor (copied asynchronously) or a
        device tensor; None keeps the buffer contents.  Noise / fake labels can be given only after
        `capture(static_noise=True)`."""
        if self.graph is None:
            raise RuntimeError("call capture() first")
        st = self._static
        if images is not None:
            if (images.dtype == torch.uint8) != (st["images_u8"] is not None):
                raise RuntimeError("graph_step: uint8 records need capture(uint8_input=True), float images need "
                                   "capture(uint8_input=False)")
            (st["images_u8"] if images.dtype == torch.uint8 else st["images"]).copy_(images, non_blocking=True)
        if labels is not None and st["labels"] is not None:
            st["labels"].copy_(labels, non_blocking=True)
        for given, key in ((noises_d, "noises_d"), (fake_labels_d, "fake_labels_d")):
            if given is not None:
                if st[key] is None:
                    raise RuntimeError(f"{key} given, but the graph was captured without static noise buffers")
                for dst, src in zip(st[key], given):
                    dst.copy_(src, non_blocking=True)
        for given, key in ((noise_g, "noise_g"), (fake_labels_g, "fake_labels_g")):
            if given is not None:
                if st[key] is None:
                    raise RuntimeError(f"{key} given, but the graph was captured without static noise buffers")
                st[key].copy_(given, non_blocking=True)
        self.graph.replay()
        return self.loss_sums
