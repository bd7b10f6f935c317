"""Oracle: SAGAN training step (torch-CPU).  TEST INFRASTRUCTURE, see oracle/__init__.py.

Follows /root/reference/sagan/main.py:
  hinge_loss_g / hinge_loss_d   main.py:21-27
  LR schedules + Adam           main.py:111-120
  Trainer.train_step            main.py:171-211
  reported loss                 main.py:216-229
Keras semantics restated from knowledge: `optimizers.Adam` (OptimizerV2, TF 2.0)
  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
  var -= lr_t * m / (sqrt(v) + eps),  eps = 1e-7   (eps OUTSIDE the bias-corrected sqrt)
`ExponentialDecay(lr0, decay_steps, rate, staircase=True)`: lr0 * rate ** floor(step / decay_steps),
evaluated at optimizer.iterations BEFORE the increment.
"""
import torch

from . import nets

ADAM_B1 = 0.0        # main.py:119-120  beta_1=0.
ADAM_B2 = 0.999
ADAM_EPS = 1e-7


def hinge_loss_g(d_fake):
    return -d_fake                                                    # main.py:21-22


def hinge_loss_d(d_real, d_fake):
    return torch.relu(1.0 - d_real) + torch.relu(1.0 + d_fake)        # main.py:24-27


def exponential_decay(lr0, step, decay_steps, rate):
    return lr0 * rate ** (step // decay_steps)


class KerasAdam:
    def __init__(self, params, lr0, decay_steps, decay_rate, b1=ADAM_B1, b2=ADAM_B2, eps=ADAM_EPS):
        self.lr0, self.decay_steps, self.decay_rate = lr0, decay_steps, decay_rate
        self.b1, self.b2, self.eps = b1, b2, eps
        self.iterations = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def apply_gradients(self, params, grads):
        lr = exponential_decay(self.lr0, self.iterations, self.decay_steps, self.decay_rate)
        t = self.iterations + 1
        lr_t = lr * (1.0 - self.b2 ** t) ** 0.5 / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for k, g in grads.items():
                if g is None:
                    continue
                self.m[k].mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                self.v[k].mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                params[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + self.eps))
        self.iterations += 1


class OracleTrainer:
    """One replica of Trainer.train_step (main.py:171-211) with injected noise."""

    def __init__(self, cfg, dtype=torch.float32, seed=0, attn_sigma=0.0, bias_scale=0.0,
                 global_batch_size=None, steps_per_epoch=1000):
        self.cfg = dict(cfg)
        self.dtype = dtype
        self.gspec = nets.generator_spec(cfg)
        self.dspec = nets.discriminator_spec(cfg)
        self.G = nets.init_params(self.gspec, seed, dtype, attn_sigma, bias_scale)
        self.D = nets.init_params(self.dspec, seed + 100, dtype, attn_sigma, bias_scale)
        self.G_sn = nets.init_sn_state(self.gspec, seed + 1, dtype)
        self.D_sn = nets.init_sn_state(self.dspec, seed + 101, dtype)
        self.bn_stats = nets.init_bn_stats(self.gspec, dtype)
        self.global_batch = global_batch_size or cfg["batch_size"]
        ur = cfg.get("update_ratio", 1)
        self.opt_G = KerasAdam(self.G, cfg["lr_g"], steps_per_epoch, cfg["decay_rate"])        # main.py:111-114,119
        self.opt_D = KerasAdam(self.D, cfg["lr_d"], steps_per_epoch * ur, cfg["decay_rate"])   # main.py:115-118,120

    def _grads(self, loss, params):
        names = list(params.keys())
        gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
        return dict(zip(names, gs))

    def d_grads(self, images, noise, labels=None, fake_labels=None):
        cfg = self.cfg
        with torch.no_grad():                                                             # main.py:178 (outside the tape)
            fake = nets.generator_forward(self.G, self.G_sn, noise, cfg, fake_labels, True, self.bn_stats)
        for v in self.D.values():
            v.requires_grad_(True)
        d_real = nets.discriminator_forward(self.D, self.D_sn, images, cfg, labels, True)     # main.py:181
        d_fake = nets.discriminator_forward(self.D, self.D_sn, fake, cfg, fake_labels, True)  # main.py:182
        loss_elems = hinge_loss_d(d_real, d_fake)                                          # main.py:183
        scalar = loss_elems.mean() * (1.0 / self.global_batch)                             # main.py:184
        grads = self._grads(scalar, self.D)
        for v in self.D.values():
            v.requires_grad_(False)
        return grads, loss_elems.detach()

    def g_grads(self, noise, fake_labels=None):
        cfg = self.cfg
        for v in self.G.values():
            v.requires_grad_(True)
        fake = nets.generator_forward(self.G, self.G_sn, noise, cfg, fake_labels, True, self.bn_stats)   # main.py:198
        d_fake = nets.discriminator_forward(self.D, self.D_sn, fake, cfg, fake_labels, True)              # main.py:199
        loss_elems = hinge_loss_g(d_fake)                                                  # main.py:200
        scalar = loss_elems.mean() * (1.0 / self.global_batch)                             # main.py:201
        grads = self._grads(scalar, self.G)
        for v in self.G.values():
            v.requires_grad_(False)
        return grads, loss_elems.detach()

    def train_step(self, images, noises_d, noise_g, labels=None, fake_labels_d=None, fake_labels_g=None):
        """noises_d: list of `update_ratio` noise batches (main.py:175-191), noise_g: one batch (main.py:194)."""
        ur = self.cfg.get("update_ratio", 1)
        accu = 0.0
        for i in range(ur):
            fl = None if fake_labels_d is None else fake_labels_d[i]
            grads, le = self.d_grads(images, noises_d[i], labels, fl)
            accu = accu + le                                                               # main.py:186
            self.opt_D.apply_gradients(self.D, grads)                                      # main.py:190
        d_loss = accu / ur                                                                 # main.py:192
        grads, g_le = self.g_grads(noise_g, fake_labels_g)
        self.opt_G.apply_gradients(self.G, grads)                                          # main.py:205
        # reported loss (main.py:216-229): sum over replicas and batch / global batch, then Keras Mean
        return dict(G_loss=float(g_le.sum(0).div(self.global_batch).mean()),
                    D_loss=float(d_loss.sum(0).div(self.global_batch).mean()))
