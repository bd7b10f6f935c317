"""torch.autograd.Function wrappers over the C ABI (raw device pointers + the current CUDA stream).

torch supplies device memory, the stream and the autograd tape (the role tf.GradientTape plays in
/root/reference/sagan/main.py:180,197); all arithmetic runs in libsagan_b200.so.  Tensors must be
CUDA fp32 and contiguous -- anything else raises (no CPU path).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_TANH, MATH_BF16_TC, MATH_FP32_STRICT, ConvGeom, SnDesc, check  # noqa: F401


def _ptr(t):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise _lib.SaganError(f"expected a contiguous CUDA float32 tensor, got {t.device} {t.dtype} "
                              f"contiguous={t.is_contiguous()} (there is no CPU fallback)")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def same_pad(n, k, s):
    """TF padding='same': returns (pad_before, out)."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, out


def conv_geom(B, H, W, Cin, Cout, kh, kw, stride, padding):
    if padding == "same":
        pt, Ho = same_pad(H, kh, stride)
        pl, Wo = same_pad(W, kw, stride)
    elif padding == "valid":
        pt = pl = 0
        Ho, Wo = (H - kh) // stride + 1, (W - kw) // stride + 1
    else:
        raise ValueError(f"padding must be 'same' or 'valid', got {padding!r}")
    return ConvGeom(B, H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pt, pl)


def deconv_geom(B, H, W, Cin, Cout, kh, kw, stride, padding):
    """Geometry of the forward conv whose gradient is Conv2DTranspose(x[B,H,W,Cin]) -> [B,H*s,W*s,Cout]:
    the conv maps the (big) output grid back onto the (small) input grid."""
    if padding != "same":
        raise ValueError("Conv2DTranspose supports padding='same' only")
    Hb, Wb = H * stride, W * stride
    pt, Ho = same_pad(Hb, kh, stride)
    pl, Wo = same_pad(Wb, kw, stride)
    assert Ho == H and Wo == W
    # conv: big [B,Hb,Wb,Cout] -> small [B,H,W,Cin]; Keras kernel [kh,kw,Cout,Cin] is its HWIO kernel
    return ConvGeom(B, Hb, Wb, Cout, H, W, Cin, kh, kw, stride, pt, pl)


# ------------------------------------------------------------------------------------ conv
class _Conv2dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, geom, act, slope, math_mode):
        lib = _lib.load()
        y = torch.empty((geom.B, geom.Ho, geom.Wo, geom.Cout), device=x.device, dtype=torch.float32)
        check(lib.sagan_conv2d_fwd(_ptr(x), _ptr(w), _ptr(bias), _ptr(y), C.byref(geom), act, slope, math_mode,
                                   _stream()), "sagan_conv2d_fwd")
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        ctx.geom, ctx.act, ctx.slope, ctx.mm, ctx.has_bias = geom, act, slope, math_mode, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w, y = ctx.saved_tensors
        g = ctx.geom
        dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dz = torch.empty_like(dy)
            check(lib.sagan_act_bwd(_ptr(y), _ptr(dy), _ptr(dz), dy.numel(), ctx.act, ctx.slope, _stream()),
                  "sagan_act_bwd")
            dy = dz
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            check(lib.sagan_conv2d_dgrad(_ptr(dy), _ptr(w), _ptr(dx), C.byref(g), ctx.mm, _stream()),
                  "sagan_conv2d_dgrad")
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = torch.empty(g.Cout, device=x.device, dtype=torch.float32)
            check(lib.sagan_conv2d_wgrad(_ptr(x), _ptr(dy), _ptr(dw), _ptr(db), C.byref(g), ctx.mm, _stream()),
                  "sagan_conv2d_wgrad")
        return dx, dw, db, None, None, None, None


def conv2d(x, w, bias, stride=1, padding="same", act=ACT_NONE, slope=0.0, math_mode=MATH_FP32_STRICT):
    """x NHWC, w HWIO [kh,kw,cin,cout]."""
    B, H, W, Cin = x.shape
    kh, kw, cin, cout = w.shape
    if cin != Cin:
        raise ValueError(f"conv2d: input has {Cin} channels, kernel expects {cin}")
    return _Conv2dFn.apply(x.contiguous(), w.contiguous(), bias, conv_geom(B, H, W, Cin, cout, kh, kw, stride, padding),
                           act, float(slope), math_mode)


class _Conv2dTransposeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, geom, math_mode):
        lib = _lib.load()
        y = torch.empty((geom.B, geom.H, geom.W, geom.Cin), device=x.device, dtype=torch.float32)
        # Conv2DTranspose forward == backward-data of the conv described by geom
        check(lib.sagan_conv2d_dgrad(_ptr(x), _ptr(w), _ptr(y), C.byref(geom), math_mode, _stream()),
              "sagan_conv2d_dgrad (deconv fwd)")
        ctx.save_for_backward(x, w)
        ctx.geom, ctx.mm = geom, math_mode
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        g = ctx.geom
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            check(lib.sagan_conv2d_fwd(_ptr(dy), _ptr(w), None, _ptr(dx), C.byref(g), ACT_NONE, 0.0, ctx.mm, _stream()),
                  "sagan_conv2d_fwd (deconv dgrad)")
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            check(lib.sagan_conv2d_wgrad(_ptr(dy), _ptr(x), _ptr(dw), None, C.byref(g), ctx.mm, _stream()),
                  "sagan_conv2d_wgrad (deconv wgrad)")
        return dx, dw, None, None


def conv2d_transpose(x, w, stride=2, padding="same", math_mode=MATH_FP32_STRICT):
    """x NHWC [B,H,W,cin], w Keras Conv2DTranspose kernel [kh,kw,cout,cin] -> [B,H*s,W*s,cout]."""
    B, H, W, Cin = x.shape
    kh, kw, cout, cin = w.shape
    if cin != Cin:
        raise ValueError(f"conv2d_transpose: input has {Cin} channels, kernel expects {cin}")
    return _Conv2dTransposeFn.apply(x.contiguous(), w.contiguous(),
                                    deconv_geom(B, H, W, Cin, cout, kh, kw, stride, padding), math_mode)


def dense(x, w, bias, math_mode=MATH_FP32_STRICT):
    """x [B, in], w [in, out] -- the H = W = k = 1 convolution."""
    B, In = x.shape
    y = _Conv2dFn.apply(x.contiguous().view(B, 1, 1, In), w.contiguous().view(1, 1, In, w.shape[1]), bias,
                        ConvGeom(B, 1, 1, In, 1, 1, w.shape[1], 1, 1, 1, 0, 0), ACT_NONE, 0.0, math_mode)
    return y.view(B, w.shape[1])


# ------------------------------------------------------------------------------------ batch norm
class _BnLreluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, moving_mean, moving_var, eps, momentum, slope):
        lib = _lib.load()
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        y = torch.empty_like(x)
        mean = torch.empty(Cc, device=x.device, dtype=torch.float32)
        invstd = torch.empty_like(mean)
        wsb = lib.sagan_bn_workspace_bytes(Cc)
        ws = torch.empty(wsb // 4, device=x.device, dtype=torch.float32)
        check(lib.sagan_bn_lrelu_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean), _ptr(invstd),
                                     _ptr(moving_mean), _ptr(moving_var), rows, Cc, eps, momentum, slope,
                                     _ptr(ws), wsb, _stream()), "sagan_bn_lrelu_fwd")
        ctx.save_for_backward(x, y, gamma, mean, invstd)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, y, gamma, mean, invstd = ctx.saved_tensors
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        wsb = lib.sagan_bn_workspace_bytes(Cc)
        ws = torch.empty(wsb // 4, device=x.device, dtype=torch.float32)
        check(lib.sagan_bn_lrelu_bwd(_ptr(dy), _ptr(x), _ptr(y), _ptr(gamma), _ptr(mean), _ptr(invstd), _ptr(dx),
                                     _ptr(dgamma), _ptr(dbeta), rows, Cc, ctx.slope, _ptr(ws), wsb, _stream()),
              "sagan_bn_lrelu_bwd")
        return dx, dgamma, dbeta, None, None, None, None, None


def batchnorm_lrelu(x, gamma, beta, moving_mean=None, moving_var=None, eps=1e-3, momentum=0.99, slope=0.1):
    """Training-mode BatchNormalization followed by LeakyReLU(slope); slope = 1 gives plain BN."""
    return _BnLreluFn.apply(x.contiguous(), gamma, beta, moving_mean, moving_var, float(eps), float(momentum),
                            float(slope))


def batchnorm_lrelu_infer(x, gamma, beta, moving_mean, moving_var, eps=1e-3, slope=0.1):
    """Inference-mode BatchNormalization (moving statistics) followed by LeakyReLU(slope).  No gradient: this is the
    `training=False` forward of the sample dumps (sagan/main.py:333)."""
    x = x.detach().contiguous()
    y = torch.empty_like(x)
    Cc = x.shape[-1]
    check(_lib.load().sagan_bn_lrelu_infer(_ptr(x), _ptr(gamma.detach()), _ptr(beta.detach()), _ptr(moving_mean),
                                           _ptr(moving_var), _ptr(y), x.numel() // Cc, Cc, float(eps), float(slope),
                                           _stream()), "sagan_bn_lrelu_infer")
    return y


# ------------------------------------------------------------------------------------ attention
class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, wo, bo, gamma, math_mode, grid):
        """grid = None: every token is a key (layers.py:93-120, paper form); grid = (H, W): keys / values max-pooled
        2x2 / stride 2 over the token grid (sagan_attn_pool_*)."""
        lib = _lib.load()
        B, N, Cc = x.shape
        dv = Cc // 2
        y = torch.empty_like(x)
        lse = torch.empty((B, N), device=x.device, dtype=torch.float32)
        a = torch.empty((B, N, dv), device=x.device, dtype=torch.float32)
        ptrs = (_ptr(x), _ptr(wq), _ptr(bq), _ptr(wk), _ptr(bk), _ptr(wv), _ptr(bv), _ptr(wo), _ptr(bo), _ptr(gamma),
                _ptr(y), _ptr(lse), _ptr(a))
        if grid is None:
            wsb = lib.sagan_attn_workspace_bytes(B, N, Cc, math_mode)
            ws = torch.empty((wsb + 3) // 4, device=x.device, dtype=torch.float32)
            check(lib.sagan_attn_fwd(*ptrs, B, N, Cc, math_mode, _ptr(ws), wsb, _stream()), "sagan_attn_fwd")
        else:
            H, W = grid
            wsb = lib.sagan_attn_pool_workspace_bytes(B, H, W, Cc, math_mode)
            ws = torch.empty((wsb + 3) // 4, device=x.device, dtype=torch.float32)
            check(lib.sagan_attn_pool_fwd(*ptrs, B, H, W, Cc, math_mode, _ptr(ws), wsb, _stream()), "sagan_attn_pool_fwd")
        ctx.save_for_backward(x, wq, bq, wk, bk, wv, bv, wo, bo, gamma, lse, a)
        ctx.mm, ctx.grid = math_mode, grid
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, wq, bq, wk, bk, wv, bv, wo, bo, gamma, lse, a = ctx.saved_tensors
        B, N, Cc = x.shape
        dy = dy.contiguous()
        need_w = any(ctx.needs_input_grad[1:10])
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        if need_w:
            gs = [torch.empty_like(t) for t in (wq, bq, wk, bk, wv, bv, wo, bo, gamma)]
        else:
            gs = [None] * 9
        if dx is None and not need_w:
            return (None,) * 12
        ptrs = (_ptr(dy), _ptr(x), _ptr(wq), _ptr(bq), _ptr(wk), _ptr(bk), _ptr(wv), _ptr(bv), _ptr(wo), _ptr(bo),
                _ptr(gamma), _ptr(lse), _ptr(a), _ptr(dx), *[_ptr(g) for g in gs])
        if ctx.grid is None:
            wsb = lib.sagan_attn_workspace_bytes(B, N, Cc, ctx.mm)
            ws = torch.empty((wsb + 3) // 4, device=x.device, dtype=torch.float32)
            check(lib.sagan_attn_bwd(*ptrs, B, N, Cc, ctx.mm, _ptr(ws), wsb, _stream()), "sagan_attn_bwd")
        else:
            H, W = ctx.grid
            wsb = lib.sagan_attn_pool_workspace_bytes(B, H, W, Cc, ctx.mm)
            ws = torch.empty((wsb + 3) // 4, device=x.device, dtype=torch.float32)
            check(lib.sagan_attn_pool_bwd(*ptrs, B, H, W, Cc, ctx.mm, _ptr(ws), wsb, _stream()), "sagan_attn_pool_bwd")
        return (dx, *gs, None, None)


def attention(x, wq, bq, wk, bk, wv, bv, wo, bo, gamma, math_mode=MATH_FP32_STRICT, pool_grid=None):
    """x [B,N,C]; wq (theta) / wk (phi) [C,C//8]; wv (g) [C,C//2]; wo [C//2,C]; gamma 0-d tensor.
    pool_grid = (H, W) with H * W == N: keys (phi) and values (g) are max-pooled 2x2 / stride 2 over the token grid,
    what layers.py:96,100,113 reaches for (N / 4 keys)."""
    if pool_grid is not None:
        H, W = int(pool_grid[0]), int(pool_grid[1])
        if H * W != x.shape[1]:
            raise ValueError(f"pool_grid {pool_grid} does not match {x.shape[1]} tokens")
        pool_grid = (H, W)
    return _AttnFn.apply(x.contiguous(), wq.contiguous(), bq, wk.contiguous(), bk, wv.contiguous(), bv,
                         wo.contiguous(), bo, gamma.reshape(1), math_mode, pool_grid)


# ------------------------------------------------------------------------------------ spectral norm
class SpectralNormGroup:
    """All spectrally-normalised kernels of one network, normalised by ONE cooperative launch.

    Owns one flat buffer [W_bar... | v... | u... | sigma...]: the persistent `u` vectors live in it
    (updated in place by every run) and the plan writes W_bar, v and sigma into it.  Every training
    forward snapshots the buffer (one copy), so several forwards inside one autograd tape
    (D(real), D(fake): sagan/main.py:181-182) keep their own W_bar / u / v / sigma for the backward.
    """

    def __init__(self, weights, u_init, Ip=1, factors=None):
        lib = _lib.load()
        self.weights = list(weights)
        n = len(self.weights)
        self.shapes = [(w.shape[-1], w.numel() // w.shape[-1]) for w in self.weights]
        dev = self.weights[0].device
        self.Ip = [Ip] * n if isinstance(Ip, int) else list(Ip)
        self.factors = [0.0 if not f else float(f) for f in (factors or [None] * n)]
        pad = lambda k: (k + 63) // 64 * 64      # keep every region 256-byte aligned
        self.off_w, self.off_v, self.off_s, self.off_u = [], [], [], []
        tot = 0
        for (R, K) in self.shapes:
            self.off_w.append(tot); tot += pad(R * K)
        for (R, K) in self.shapes:
            self.off_v.append(tot); tot += pad(K)
        for (R, K) in self.shapes:
            self.off_u.append(tot); tot += pad(R)
        for _ in self.shapes:
            self.off_s.append(tot); tot += 64
        self.out = torch.zeros(tot, device=dev, dtype=torch.float32)
        descs = (SnDesc * n)()
        for i, w in enumerate(self.weights):
            R, K = self.shapes[i]
            u0 = u_init[i].detach().reshape(-1).to(device=dev, dtype=torch.float32)
            if u0.numel() != R:
                raise _lib.SaganError(f"spectral-norm u of kernel {i} must have {R} elements, got {u0.numel()}")
            self.u(i).copy_(u0)
            descs[i] = SnDesc(_ptr(w.detach()), self.u(i).data_ptr(), self.v(i).data_ptr(),
                              self.out[self.off_w[i]:].data_ptr(), None, self.out[self.off_s[i]:].data_ptr(),
                              R, K, self.Ip[i], self.factors[i])
        plan = C.c_void_p()
        check(lib.sagan_sn_plan_create(descs, n, dev.index if dev.index is not None else torch.cuda.current_device(),
                                       C.byref(plan)), "sagan_sn_plan_create")
        self.plan = plan
        self.algorithmic_bytes = int(lib.sagan_sn_plan_algorithmic_bytes(plan))
        self._bws = torch.empty(max(lib.sagan_sn_backward_workspace_bytes(0) // 4, 16 * 64), device=dev, dtype=torch.float32)

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                _lib.load().sagan_sn_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # persistent state views
    def u(self, i):
        return self.out[self.off_u[i]:self.off_u[i] + self.shapes[i][0]]

    def v(self, i):
        return self.out[self.off_v[i]:self.off_v[i] + self.shapes[i][1]]

    def sigma(self, i):
        return self.out[self.off_s[i]:self.off_s[i] + 1]

    def w_bar(self, i):
        R, K = self.shapes[i]
        return self.out[self.off_w[i]:self.off_w[i] + R * K].view(self.weights[i].shape)

    def run(self):
        """Power iteration + W/sigma for every kernel of the group (layers.py:50-68), one launch."""
        check(_lib.load().sagan_sn_plan_run(self.plan, _stream()), "sagan_sn_plan_run")

    def phase_times_ms(self):
        """Device-side durations of the five phases of the last run (diagnostics; synchronises)."""
        out = (C.c_float * 5)()
        check(_lib.load().sagan_sn_plan_phase_times(self.plan, out), "sagan_sn_plan_phase_times")
        return list(out)

    def views(self, snap, i):
        R, K = self.shapes[i]
        wbar = snap[self.off_w[i]:self.off_w[i] + R * K].view(self.weights[i].shape)
        return (wbar, snap[self.off_u[i]:self.off_u[i] + R], snap[self.off_v[i]:self.off_v[i] + K],
                snap[self.off_s[i]:self.off_s[i] + 1])

    def refresh(self):
        """training=False: sigma from the stored u, v and the current kernels, W_bar = W / sigma, no iteration."""
        check(_lib.load().sagan_sn_plan_refresh(self.plan, _stream()), "sagan_sn_plan_refresh")

    def normalized(self, update=True):
        """Returns the list of W_bar tensors (autograd-connected to the raw kernels).  update=True (training): one
        power iteration first (layers.py:46 in the oracle's reading); update=False: `refresh` -- u and v stay."""
        if update:
            self.run()
        else:
            self.refresh()
        holder = _Snap(self.out.clone())
        return list(_SnGroupFn.apply(self, holder, *self.weights))


class _Snap:
    """Opaque holder so autograd does not treat the snapshot as a differentiable input."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


class _SnGroupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, group, holder, *weights):
        ctx.group, ctx.holder = group, holder
        return tuple(group.views(holder.t, i)[0] for i in range(len(weights)))

    @staticmethod
    def backward(ctx, *dwbars):
        """All kernels of the group in two launches (sagan_sn_backward_multi).  A kernel whose `.grad` is already
        allocated (the networks' flat gradient bucket) is ACCUMULATED in place and autograd gets None for it, which
        removes one elementwise add launch per kernel."""
        lib = _lib.load()
        g, snap = ctx.group, ctx.holder.t
        todo = [i for i, dwb in enumerate(dwbars) if dwb is not None and ctx.needs_input_grad[2 + i]]
        grads = [None] * len(dwbars)
        keep = []
        for mode in (1, 0):                       # 1: accumulate into existing .grad, 0: fresh output tensors
            sel = [i for i in todo if (g.weights[i].grad is not None) == bool(mode)]
            for a in range(0, len(sel), 16):
                chunk = sel[a:a + 16]
                descs = (_lib.SnBwdDesc * len(chunk))()
                for j, i in enumerate(chunk):
                    wbar, u, v, sigma = g.views(snap, i)
                    R, K = g.shapes[i]
                    dwb = dwbars[i].contiguous()
                    keep.append(dwb)
                    if mode:
                        out = g.weights[i].grad
                    else:
                        out = torch.empty_like(dwb)
                        grads[i] = out
                    descs[j] = _lib.SnBwdDesc(_ptr(dwb), _ptr(wbar), _ptr(u), _ptr(v), _ptr(sigma), _ptr(out),
                                              g.factors[i], R, K)
                check(lib.sagan_sn_backward_multi(descs, len(chunk), mode, _ptr(g._bws), g._bws.numel() * 4, _stream()),
                      "sagan_sn_backward_multi")
        return (None, None, *grads)


class _ParamSinkFn(torch.autograd.Function):
    """Proxies for a network's parameters that are NOT spectrally normalised (biases, BatchNorm gamma / beta, attention
    gamma, un-normalised head kernels).  Forward hands out aliases of the parameters; backward receives all their
    gradients at once and adds them into the pre-allocated `.grad` views (the network's flat gradient bucket) with ONE
    multi-tensor launch, returning None to autograd -- instead of one accumulation kernel per parameter."""

    @staticmethod
    def forward(ctx, holder, *params):
        ctx.params = params
        return tuple(p.detach() for p in params)

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        out = [None] * len(grads)
        todo = []
        for i, (p, g) in enumerate(zip(ctx.params, grads)):
            if g is None or not ctx.needs_input_grad[1 + i]:
                continue
            if p.grad is None:
                out[i] = g                     # no bucket: let autograd keep the gradient
            else:
                todo.append((p.grad, g.contiguous()))
        for a in range(0, len(todo), 64):
            chunk = todo[a:a + 64]
            descs = (_lib.AccDesc * len(chunk))()
            for j, (dst, src) in enumerate(chunk):
                descs[j] = _lib.AccDesc(_ptr(dst), _ptr(src), src.numel())
            check(lib.sagan_accumulate_multi(descs, len(chunk), _stream()), "sagan_accumulate_multi")
        return (None, *out)


def param_proxies(params):
    """Aliases of `params` whose gradients are accumulated into `p.grad` by one launch (see _ParamSinkFn)."""
    return _ParamSinkFn.apply(None, *params)


# ------------------------------------------------------------------------------------ losses / optimiser
def hinge_d_grads(d_real, d_fake, global_batch, loss_sum):
    """sagan/main.py:24-27,183-184: accumulates sum(L) into loss_sum[0]; returns d(mean(L)/global_batch)/d logits."""
    lib = _lib.load()
    n = d_real.numel()
    g_real, g_fake = torch.empty_like(d_real), torch.empty_like(d_fake)
    check(lib.sagan_hinge_d(_ptr(d_real.detach()), _ptr(d_fake.detach()), n, 1.0 / (n * global_batch), _ptr(loss_sum),
                            _ptr(g_real), _ptr(g_fake), _stream()), "sagan_hinge_d")
    return g_real, g_fake


def hinge_g_grads(d_fake, global_batch, loss_sum):
    """sagan/main.py:21-22,200-201."""
    lib = _lib.load()
    n = d_fake.numel()
    g_fake = torch.empty_like(d_fake)
    check(lib.sagan_hinge_g(_ptr(d_fake.detach()), n, 1.0 / (n * global_batch), _ptr(loss_sum), _ptr(g_fake), _stream()),
          "sagan_hinge_g")
    return g_fake


def adam_step(param, grad, v, hyper, m=None, grad_scale=1.0):
    """Keras Adam over a flat bucket (sagan/main.py:119-120); hyper = device [lr_t, b1, b2, eps]."""
    check(_lib.load().sagan_adam_step(_ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(), _ptr(hyper),
                                      float(grad_scale), _stream()), "sagan_adam_step")


def adam_schedule(hyper, iterations, lr0, decay_rate, decay_steps, b1, b2, eps):
    """sagan/main.py:111-120 on the device: hyper <- [lr_t, b1, b2, eps] from the device counter `iterations`
    (int64 [1]), which is then incremented."""
    if not (iterations.is_cuda and iterations.dtype == torch.int64):
        raise _lib.SaganError("adam_schedule: `iterations` must be a CUDA int64 tensor")
    check(_lib.load().sagan_adam_schedule(_ptr(hyper), iterations.data_ptr(), float(lr0), float(decay_rate),
                                          int(decay_steps), float(b1), float(b2), float(eps), _stream()),
          "sagan_adam_schedule")


# ------------------------------------------------------------------------------------ weight normalisation / records
class _WeightNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, g):
        lib = _lib.load()
        cols = v.shape[-1]
        rows = v.numel() // cols
        w = torch.empty_like(v)
        inv = torch.empty(cols, device=v.device, dtype=torch.float32)
        check(lib.sagan_wn_fwd(_ptr(v), _ptr(g), _ptr(w), _ptr(inv), rows, cols, _stream()), "sagan_wn_fwd")
        ctx.save_for_backward(v, g, inv)
        return w

    @staticmethod
    def backward(ctx, dw):
        lib = _lib.load()
        v, g, inv = ctx.saved_tensors
        cols = v.shape[-1]
        rows = v.numel() // cols
        dv, dg, ws = torch.empty_like(v), torch.empty_like(g), torch.empty_like(g)
        check(lib.sagan_wn_bwd(_ptr(dw.contiguous()), _ptr(v), _ptr(g), _ptr(inv), _ptr(dv), _ptr(dg), _ptr(ws), rows, cols,
                               _stream()), "sagan_wn_bwd")
        return dv, dg


def weight_norm(v, g):
    """sagan/layers.py:124: kernel = l2_normalize(v, all axes but the last) * g."""
    return _WeightNormFn.apply(v.contiguous(), g.contiguous())


def batch_moments(x, eps):
    """Per-channel (last axis) mean and 1 / sqrt(var + eps) of x over all other axes (biased variance, tf.nn.moments)."""
    lib = _lib.load()
    x = x.contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    ones, zeros = torch.ones(Cc, device=x.device), torch.zeros(Cc, device=x.device)
    y = torch.empty_like(x)
    mean, invstd = torch.empty(Cc, device=x.device), torch.empty(Cc, device=x.device)
    wsb = lib.sagan_bn_workspace_bytes(Cc)
    ws = torch.empty(wsb // 4, device=x.device, dtype=torch.float32)
    check(lib.sagan_bn_lrelu_fwd(_ptr(x), _ptr(ones), _ptr(zeros), _ptr(y), _ptr(mean), _ptr(invstd), None, None, rows, Cc,
                                 float(eps), 0.0, 1.0, _ptr(ws), wsb, _stream()), "sagan_bn_lrelu_fwd")
    return mean, invstd


def decode_records(raw_u8, out=None):
    """sagan/dataset.py:31-34: uint8 HWC records -> float32 in [-1, 1] (`x * (2. / 255) - 1.`), on the device."""
    if not (raw_u8.is_cuda and raw_u8.dtype == torch.uint8 and raw_u8.is_contiguous()):
        raise _lib.SaganError("decode_records expects a contiguous CUDA uint8 tensor")
    if out is None:
        out = torch.empty(raw_u8.shape, device=raw_u8.device, dtype=torch.float32)
    check(_lib.load().sagan_u8_to_f32(raw_u8.data_ptr(), _ptr(out), raw_u8.numel(), 2.0 / 255, -1.0, _stream()),
          "sagan_u8_to_f32")
    return out


# ------------------------------------------------------------------------------------ elementwise glue (residual nets)
class _EwFn(torch.autograd.Function):
    """y = act(a + bias[c] + residual)."""

    @staticmethod
    def forward(ctx, a, bias, residual, act, slope):
        lib = _lib.load()
        y = torch.empty_like(a)
        Cc = a.shape[-1]
        check(lib.sagan_ew_fwd(_ptr(a), _ptr(bias), _ptr(residual), _ptr(y), a.numel(), Cc, act, slope, _stream()),
              "sagan_ew_fwd")
        ctx.save_for_backward(y if act != ACT_NONE else None)
        ctx.act, ctx.slope, ctx.C = act, slope, Cc
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dz = torch.empty_like(dy)
            check(lib.sagan_act_bwd(_ptr(y), _ptr(dy), _ptr(dz), dy.numel(), ctx.act, ctx.slope, _stream()), "sagan_act_bwd")
            dy = dz
        db = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            db = torch.empty(ctx.C, device=dy.device, dtype=torch.float32)
            check(lib.sagan_colsum(_ptr(dy), _ptr(db), dy.numel() // ctx.C, ctx.C, _stream()), "sagan_colsum")
        return dy, db, (dy if ctx.has_res else None), None, None


def activation(x, act=ACT_LRELU, slope=0.0):
    """Stand-alone ReLU (slope 0) / LeakyReLU / tanh."""
    return _EwFn.apply(x.contiguous(), None, None, act, float(slope))


def bias_add(x, bias, act=ACT_NONE, slope=0.0):
    return _EwFn.apply(x.contiguous(), bias, None, act, float(slope))


def add(a, b):
    """keras `layers.add([a, b])`."""
    return _EwFn.apply(a.contiguous(), None, b.contiguous(), ACT_NONE, 0.0)
