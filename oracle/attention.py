"""Oracle: SAGAN self-attention block (numpy).  TEST INFRASTRUCTURE, see oracle/__init__.py.

Follows /root/reference/layers.py:71-120 (Attention_Layer.build / call) in the
well-formed reading fixed in SURVEY.md §8c(3): paper form, no key/value pooling,
phi^T a true transpose, logits NOT scaled by 1/sqrt(d).

  phi   = X Wphi   + bphi      [B,N,d]     layers.py:99      (keys)
  theta = X Wtheta + btheta    [B,N,d]     layers.py:104-105 (queries)
  g     = X Wg     + bg        [B,N,dv]    layers.py:112-114 (values)
  S     = theta phi^T          [B,N,N]     layers.py:108
  P     = softmax(S, -1)                   layers.py:109
  A     = P g                  [B,N,dv]    layers.py:116-117
  O     = A Wo + bo            [B,N,C]     layers.py:119
  Y     = X + gamma * O                    layers.py:120  (gamma = scalar `sigma`, layers.py:76-79)

The 1x1 kernels passed in are the EFFECTIVE kernels (spectral normalisation is
applied by the caller, oracle.sn); Keras kernel [1,1,cin,cout] == row-major [cin,cout].
Everything is materialised op by op (the [B,N,N] map included), like the TF graph.
"""
import numpy as np

WEIGHT_NAMES = ("Wphi", "bphi", "Wtheta", "btheta", "Wg", "bg", "Wo", "bo", "gamma")


def channel_split(C):
    # layers.py:82-85   c//8, c//8, c//2, c
    return C // 8, C // 2


def forward(X, Wphi, bphi, Wtheta, btheta, Wg, bg, Wo, bo, gamma, return_cache=False):
    """X [B,N,C] (NHWC with H*W flattened).  Returns Y [B,N,C]."""
    phi = X @ Wphi + bphi
    theta = X @ Wtheta + btheta
    g = X @ Wg + bg
    S = theta @ np.swapaxes(phi, 1, 2)
    S = S - S.max(axis=-1, keepdims=True)
    E = np.exp(S)
    P = E / E.sum(axis=-1, keepdims=True)
    A = P @ g
    O = A @ Wo + bo
    Y = X + gamma * O
    if return_cache:
        return Y, dict(phi=phi, theta=theta, g=g, P=P, A=A, O=O)
    return Y


def backward(dY, X, Wphi, bphi, Wtheta, btheta, Wg, bg, Wo, bo, gamma):
    """Analytic gradients; returns dict with dX and the 9 parameter grads.

    The reference has no hand-written backward (tf.GradientTape differentiates
    the ops above); tests cross-check this against torch autograd of `forward`
    and against finite differences.
    """
    Y, c = forward(X, Wphi, bphi, Wtheta, btheta, Wg, bg, Wo, bo, gamma, return_cache=True)
    phi, theta, g, P, A, O = c["phi"], c["theta"], c["g"], c["P"], c["A"], c["O"]
    C = X.shape[-1]
    dgamma = np.sum(dY * O)
    dO = gamma * dY
    dWo = np.einsum("bnv,bnc->vc", A, dO)
    dbo = dO.sum(axis=(0, 1))
    dA = dO @ Wo.T
    dg = np.swapaxes(P, 1, 2) @ dA
    dP = dA @ np.swapaxes(g, 1, 2)
    dS = P * (dP - np.sum(dP * P, axis=-1, keepdims=True))
    dtheta = dS @ phi
    dphi = np.swapaxes(dS, 1, 2) @ theta
    Xf = X.reshape(-1, C)
    out = dict(
        dX=dY + dtheta @ Wtheta.T + dphi @ Wphi.T + dg @ Wg.T,
        dWphi=Xf.T @ dphi.reshape(-1, dphi.shape[-1]), dbphi=dphi.sum(axis=(0, 1)),
        dWtheta=Xf.T @ dtheta.reshape(-1, dtheta.shape[-1]), dbtheta=dtheta.sum(axis=(0, 1)),
        dWg=Xf.T @ dg.reshape(-1, dg.shape[-1]), dbg=dg.sum(axis=(0, 1)),
        dWo=dWo, dbo=dbo, dgamma=dgamma,
    )
    return out


def make_inputs(B, N, C, seed=0, gamma=0.37, dtype=np.float64):
    """Synthetic block inputs of SURVEY.md §8d config 5: X ~ N(0,1), 1x1 weights ~ N(0, 1/C)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    d, dv = channel_split(C)
    s = 1.0 / np.sqrt(C)
    w = dict(
        Wphi=rng.standard_normal((C, d)) * s, bphi=rng.standard_normal(d) * 0.1,
        Wtheta=rng.standard_normal((C, d)) * s, btheta=rng.standard_normal(d) * 0.1,
        Wg=rng.standard_normal((C, dv)) * s, bg=rng.standard_normal(dv) * 0.1,
        Wo=rng.standard_normal((dv, C)) * (1.0 / np.sqrt(dv)), bo=rng.standard_normal(C) * 0.1,
    )
    w = {k: v.astype(dtype) for k, v in w.items()}
    w["gamma"] = dtype(gamma)
    X = rng.standard_normal((B, N, C)).astype(dtype)
    dY = rng.standard_normal((B, N, C)).astype(dtype)
    return X, dY, w


# ----------------------------------------------------------------------------------------------------------------------
# Down-sampled keys / values: what /root/reference/layers.py:96,100,113 reaches for -- the SAGAN
# paper's memory-saving variant max-pools phi and g over 2x2 windows before the attention map, so a query attends to
# N/4 keys.  The reference writes MaxPool2D(2, 1) (pool 2, stride 1: ill-formed); the well-formed
# reading fixed here is pool 2 / stride 2 / 'valid' on the [H, W] grid.  Not built in CUDA yet: this is the oracle the
# kernels of that row will be tested against.
def _pool2x2(t, hw):
    """t [B, H*W, c] -> pooled [B, (H/2)*(W/2), c] and the flat argmax index [B, (H/2)*(W/2), c] into H*W."""
    H, W = hw
    B, N, c = t.shape
    assert N == H * W and H % 2 == 0 and W % 2 == 0
    v = t.reshape(B, H // 2, 2, W // 2, 2, c).transpose(0, 1, 3, 2, 4, 5).reshape(B, H // 2, W // 2, 4, c)
    k = v.argmax(axis=3)                                   # first maximum wins, as in TF / cuDNN
    pooled = np.take_along_axis(v, k[:, :, :, None, :], axis=3)[:, :, :, 0, :]
    hh = np.arange(H // 2)[None, :, None, None] * 2 + k // 2
    ww = np.arange(W // 2)[None, None, :, None] * 2 + k % 2
    idx = hh * W + ww
    return pooled.reshape(B, -1, c), idx.reshape(B, -1, c)


def forward_pooled(X, Wphi, bphi, Wtheta, btheta, Wg, bg, Wo, bo, gamma, hw, return_cache=False):
    """As `forward`, with keys phi and values g max-pooled 2x2 / stride 2 over the [H, W] = hw grid."""
    phi = X @ Wphi + bphi
    theta = X @ Wtheta + btheta
    g = X @ Wg + bg
    phi_p, iphi = _pool2x2(phi, hw)
    g_p, ig = _pool2x2(g, hw)
    S = theta @ np.swapaxes(phi_p, 1, 2)                   # [B, N, N/4]
    S = S - S.max(axis=-1, keepdims=True)
    E = np.exp(S)
    P = E / E.sum(axis=-1, keepdims=True)
    A = P @ g_p
    O = A @ Wo + bo
    Y = X + gamma * O
    if return_cache:
        return Y, dict(phi=phi, theta=theta, g=g, phi_p=phi_p, g_p=g_p, iphi=iphi, ig=ig, P=P, A=A, O=O)
    return Y


def backward_pooled(dY, X, Wphi, bphi, Wtheta, btheta, Wg, bg, Wo, bo, gamma, hw):
    """Analytic gradients of `forward_pooled`: the pooled-key / pooled-value gradients are scattered to the argmax
    positions (everything else as in `backward`)."""
    Y, c = forward_pooled(X, Wphi, bphi, Wtheta, btheta, Wg, bg, Wo, bo, gamma, hw, return_cache=True)
    theta, phi_p, g_p, P, A, O = c["theta"], c["phi_p"], c["g_p"], c["P"], c["A"], c["O"]
    B, N, C = X.shape
    dgamma = np.sum(dY * O)
    dO = gamma * dY
    dWo = np.einsum("bnv,bnc->vc", A, dO)
    dbo = dO.sum(axis=(0, 1))
    dA = dO @ Wo.T
    dg_p = np.swapaxes(P, 1, 2) @ dA
    dP = dA @ np.swapaxes(g_p, 1, 2)
    dS = P * (dP - np.sum(dP * P, axis=-1, keepdims=True))
    dtheta = dS @ phi_p
    dphi_p = np.swapaxes(dS, 1, 2) @ theta

    def scatter(dp, idx, width):
        out = np.zeros((B, N, width), dtype=dp.dtype)
        bb = np.arange(B)[:, None, None]
        cc = np.arange(width)[None, None, :]
        np.add.at(out, (bb, idx, cc), dp)
        return out

    dphi = scatter(dphi_p, c["iphi"], phi_p.shape[-1])
    dg = scatter(dg_p, c["ig"], g_p.shape[-1])
    Xf = X.reshape(-1, C)
    return dict(
        dX=dY + dtheta @ Wtheta.T + dphi @ Wphi.T + dg @ Wg.T,
        dWphi=Xf.T @ dphi.reshape(-1, dphi.shape[-1]), dbphi=dphi.sum(axis=(0, 1)),
        dWtheta=Xf.T @ dtheta.reshape(-1, dtheta.shape[-1]), dbtheta=dtheta.sum(axis=(0, 1)),
        dWg=Xf.T @ dg.reshape(-1, dg.shape[-1]), dbg=dg.sum(axis=(0, 1)),
        dWo=dWo, dbo=dbo, dgamma=dgamma,
    )
