"""Oracle: spectral normalisation (numpy).  TEST INFRASTRUCTURE, see oracle/__init__.py.

Follows /root/reference/layers.py:
  l2normalize            layers.py:4-5
  _make_param (u, v)     layers.py:30-38
  update_uv              layers.py:50-68
Fixed readings: the normalised weight W/sigma IS what the wrapped
layer uses, u/v persist across calls, and they are constants in the backward.
"""
import numpy as np

EPS = 1e-12


def l2normalize(v, eps=EPS):
    # layers.py:4-5   v / (||v||_2 + eps)   (tf.norm of the whole tensor)
    return v / (np.sqrt(np.sum(v * v)) + eps)


def matricize(W):
    # layers.py:56   tf.reshape(W, [W.shape[-1], -1]) -- a RAW row-major
    # reinterpretation [R = last dim, K = numel / R], not a transpose.
    return np.reshape(W, (W.shape[-1], -1))


def make_param(W, rng):
    # layers.py:30-38   u ~ N(0,1) [1, R], v ~ N(0,1) [1, K], both l2-normalised
    R = W.shape[-1]
    K = W.size // R
    u = l2normalize(rng.standard_normal((1, R)).astype(W.dtype))
    v = l2normalize(rng.standard_normal((1, K)).astype(W.dtype))
    return u, v


def power_iteration(W, u, Ip=1, factor=None):
    """layers.py:50-68.  Returns (u_new [1,R], v_new [1,K], sigma, W_bar)."""
    if not Ip >= 1:
        # layers.py:17-18
        raise ValueError("The number of power iterations should be positive integer")
    Wm = matricize(W)
    v = None
    for _ in range(Ip):
        v = l2normalize(u @ Wm)            # layers.py:59
        u = l2normalize(v @ Wm.T)          # layers.py:60
    sigma = np.sum((u @ Wm) * v)           # layers.py:62
    if factor:
        sigma = sigma / factor             # layers.py:65-66
    W_bar = W / sigma                      # layers.py:68
    return u, v, sigma, W_bar


def backward(dW_bar, W_bar, u, v, sigma, factor=None):
    """Gradient of L w.r.t. W through W_bar = W / sigma(W), u and v constants.

    sigma = (u Wm v^T) / factor  =>  d sigma / d Wm = u^T v / factor
    dW = (dW_bar - (sum dW_bar * W_bar) * reshape(u^T v) / factor) / sigma
    """
    f = 1.0 if not factor else factor
    outer = (u.reshape(-1, 1) @ v.reshape(1, -1)).reshape(W_bar.shape)
    s = np.sum(dW_bar * W_bar)
    return (dW_bar - s * outer / f) / sigma
