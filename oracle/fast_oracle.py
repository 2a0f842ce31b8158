"""TEST INFRASTRUCTURE — tap-wise float64 evaluation of the convolution sums, for FULL-SIZE parity checks.

`oracle/conv_oracle.py` restates the reference literally (index-gather im2col, one dgemm, `np.add.at`
col2im) and is pinned to the reference's outputs; at the BASELINE.json sizes it needs 35-55 s and up to
11 GB per forward+backward, almost all of it in the gather / scatter.  The functions here compute THE SAME
sums (a5, a11, a13) one filter tap at a time, as strided views of the zero-padded volume times
a (C x K) slice of the weights, so a whole benchmark configuration takes seconds:

    Z [n,oh,ow,k] = sum_{r,t,c} Xpad[n, oh*s + r*D, ow*s + t*D, c] * W[r,t,c,k]        utils/utils.py:603-665
    dW[r,t,c,k]   = sum_{n,oh,ow} Xpad[n, oh*s + r*D, ow*s + t*D, c] * dZ[n,oh,ow,k]    layers/layers.py:3104-3110
    dXpad[n, oh*s + r*D, ow*s + t*D, c] += sum_k dZ[n,oh,ow,k] * W[r,t,c,k]             layers/layers.py:3113-3114

They are not a second opinion on WHAT is computed: `tests/test_oracle_golden.py` pins every function here
to `conv_oracle` (and through it to the reference's golden vectors) at <= 1e-12 on all miniature cases; the
only difference is float64 summation order.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product never does.
"""
import numpy as np

from . import conv_oracle as O


def _geom(X_shape, W_shape, stride, pad, dilation):
    p = O.resolve_pad(X_shape, pad, W_shape[:2], stride, dilation)
    fr, fc = W_shape[:2]
    _, h, w, _ = X_shape
    oh = (h + p[0] + p[1] - O.eff_kernel(fr, dilation)) // stride + 1
    ow = (w + p[2] + p[3] - O.eff_kernel(fc, dilation)) // stride + 1
    if oh <= 0 or ow <= 0:
        raise ValueError("Dimension mismatch during convolution: out_rows = {}, out_cols = {}".format(oh, ow))
    return p, oh, ow


def _padded(X, p):
    return np.pad(X, ((0, 0), (p[0], p[1]), (p[2], p[3]), (0, 0)))


def _tap_view(Xp, r, t, D, s, oh, ow):
    return Xp[:, r * D: r * D + (oh - 1) * s + 1: s, t * D: t * D + (ow - 1) * s + 1: s, :]


def conv2d(X, W, stride, pad, dilation=0):
    """conv2D (utils/utils.py:603-665) as fr*fc strided-view GEMMs."""
    p, oh, ow = _geom(X.shape, W.shape, stride, pad, dilation)
    Xp, D = _padded(np.asarray(X, dtype=np.float64), p), dilation + 1
    fr, fc, _, k = W.shape
    Z = np.zeros((X.shape[0], oh, ow, k))
    for r in range(fr):
        for t in range(fc):
            Z += np.ascontiguousarray(_tap_view(Xp, r, t, D, stride, oh, ow)) @ W[r, t]
    return Z


def conv2d_layer_fwd(X, W, b, stride, pad, dilation=0, act="identity", p0=0.0, p1=0.0):
    """Conv2D.forward (layers/layers.py:3006-3046)."""
    Z = conv2d(X, W, stride, pad, dilation) + b
    return O.act_fn(act, Z, p0, p1), Z


def conv2d_layer_bwd(dLdy, X, Z, W, stride, pad, dilation=0, act="identity", p0=0.0, p1=0.0):
    """Conv2D._bwd (layers/layers.py:3093-3116): (dX, dW, db)."""
    p, oh, ow = _geom(X.shape, W.shape, stride, pad, dilation)
    Xp, D = _padded(np.asarray(X, dtype=np.float64), p), dilation + 1
    fr, fc, c, k = W.shape
    dZ = dLdy * O.act_grad(act, Z, p0, p1)
    dZ2 = np.ascontiguousarray(dZ).reshape(-1, k)
    db = dZ2.sum(axis=0).reshape(1, 1, 1, -1)
    dW = np.zeros(W.shape)
    dXp = np.zeros(Xp.shape)
    for r in range(fr):
        for t in range(fc):
            v = np.ascontiguousarray(_tap_view(Xp, r, t, D, stride, oh, ow)).reshape(-1, c)
            dW[r, t] = v.T @ dZ2
            _tap_view(dXp, r, t, D, stride, oh, ow)[...] += (dZ2 @ W[r, t].T).reshape(dZ.shape[:3] + (c,))
    h, w = X.shape[1:3]
    return dXp[:, p[0]: p[0] + h, p[2]: p[2] + w, :], dW, db


def _sym_pad(pad, X_shape, W_shape):
    """Deconv2D's pinned domain: int / 2-tuple symmetric pads, and 'same' at stride 1 with
    an odd kernel."""
    if isinstance(pad, int):
        return pad, pad
    if isinstance(pad, tuple) and len(pad) == 2:
        return pad
    if pad == "same":
        fr, fc = W_shape[:2]
        assert fr % 2 == 1 and fc % 2 == 1, "'same' with an even kernel is outside the pinned domain"
        return (fr - 1) // 2, (fc - 1) // 2
    raise ValueError("fast_oracle: Deconv2D pad {} is outside the pinned (symmetric) domain".format(pad))


def deconv2d(X, W, stride, pad):
    """deconv2D_naive on its pinned domain (utils/utils.py:720-794): every input pixel
    scatters X[n,iy,ix,:] @ W[r,t] to (iy*s + r, ix*s + t); the symmetric pad is cropped."""
    pr, pc = _sym_pad(pad, X.shape, W.shape)
    n, h, w, _ = X.shape
    fr, fc, _, k = W.shape
    s = stride
    full = np.zeros((n, s * (h - 1) + fr, s * (w - 1) + fc, k))
    X = np.ascontiguousarray(X, dtype=np.float64)
    for r in range(fr):
        for t in range(fc):
            full[:, r: r + (h - 1) * s + 1: s, t: t + (w - 1) * s + 1: s, :] += X @ W[r, t]
    return full[:, pr: full.shape[1] - pr, pc: full.shape[2] - pc, :]


def deconv2d_layer_fwd(X, W, b, stride, pad, act="identity", p0=0.0, p1=0.0):
    """Deconv2D.forward (layers/layers.py:3438-3478)."""
    Z = deconv2d(X, W, stride, pad) + b
    return O.act_fn(act, Z, p0, p1), Z


def deconv2d_layer_bwd(dLdY, X, Z, W, stride, pad, act="identity", p0=0.0, p1=0.0):
    """Deconv2D._bwd (layers/layers.py:3521-3568) in the closed form SURVEY.md 8 a13 measured against brute
    force: dX is a strided forward convolution of dZ, dW[r,t] = X^T (dZ shifted by the tap)."""
    pr, pc = _sym_pad(pad, X.shape, W.shape)
    n, h, w, c = X.shape
    fr, fc, _, k = W.shape
    s = stride
    dZ = dLdY * O.act_grad(act, Z, p0, p1)
    db = dZ.reshape(-1, k).sum(axis=0).reshape(1, 1, 1, -1)
    dZf = np.pad(dZ, ((0, 0), (pr, pr), (pc, pc), (0, 0)))
    X2 = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, c)
    dX = np.zeros(X.shape)
    dW = np.zeros(W.shape)
    for r in range(fr):
        for t in range(fc):
            v = np.ascontiguousarray(dZf[:, r: r + (h - 1) * s + 1: s, t: t + (w - 1) * s + 1: s, :]).reshape(-1, k)
            dW[r, t] = X2.T @ v
            dX += (v @ W[r, t].T).reshape(X.shape)
    return dX, dW, db


class _tapwise(object):
    """conv_oracle's module wiring with the tap-wise convolutions above (swapped in for the duration of a call)."""

    def __enter__(self):
        self.saved = O.conv2d_layer_fwd, O.conv2d_layer_bwd
        O.conv2d_layer_fwd, O.conv2d_layer_bwd = conv2d_layer_fwd, conv2d_layer_bwd

    def __exit__(self, *exc):
        O.conv2d_layer_fwd, O.conv2d_layer_bwd = self.saved


def skip_conv_module_fwd(*args, **kw):
    """SkipConnectionConvModule forward (modules/modules.py:905-937)."""
    with _tapwise():
        return O.skip_conv_module_fwd(*args, **kw)


def skip_conv_module_bwd(*args, **kw):
    """SkipConnectionConvModule backward (modules/modules.py:939-984)."""
    with _tapwise():
        return O.skip_conv_module_bwd(*args, **kw)


def skip_conv_module_fwd_bwd(*args, **kw):
    with _tapwise():
        return O.skip_conv_module_fwd_bwd(*args, **kw)
