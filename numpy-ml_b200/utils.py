"""Functional API of the convolution hot path, signature-compatible with
numpy_ml/neural_nets/utils/utils.py (pad2D, calc_pad_dims_2D, calc_conv_out_dims, im2col, col2im,
conv2D, deconv2D_naive, dilate) and of their 1-D forms (pad1D, calc_pad_dims_1D, conv1D).

All geometry (pad resolution incl. 'same', output dims, the reference's exceptions) is resolved here
on the host as plain ints; the arithmetic runs in libnpmlconv.so.  NumPy inputs give float64 NumPy
outputs (like the reference); device tensors give device tensors.  Outputs are contiguous, where
the reference returns transposed views.
"""
import numpy as np
import torch

from . import _lib as L


# --------------------------------------------------------------------------------------------------
#  geometry (host only)
# --------------------------------------------------------------------------------------------------
def calc_pad_dims_2D(X_shape, out_dim, kernel_shape, stride, dilation=0):
    """(top, bottom, left, right) padding so that a conv produces `out_dim`  (utils.py:51-120).

    Symmetric floor padding; one extra bottom/right row/col if that is exactly one short; an
    unattainable size is an AssertionError, a negative pad or a wrongly typed argument a ValueError."""
    for name, val, typ in (("X_shape", X_shape, tuple), ("out_dim", out_dim, tuple),
                           ("kernel_shape", kernel_shape, tuple), ("stride", stride, int)):
        if not isinstance(val, typ):
            raise ValueError("`{}` must be of type {}".format(name, typ.__name__))
    d = dilation
    fr, fc = kernel_shape
    out_rows, out_cols = out_dim
    _, in_rows, in_cols, _ = X_shape
    _fr, _fc = fr * (d + 1) - d, fc * (d + 1) - d
    pr = int((stride * (out_rows - 1) + _fr - in_rows) / 2)
    pc = int((stride * (out_cols - 1) + _fc - in_cols) / 2)
    got_rows = int(1 + (in_rows + 2 * pr - _fr) / stride)
    got_cols = int(1 + (in_cols + 2 * pc - _fc) / stride)
    pr1 = pr2 = pr
    if got_rows == out_rows - 1:
        pr2 = pr + 1
    elif got_rows != out_rows:
        raise AssertionError
    pc1 = pc2 = pc
    if got_cols == out_cols - 1:
        pc2 = pc + 1
    elif got_cols != out_cols:
        raise AssertionError
    if any(np.array([pr1, pr2, pc1, pc2]) < 0):
        raise ValueError("Padding cannot be less than 0. Got: {}".format((pr1, pr2, pc1, pc2)))
    return (pr1, pr2, pc1, pc2)


def resolve_pad_2D(X_shape, pad, kernel_shape=None, stride=None, dilation=0):
    """int | (r, c) | (top, bottom, left, right) | 'same' -> 4-tuple, as pad2D does (utils.py:285-305)."""
    p = pad
    if isinstance(p, int):
        p = (p, p, p, p)
    if isinstance(p, tuple) and len(p) == 2:
        p = (p[0], p[0], p[1], p[1])
    if p == "same" and kernel_shape and stride is not None:
        p = calc_pad_dims_2D(tuple(X_shape), tuple(X_shape[1:3]), tuple(kernel_shape), stride, dilation=dilation)
    if not (isinstance(p, tuple) and len(p) == 4):
        raise ValueError("Unrecognized pad: {}".format(pad))
    return tuple(int(v) for v in p)


def calc_pad_dims_1D(X_shape, l_out, kernel_width, stride, dilation=0, causal=False):
    """(left, right) padding so that a 1-D conv produces length `l_out` (utils.py:123-192): half of the total on each
    side plus one on the right if that is exactly one short; all of it on the left for a causal convolution."""
    for name, val, typ in (("X_shape", X_shape, tuple), ("l_out", l_out, int), ("kernel_width", kernel_width, int),
                           ("stride", stride, int)):
        if not isinstance(val, typ):
            raise ValueError("`{}` must be of type {}".format(name, typ.__name__))
    _, l_in, _ = X_shape
    _fw = kernel_width * (dilation + 1) - dilation
    total_pad = int(stride * (l_out - 1) + _fw - l_in)
    if causal:
        pw1, pw2 = total_pad, 0
        assert int(1 + (l_in + total_pad - _fw) / stride) == l_out
    else:
        pw1 = pw2 = total_pad // 2
        got = int(1 + (l_in + 2 * pw1 - _fw) / stride)
        if got == l_out - 1:
            pw2 = pw1 + 1
        elif got != l_out:
            raise AssertionError
    if pw1 < 0 or pw2 < 0:
        raise ValueError("Padding cannot be less than 0. Got: {}".format((pw1, pw2)))
    return (pw1, pw2)


def resolve_pad_1D(X_shape, pad, kernel_width=None, stride=None, dilation=0):
    """int | (left, right) | 'same' | 'causal' -> 2-tuple, as pad1D does (utils.py:229-247)."""
    p = pad
    if isinstance(p, int):
        p = (p, p)
    if p in ("same", "causal") and kernel_width and stride:
        p = calc_pad_dims_1D(tuple(X_shape), int(X_shape[1]), kernel_width, stride, dilation=dilation,
                             causal=(p == "causal"))
    if not (isinstance(p, tuple) and len(p) == 2):
        raise ValueError("Unrecognized pad: {}".format(pad))
    return tuple(int(v) for v in p)


def calc_conv_out_dims(X_shape, W_shape, stride=1, pad=0, dilation=0):
    """Output volume shape of a 1D / 2D convolution (utils.py:380-439)."""
    if len(X_shape) == 3:
        pw1, pw2 = resolve_pad_1D(X_shape, pad)
        fw, _, out_ch = W_shape
        return (X_shape[0], (X_shape[1] + pw1 + pw2 - (fw * (dilation + 1) - dilation)) // stride + 1, out_ch)
    if len(X_shape) != 4:
        raise ValueError("Unrecognized number of input dims: {}".format(len(X_shape)))
    pr1, pr2, pc1, pc2 = resolve_pad_2D(X_shape, pad)
    fr, fc, _, out_ch = W_shape
    n_ex, in_rows, in_cols, _ = X_shape
    d = dilation
    _fr, _fc = fr * (d + 1) - d, fc * (d + 1) - d
    return (n_ex, (in_rows + pr1 + pr2 - _fr) // stride + 1, (in_cols + pc1 + pc2 - _fc) // stride + 1, out_ch)


def conv_out_hw(in_rows, in_cols, fr, fc, p4, stride, dilation):
    """(out_rows, out_cols) as conv2D computes them (utils.py:653-657); ValueError when empty (463-467)."""
    d = dilation
    _fr, _fc = fr * (d + 1) - d, fc * (d + 1) - d
    out_rows = int((in_rows + p4[0] + p4[1] - _fr) / stride + 1)
    out_cols = int((in_cols + p4[2] + p4[3] - _fc) / stride + 1)
    if (in_rows + p4[0] + p4[1] - _fr) < 0 or (in_cols + p4[2] + p4[3] - _fc) < 0 or out_rows <= 0 or out_cols <= 0:
        raise ValueError("Dimension mismatch during convolution: out_rows = {}, out_cols = {}".format(out_rows,
                                                                                                 out_cols))
    return out_rows, out_cols


def deconv_geometry(in_rows, in_cols, fr, fc, pad, stride):
    """Effective (top, bottom, left, right) crop and output size of deconv2D_naive (utils.py:764-794).

    The reference zero-stuffs X, pads it by `pad`, pads again so that a stride-1 conv with the
    flipped kernel yields the full transposed-conv size, and crops `pad` off the result.  Written
    out, Z[y] receives X[iy]*W[r] at y = iy*stride + r - e with e = (fr-1) - int((2*(fr-1) - p1 - p2)/2);
    for symmetric pads e = p (the conv_transpose2d definition)."""
    hd, wd = stride * (in_rows - 1) + 1, stride * (in_cols - 1) + 1
    p4 = resolve_pad_2D((1, hd, wd, 1), pad, (fr, fc), 1, 0)
    pr1, pr2, pc1, pc2 = p4
    out_rows, out_cols = hd - 1 + fr, wd - 1 + fc                 # before the crop
    extra = calc_pad_dims_2D((1, hd + pr1 + pr2, wd + pc1 + pc2, 1), (out_rows, out_cols), (fr, fc), 1, 0)
    er, ec = (fr - 1) - extra[0], (fc - 1) - extra[2]
    oh, ow = out_rows - pr1 - pr2, out_cols - pc1 - pc2
    if oh <= 0 or ow <= 0 or er < 0 or ec < 0:
        raise ValueError("Dimension mismatch during deconvolution: out_rows = {}, out_cols = {}".format(oh, ow))
    return (er, pr1 + pr2 - er, ec, pc1 + pc2 - ec), oh, ow


def make_desc(N, H, W, C, K, fr, fc, stride, dil, p4, OH, OW, mode="fp32", act=None, accumulate=0):
    d = L.ConvDesc()
    d.N, d.H, d.W, d.C, d.K, d.fr, d.fc = N, H, W, C, K, fr, fc
    d.stride, d.dil = stride, dil
    d.pr1, d.pr2, d.pc1, d.pc2 = p4
    d.OH, d.OW = OH, OW
    d.mode = L.MODES[mode]
    d.act = act.code if act is not None else 0
    d.act_p0 = act.p0 if act is not None else 0.0
    d.act_p1 = act.p1 if act is not None else 0.0
    d.accumulate = accumulate
    return d


def _is_np(x):
    return isinstance(x, np.ndarray)


def _wrap(t, as_np):
    return L.to_numpy(t) if as_np else t


# --------------------------------------------------------------------------------------------------
#  tensor functions
# --------------------------------------------------------------------------------------------------
def pad2D(X, pad, kernel_shape=None, stride=None, dilation=0):
    """Zero-pad rows/cols of an NHWC volume; returns (X_pad, p4)  (utils.py:252-306).
    API parity only: the convolution kernels never materialise the padded volume."""
    as_np = _is_np(X)
    p = resolve_pad_2D(X.shape, pad, kernel_shape, stride, dilation)
    Xd = L.to_device(X, torch.float32) if as_np else L.to_device(X, X.dtype)
    n, h, w, c = Xd.shape
    out = torch.zeros((n, h + p[0] + p[1], w + p[2] + p[3], c), dtype=Xd.dtype, device=Xd.device)
    out[:, p[0]:p[0] + h, p[2]:p[2] + w, :].copy_(Xd)
    return _wrap(out, as_np), p


def pad1D(X, pad, kernel_width=None, stride=None, dilation=0):
    """Zero-pad the length axis of an (n_ex, l_in, in_ch) volume; returns (X_pad, (left, right)) (utils.py:195-249).
    API parity only: Conv1D never materialises the padded volume."""
    as_np = _is_np(X)
    p = resolve_pad_1D(X.shape, pad, kernel_width, stride, dilation)
    Xd = L.to_device(X, torch.float32) if as_np else L.to_device(X, X.dtype)
    n, l, c = Xd.shape
    out = torch.zeros((n, l + p[0] + p[1], c), dtype=Xd.dtype, device=Xd.device)
    out[:, p[0]:p[0] + l, :].copy_(Xd)
    return _wrap(out, as_np), p


def dilate(X, d):
    """Insert `d` zero rows/cols between adjacent pixels (utils.py:309-344).  API parity only."""
    as_np = _is_np(X)
    Xd = L.to_device(X, torch.float32) if as_np else L.to_device(X, X.dtype)
    n, h, w, c = Xd.shape
    out = torch.zeros((n, h + d * (h - 1), w + d * (w - 1), c), dtype=Xd.dtype, device=Xd.device)
    out[:, ::d + 1, ::d + 1, :].copy_(Xd)
    return _wrap(out, as_np)


def im2col(X, W_shape, pad, stride, dilation=0):
    """(X_col of shape (fr*fc*in_ch, out_rows*out_cols*n_ex), p4)  (utils.py:486-543).
    Exported for API parity (fp32); conv2D and the layers use implicit GEMM instead."""
    as_np = _is_np(X)
    fr, fc, _, out_ch = W_shape
    Xd = L.to_device(X, torch.float32)
    n, h, w, c = Xd.shape
    p = resolve_pad_2D(Xd.shape, pad, (fr, fc), stride, dilation)
    oh, ow = conv_out_hw(h, w, fr, fc, p, stride, dilation)
    d = make_desc(n, h, w, c, out_ch, fr, fc, stride, dilation + 1, p, oh, ow)
    out = torch.empty((fr * fc * c, oh * ow * n), dtype=torch.float32, device=Xd.device)
    L.check(L.load().npml_im2col(d, L.ptr(Xd), L.ptr(out), L.stream()), "npml_im2col")
    return _wrap(out, as_np), p


def col2im(X_col, X_shape, W_shape, pad, stride, dilation=0):
    """Adjoint of im2col; returns (n_ex, in_ch, in_rows, in_cols) like the reference (utils.py:546-595)."""
    if not (isinstance(pad, tuple) and len(pad) == 4):
        raise TypeError("pad must be a 4-tuple, but got: {}".format(pad))
    as_np = _is_np(X_col)
    fr, fc, _, out_ch = W_shape
    n, h, w, c = X_shape
    oh, ow = conv_out_hw(h, w, fr, fc, pad, stride, dilation)
    Xc = L.to_device(X_col, torch.float32)
    d = make_desc(n, h, w, c, out_ch, fr, fc, stride, dilation + 1, pad, oh, ow)
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=Xc.device)
    L.check(L.load().npml_col2im(d, L.ptr(Xc), L.ptr(out), L.stream()), "npml_col2im")
    return _wrap(out, as_np)


def _weights(W):
    return L.to_device(W, torch.float32)


def conv2D(X, W, stride, pad, dilation=0, mode="fp32"):
    """2D cross-correlation of NHWC `X` with (fr, fc, in_ch, out_ch) `W`, no bias (utils.py:603-665)."""
    as_np = _is_np(X)
    Xd = L.to_device(X, L.torch_dtype(mode))
    Wd = _weights(W)
    fr, fc, in_ch, out_ch = Wd.shape
    n, h, w, c = Xd.shape
    p = resolve_pad_2D(Xd.shape, pad, (fr, fc), stride, dilation)
    oh, ow = conv_out_hw(h, w, fr, fc, p, stride, dilation)
    d = make_desc(n, h, w, c, out_ch, fr, fc, stride, dilation + 1, p, oh, ow, mode)
    Z = torch.empty((n, oh, ow, out_ch), dtype=Xd.dtype, device=Xd.device)
    lib = L.load()
    nws = lib.npml_workspace_bytes(d, L.OP_CONV_FPROP)
    ws = L.workspace(nws)
    L.check(lib.npml_conv2d_fprop(d, L.ptr(Xd), L.ptr(Wd), None, L.ptr(Z), None, L.ptr(ws), ws.numel(), L.stream()),
            "npml_conv2d_fprop")
    return _wrap(Z, as_np)


def conv1D(X, W, stride, pad, dilation=0, mode="fp32"):
    """1D cross-correlation of (n_ex, l_in, in_ch) `X` with (kernel_width, in_ch, out_ch) `W` (utils.py:668-717): the
    same kernels as conv2D on a volume with one row."""
    as_np = _is_np(X)
    Xd = L.to_device(X, L.torch_dtype(mode))
    Wd = _weights(W)
    p = resolve_pad_1D(Xd.shape, pad, Wd.shape[0], stride, dilation)
    Z = conv2D(Xd.unsqueeze(1), Wd.unsqueeze(0), stride, (0, 0, p[0], p[1]), dilation, mode)
    return _wrap(Z.squeeze(1), as_np)


def conv2D_naive(X, W, stride, pad, dilation=0):
    """The reference keeps a 4-loop version as a test oracle (utils.py:797-856); here it is the same
    kernel in fp32 exactness mode."""
    return conv2D(X, W, stride, pad, dilation, mode="fp32")


def deconv2D_naive(X, W, stride, pad, dilation=0, mode="fp32"):
    """Transposed convolution (utils.py:720-794) without zero-stuffing: every output pixel gathers only
    the taps that reach it."""
    if dilation != 0:
        raise NotImplementedError("deconv2D_naive: the layers always pass dilation=0 (layers.py:3469)")
    as_np = _is_np(X)
    Xd = L.to_device(X, L.torch_dtype(mode))
    Wd = _weights(W)
    fr, fc, in_ch, out_ch = Wd.shape
    n, h, w, c = Xd.shape
    p, oh, ow = deconv_geometry(h, w, fr, fc, pad, stride)
    d = make_desc(n, h, w, c, out_ch, fr, fc, stride, 1, p, oh, ow, mode)
    Z = torch.empty((n, oh, ow, out_ch), dtype=Xd.dtype, device=Xd.device)
    lib = L.load()
    ws = L.workspace(lib.npml_workspace_bytes(d, L.OP_DECONV_FPROP))
    L.check(lib.npml_deconv2d_fprop(d, L.ptr(Xd), L.ptr(Wd), None, L.ptr(Z), None, L.ptr(ws), ws.numel(),
                                    L.stream()), "npml_deconv2d_fprop")
    return _wrap(Z, as_np)


# --------------------------------------------------------------------------------------------------
#  weight initialisation (host; runs once per layer)
# --------------------------------------------------------------------------------------------------
def calc_fan(weight_shape):
    """fan_in / fan_out of a weight tensor (utils.py:352-377)."""
    if len(weight_shape) == 2:
        fan_in, fan_out = weight_shape
    elif len(weight_shape) in [3, 4]:
        in_ch, out_ch = weight_shape[-2:]
        kernel_size = np.prod(weight_shape[:-2])
        fan_in, fan_out = in_ch * kernel_size, out_ch * kernel_size
    else:
        raise ValueError("Unrecognized weight dimension: {}".format(weight_shape))
    return fan_in, fan_out


def truncated_normal(mean, std, out_shape):
    """Normal draws re-sampled until within two std (utils.py:1002-1030); consumes np.random like the reference."""
    samples = np.random.normal(loc=mean, scale=std, size=out_shape)
    reject = np.logical_or(samples >= mean + 2 * std, samples <= mean - 2 * std)
    while any(reject.flatten()):
        samples[reject] = np.random.normal(loc=mean, scale=std, size=reject.sum())
        reject = np.logical_or(samples >= mean + 2 * std, samples <= mean - 2 * std)
    return samples


def he_uniform(weight_shape):
    fan_in, _ = calc_fan(weight_shape)
    b = np.sqrt(6 / fan_in)
    return np.random.uniform(-b, b, size=weight_shape)


def he_normal(weight_shape):
    fan_in, _ = calc_fan(weight_shape)
    return truncated_normal(0, np.sqrt(2 / fan_in), weight_shape)


def glorot_uniform(weight_shape, gain=1.0):
    fan_in, fan_out = calc_fan(weight_shape)
    b = gain * np.sqrt(6 / (fan_in + fan_out))
    return np.random.uniform(-b, b, size=weight_shape)


def glorot_normal(weight_shape, gain=1.0):
    fan_in, fan_out = calc_fan(weight_shape)
    return truncated_normal(0, gain * np.sqrt(2 / (fan_in + fan_out)), weight_shape)
