"""
CPU oracle for the callers either side of the convolution hot path  --  TEST INFRASTRUCTURE ONLY.

float64 NumPy restatement of
  * calc_pad_dims_1D / pad1D / conv1D            numpy_ml/neural_nets/utils/utils.py:123-249, 668-717
  * Conv1D forward / _bwd                        numpy_ml/neural_nets/layers/layers.py:2603-2838
  * Pool2D forward / backward                    numpy_ml/neural_nets/layers/layers.py:3181-3350
  * Flatten forward / backward                   numpy_ml/neural_nets/layers/layers.py:859-966
  * SkipConnectionIdentityModule                 numpy_ml/neural_nets/modules/modules.py:360-612
Parity status: PINNED -- tests/golden/{conv1d_layer,pool2d,skip_identity_module}.npz were produced by the unmodified
reference (tests/golden/make_golden.py) and tests/test_oracle_golden.py checks every function here against them.
One case is outside the pinned domain: the reference's max-mode Pool2D.backward reads its windows from the UNPADDED
input (layers.py:3326) while the forward pools the padded one, so with pad > 0 it raises or masks a shifted window;
`pool2d_bwd` uses the padded windows of the forward there (identical to the reference whenever pad == 0).
Only tests/ may import this module; the product never does.
"""
import numpy as np

from . import conv_oracle as O


# --------------------------------------------------------------------------- #
#  1-D convolution                                                             #
# --------------------------------------------------------------------------- #
def calc_pad_dims_1D(X_shape, l_out, kernel_width, stride, dilation=0, causal=False):
    """(left, right) padding (utils.py:123-192): half the total on each side, one extra on the right if that is one
    short; everything on the left when `causal`."""
    for name, val, typ in (("X_shape", X_shape, tuple), ("l_out", l_out, int), ("kernel_width", kernel_width, int),
                           ("stride", stride, int)):
        if not isinstance(val, typ):
            raise ValueError("`{}` must be of type {}".format(name, typ.__name__))
    _, l_in, _ = X_shape
    _fw = O.eff_kernel(kernel_width, dilation)
    total = int(stride * (l_out - 1) + _fw - l_in)
    if causal:
        pw1, pw2 = total, 0
        assert int(1 + (l_in + total - _fw) / stride) == l_out
    else:
        pw = total // 2
        got = int(1 + (l_in + 2 * pw - _fw) / stride)
        pw1, pw2 = pw, pw
        if got == l_out - 1:
            pw2 = pw + 1
        elif got != l_out:
            raise AssertionError
    if pw1 < 0 or pw2 < 0:
        raise ValueError("Padding cannot be less than 0. Got: {}".format((pw1, pw2)))
    return (pw1, pw2)


def resolve_pad_1D(X_shape, pad, kernel_width=None, stride=None, dilation=0):
    """int | (left, right) | 'same' | 'causal' -> 2-tuple (utils.py:229-247)."""
    p = pad
    if isinstance(p, int):
        p = (p, p)
    if p in ("same", "causal") and kernel_width and stride:
        p = calc_pad_dims_1D(tuple(X_shape), X_shape[1], kernel_width, stride, dilation=dilation, causal=(p == "causal"))
    return tuple(int(v) for v in p)


def pad1D(X, pad, kernel_width=None, stride=None, dilation=0):
    p = resolve_pad_1D(X.shape, pad, kernel_width, stride, dilation)
    return np.pad(X, ((0, 0), (p[0], p[1]), (0, 0)), mode="constant"), p


def conv1D(X, W, stride, pad, dilation=0):
    """utils.py:668-717: a unit row dimension, then the 2-D path."""
    p = resolve_pad_1D(X.shape, pad, W.shape[0], stride, dilation)
    Z = O.conv2D(X[:, None], W[None], stride, (0, 0, p[0], p[1]), dilation)
    return Z[:, 0]


def conv1d_layer_fwd(X, W, b, stride, pad, dilation=0, act="identity", p0=0.0, p1=0.0):
    """Conv1D.forward (layers.py:2714-2752); b is (1, 1, out_ch)."""
    Z = conv1D(X, W, stride, pad, dilation) + b
    return O.act_fn(act, Z, p0, p1), Z


def conv1d_layer_bwd(dLdy, X, Z, W, stride, pad, dilation=0, act="identity", p0=0.0, p1=0.0):
    """Conv1D._bwd (layers.py:2804-2838): Conv2D._bwd on the row-expanded tensors."""
    p = resolve_pad_1D(X.shape, pad, W.shape[0], stride, dilation)
    dX, dW, db = O.conv2d_layer_bwd(dLdy[:, None], X[:, None], Z[:, None], W[None], stride, (0, 0, p[0], p[1]),
                                    dilation, act, p0, p1)
    return dX[:, 0], dW[0], db.reshape(1, 1, -1)


# --------------------------------------------------------------------------- #
#  Pool2D / Flatten                                                            #
# --------------------------------------------------------------------------- #
def pool2d_fwd(X, kernel_shape, stride=1, pad=0, mode="max"):
    """Pool2D.forward (layers.py:3234-3290): max / mean over every (fr, fc) window of the ZERO-padded input (the pad
    pixels take part in both the max and the mean)."""
    fr, fc = kernel_shape
    Xp, (pr1, pr2, pc1, pc2) = O.pad2D(X, pad, kernel_shape, stride)
    n, h, w, c = X.shape
    oh = int(np.floor(1 + (h + pr1 + pr2 - fr) / stride))
    ow = int(np.floor(1 + (w + pc1 + pc2 - fc) / stride))
    Y = np.zeros((n, oh, ow, c))
    fn = np.max if mode == "max" else np.mean
    for i in range(oh):
        for j in range(ow):
            Y[:, i, j, :] = fn(Xp[:, i * stride:i * stride + fr, j * stride:j * stride + fc, :], axis=(1, 2))
    return Y


def pool2d_bwd(dLdY, X, kernel_shape, stride=1, pad=0, mode="max"):
    """Pool2D.backward (layers.py:3292-3350).  average: every pixel of a window receives dy / (fr*fc); max: only the
    FIRST maximal pixel of the window in row-major order does (layers.py:3332-3338).  The gradient that lands on pad
    pixels is cropped."""
    fr, fc = kernel_shape
    Xp, (pr1, pr2, pc1, pc2) = O.pad2D(X, pad, kernel_shape, stride)
    n, oh, ow, c = dLdY.shape
    dXp = np.zeros_like(Xp)
    for i in range(oh):
        for j in range(ow):
            i0, j0 = i * stride, j * stride
            dy = dLdY[:, i, j, :]
            if mode == "max":
                win = Xp[:, i0:i0 + fr, j0:j0 + fc, :].reshape(n, fr * fc, c)
                first = np.argmax(win == win.max(axis=1, keepdims=True), axis=1)         # (n, c) row-major index
                mask = (np.arange(fr * fc)[None, :, None] == first[:, None, :]).reshape(n, fr, fc, c)
                dXp[:, i0:i0 + fr, j0:j0 + fc, :] += mask * dy[:, None, None, :]
            else:
                dXp[:, i0:i0 + fr, j0:j0 + fc, :] += dy[:, None, None, :] / (fr * fc)
    h, w = X.shape[1:3]
    return dXp[:, pr1:pr1 + h, pc1:pc1 + w, :]


def flatten_fwd(X, keep_dim="first"):
    """Flatten.forward (layers.py:912-936)."""
    if keep_dim == -1:
        return X.flatten().reshape(1, -1)
    return X.reshape(X.shape[0], -1) if keep_dim == "first" else X.reshape(-1, X.shape[-1])


# --------------------------------------------------------------------------- #
#  SkipConnectionIdentityModule (modules.py:360-612)                           #
# --------------------------------------------------------------------------- #
def skip_identity_module_fwd_bwd(X, dLdY, P, k1, k2, s1=1, s2=1, act="identity", p0=0.0, p1=0.0, momentum=0.9,
                                 epsilon=1e-5):
    """conv1('same', act) -> BN1 -> conv2('same', affine, out_ch = in_ch) -> BN2 -> Add([X, BN2]) -> act
    (modules.py:574-612).  P: W1,b1,W2,b2, g<x>,h<x>,rm<x>,rv<x> for x in (1, 2)."""
    o = {}
    o["conv1_out"], z1 = O.conv2d_layer_fwd(X, P["W1"], P["b1"], s1, "same", 0, act, p0, p1)
    o["batchnorm1_out"], o["rm1"], o["rv1"] = O.bn2d_fwd(o["conv1_out"], P["g1"], P["h1"], P["rm1"], P["rv1"], momentum, epsilon)
    o["conv2_out"], z2 = O.conv2d_layer_fwd(o["batchnorm1_out"], P["W2"], P["b2"], s2, "same", 0)
    o["batchnorm2_out"], o["rm2"], o["rv2"] = O.bn2d_fwd(o["conv2_out"], P["g2"], P["h2"], P["rm2"], P["rv2"], momentum, epsilon)
    o["Y"], total = O.add_fwd([X, o["batchnorm2_out"]], act, p0, p1)

    dX0, d_bn2 = O.add_bwd(dLdY, total, 2, act, p0, p1)
    o["dLdBn2"] = d_bn2
    o["dLdConv2"], o["dg2"], o["dh2"] = O.bn2d_bwd(d_bn2, o["conv2_out"], P["g2"], epsilon)
    o["dLdBn1"], o["dW2"], o["db2"] = O.conv2d_layer_bwd(o["dLdConv2"], o["batchnorm1_out"], z2, P["W2"], s2, "same", 0)
    o["dLdConv1"], o["dg1"], o["dh1"] = O.bn2d_bwd(o["dLdBn1"], o["conv1_out"], P["g1"], epsilon)
    dX1, o["dW1"], o["db1"] = O.conv2d_layer_bwd(o["dLdConv1"], X, z1, P["W1"], s1, "same", 0, act, p0, p1)
    o["dX"] = dX0 + dX1
    return o


# --------------------------------------------------------------------------- #
#  Multiply / FullyConnected / WavenetResidualModule / the VAE encoder chain    #
# --------------------------------------------------------------------------- #
def multiply_fwd(Xs, act="identity", p0=0.0, p1=0.0):
    """Multiply.forward (layers.py:800-822).  Returns (Y, product)."""
    prod = Xs[0].copy()
    for x in Xs[1:]:
        prod = prod * x
    return O.act_fn(act, prod, p0, p1), prod


def multiply_bwd(dLdY, Xs, prod, act="identity", p0=0.0, p1=0.0):
    """Multiply._bwd (layers.py:849-856): input i receives dLdY * act'(prod) * prod_{j != i} x_j."""
    g = dLdY * O.act_grad(act, prod, p0, p1)
    out = []
    for i in range(len(Xs)):
        gi = g.copy()
        for j, x in enumerate(Xs):
            if j != i:
                gi = gi * x
        out.append(gi)
    return out


def fc_fwd(X, W, b, act="identity", p0=0.0, p1=0.0):
    """FullyConnected._fwd (layers.py:2100-2106): Z = X @ W + b."""
    Z = X @ W + b
    return O.act_fn(act, Z, p0, p1), Z


def fc_bwd(dLdy, X, W, b, act="identity", p0=0.0, p1=0.0):
    """FullyConnected._bwd (layers.py:2144-2155)."""
    Z = X @ W + b
    dZ = dLdy * O.act_grad(act, Z, p0, p1)
    return dZ @ W.T, X.T @ dZ, dZ.sum(axis=0, keepdims=True)


def wavenet_module_fwd_bwd(X_main, X_skip, dY_skip, dY_main, P, dilation):
    """WavenetResidualModule.forward / backward (modules.py:283-357).  conv_dilation is a causal Conv1D of width 2
    (the constructor's kernel_width is stored but not used, modules.py:161-169), conv_1x1 a 'same' Conv1D of width 1.
    P: Wd, bd, W1, b1."""
    o = {}
    o["conv_dilation_out"], zd = conv1d_layer_fwd(X_main, P["Wd"], P["bd"], 1, "causal", dilation)
    o["tanh_out"], o["sigm_out"] = np.tanh(zd), 1.0 / (1.0 + np.exp(-zd))
    o["multiply_gate_out"], prod = multiply_fwd([o["tanh_out"], o["sigm_out"]])
    o["conv_1x1_out"], z1 = conv1d_layer_fwd(o["multiply_gate_out"], P["W1"], P["b1"], 1, "same", 0)
    o["Y_skip"] = X_skip + o["conv_1x1_out"]
    o["Y_main"] = X_main + o["conv_1x1_out"]
    d1 = dY_skip.copy()
    o["dX_skip"] = dY_skip.copy()
    dX_main = np.zeros_like(X_main)
    if dY_main is not None:
        dX_main = dY_main.copy()
        d1 = d1 + dY_main
    o["dLdConv_1x1"] = d1
    o["dLdMultiply"], o["dW1"], o["db1"] = conv1d_layer_bwd(d1, o["multiply_gate_out"], z1, P["W1"], 1, "same", 0)
    o["dLdTanh"], o["dLdSigmoid"] = multiply_bwd(o["dLdMultiply"], [o["tanh_out"], o["sigm_out"]], prod)
    dd = o["dLdTanh"] * O.act_grad("tanh", zd) + o["dLdSigmoid"] * O.act_grad("sigmoid", zd)
    o["dLdConv_dilation"] = dd
    back, o["dWd"], o["dbd"] = conv1d_layer_bwd(dd, X_main, zd, P["Wd"], 1, "causal", dilation)
    o["dX_main"] = dX_main + back
    return o
