"""CPU check of the host logic behind the fp32 register-tiled convolution kernel (csrc/direct_fp32.cu): for the
transposed form with stride > 1 it is launched once per output parity class with only the taps that reach the class
(build_rt_classes).  The C++ builder itself is read back through the host-only hook npml_debug_rt_classes and its
tables are evaluated in NumPy float64: they must reproduce the oracle's Conv2D forward / dX and Deconv2D forward / dX
exactly, every output pixel must belong to exactly one class and every tap must be used.  The GPU parity of the kernel
is tests/test_gpu_parity.py (fp32 mode, miniature and full-size cases)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import npml  # noqa: E402
from oracle import conv_oracle as O  # noqa: E402


def classes_of(pkg, d, op):
    L = pkg._lib
    out = (C.c_int32 * 4096)()
    n = L.load().npml_debug_rt_classes(d, op, C.cast(out, C.c_void_p), 4096)
    assert n > 0
    v, i, res = list(out[:n]), 1, []
    for _ in range(v[0]):
        ntaps, y0, x0, so, sin, oh, ow = v[i:i + 7]
        i += 7
        taps = [tuple(v[i + 3 * j:i + 3 * j + 3]) for j in range(ntaps)]
        i += 3 * ntaps
        res.append(dict(y0=y0, x0=x0, so=so, sin=sin, oh=oh, ow=ow, taps=taps))
    assert i == n
    return res


def evaluate(classes, src, wtap, out_shape):
    """out[n, y0 + so yy, x0 + so xx, :] = sum_j src[n, yy sin + dy_j, xx sin + dx_j, :] @ wtap[tap_j] (zero outside)."""
    N, IH, IW, _ = src.shape
    out = np.zeros(out_shape)
    seen = np.zeros(out_shape[1:3], dtype=int)
    for c in classes:
        for yy in range(c["oh"]):
            for xx in range(c["ow"]):
                y, x = c["y0"] + c["so"] * yy, c["x0"] + c["so"] * xx
                seen[y, x] += 1
                for dy, dx, tp in c["taps"]:
                    iy, ix = yy * c["sin"] + dy, xx * c["sin"] + dx
                    if 0 <= iy < IH and 0 <= ix < IW:
                        out[:, y, x, :] += src[:, iy, ix, :] @ wtap[tp]
    assert (seen == 1).all()                       # the classes tile the output exactly once
    return out


CASES = [
    # kind, X shape, out_ch, kernel, pad, stride, dilation
    ("conv", (2, 9, 8, 4), 8, (3, 3), 1, 1, 0),
    ("conv", (2, 11, 10, 4), 8, (3, 3), 1, 2, 0),          # dX: 4 classes with 4 / 2 / 2 / 1 taps
    ("conv", (1, 12, 12, 4), 4, (4, 4), 1, 2, 0),          # dX: 4 taps per class
    ("conv", (1, 13, 11, 4), 4, (3, 2), (1, 0, 2, 1), 3, 0),   # stride 3, asymmetric pads: classes without a tap
    ("conv", (1, 14, 14, 4), 8, (3, 3), 0, 2, 2),          # cfg3's geometry (stride 2, dilation 2)
    ("conv", (1, 9, 9, 4), 4, (1, 1), 0, 2, 0),            # 1x1 stride 2: three of four dX classes are all zero
    ("deconv", (2, 5, 6, 4), 8, (4, 4), 1, 2, 0),          # cfg5's geometry
    ("deconv", (1, 6, 5, 4), 4, (3, 3), 1, 2, 0),
    ("deconv", (1, 4, 4, 8), 4, (2, 3), (1, 0), 3, 0),
    ("deconv", (2, 6, 6, 4), 4, (3, 3), 1, 1, 0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_rt_classes_reproduce_the_oracle(case):
    pkg = npml()
    U, L = pkg.utils, pkg._lib
    kind, xs, oc, ks, pad, s, dil = CASES[case]
    rng = np.random.default_rng(50 + case)
    X = rng.standard_normal(xs)
    W = rng.standard_normal(ks + (xs[3], oc))
    wt = W.reshape(ks[0] * ks[1], xs[3], oc)
    if kind == "conv":
        p4 = O.resolve_pad(xs, pad, ks, s, dil)
        oh, ow = O.calc_conv_out_dims(xs, W.shape, s, pad, dil)[1:3]
        d = U.make_desc(xs[0], xs[1], xs[2], xs[3], oc, ks[0], ks[1], s, dil + 1, tuple(p4), oh, ow, "fp32")
        Zr, _ = O.conv2d_layer_fwd(X, W, np.zeros((1, 1, 1, oc)), s, pad, dil)
        dY = rng.standard_normal(Zr.shape)
        dXr, _, _ = O.conv2d_layer_bwd(dY, X, Zr, W, s, pad, dil)
        fwd = classes_of(pkg, d, L.OP_CONV_FPROP)
        assert len(fwd) == 1 and len(fwd[0]["taps"]) == ks[0] * ks[1]
        np.testing.assert_allclose(evaluate(fwd, X, wt, Zr.shape), Zr, rtol=0, atol=1e-10)
        bwd = classes_of(pkg, d, L.OP_CONV_DGRAD)
        assert sorted(t[2] for c in bwd for t in c["taps"]) == sorted(set(t[2] for c in bwd for t in c["taps"]))
        np.testing.assert_allclose(evaluate(bwd, dY, wt.transpose(0, 2, 1), X.shape), dXr, rtol=0, atol=1e-10)
    else:
        p4 = O.resolve_pad(xs, pad, ks, s, 0) if not isinstance(pad, int) else (pad,) * 4
        if isinstance(pad, tuple) and len(pad) == 2:
            p4 = (pad[0], pad[0], pad[1], pad[1])
        Zr, _ = O.deconv2d_layer_fwd(X, W, np.zeros((1, 1, 1, oc)), s, pad)
        oh, ow = Zr.shape[1:3]
        d = U.make_desc(xs[0], xs[1], xs[2], xs[3], oc, ks[0], ks[1], s, 1, tuple(p4), oh, ow, "fp32")
        dY = rng.standard_normal(Zr.shape)
        dXr, _, _ = O.deconv2d_layer_bwd(dY, X, Zr, W, s, pad)
        fwd = classes_of(pkg, d, L.OP_DECONV_FPROP)
        assert sorted(t[2] for c in fwd for t in c["taps"]) == list(range(ks[0] * ks[1]))      # every tap exactly once
        np.testing.assert_allclose(evaluate(fwd, X, wt, Zr.shape), Zr, rtol=0, atol=1e-10)
        bwd = classes_of(pkg, d, L.OP_DECONV_DGRAD)
        assert len(bwd) == 1
        np.testing.assert_allclose(evaluate(bwd, dY, wt.transpose(0, 2, 1), X.shape), dXr, rtol=0, atol=1e-10)
