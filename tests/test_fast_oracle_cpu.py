"""Pins oracle/fast_oracle.py (tap-wise float64 sums used for the FULL-SIZE GPU parity checks) to the literal
restatement oracle/conv_oracle.py -- which tests/test_oracle_golden.py pins to the reference's own outputs -- and
to the reference's golden vectors directly.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, to_pad, rel_err
from oracle import conv_oracle as O
from oracle import fast_oracle as F

TOL = 1e-12

CONV = [
    # xs, out_ch, kernel, pad, stride, dilation, act
    ((3, 9, 8, 5), 7, (3, 3), "same", 1, 0, "identity"),
    ((2, 28, 28, 6), 4, (3, 3), 0, 2, 2, "tanh"),
    ((2, 14, 14, 3), 5, (3, 3), "same", 2, 2, "relu"),          # (16,17,16,17)-style asymmetric 'same'
    ((2, 9, 11, 4), 3, (3, 2), (1, 2), 1, 0, "sigmoid"),
    ((2, 10, 10, 2), 3, (2, 2), (1, 0, 2, 1), 3, 0, "identity"),  # stride leftovers, 4-tuple pad
    ((1, 6, 6, 2), 2, (1, 1), 0, 1, 0, "leaky_relu"),
    ((2, 7, 7, 3), 3, (5, 5), 2, 1, 1, "gelu"),
]
DECONV = [
    ((2, 6, 6, 5), 3, (4, 4), 1, 2, "identity"),
    ((3, 6, 5, 4), 4, (3, 3), 1, 1, "relu"),
    ((2, 5, 5, 3), 4, (2, 3), (1, 0), 3, "tanh"),
    ((2, 7, 7, 2), 3, (3, 3), "same", 1, "identity"),
    ((2, 4, 4, 3), 2, (5, 5), 2, 2, "sigmoid"),
]


def _mk(rng, xs, oc, ks):
    X = rng.standard_normal(xs)
    W = rng.standard_normal(ks + (xs[3], oc)) / np.sqrt(ks[0] * ks[1] * xs[3])
    b = 0.1 * rng.standard_normal((1, 1, 1, oc))
    return X, W, b


@pytest.mark.parametrize("case", range(len(CONV)))
def test_conv_taps_equal_im2col_restatement(case):
    xs, oc, ks, pad, s, d, act = CONV[case]
    rng = np.random.default_rng(case)
    X, W, b = _mk(rng, xs, oc, ks)
    p0 = 0.3 if act == "leaky_relu" else 0.0
    Y, Z = O.conv2d_layer_fwd(X, W, b, s, pad, d, act, p0)
    Yf, Zf = F.conv2d_layer_fwd(X, W, b, s, pad, d, act, p0)
    assert rel_err(Yf, Y) <= TOL and rel_err(Zf, Z) <= TOL
    dY = rng.standard_normal(Y.shape)
    for a, r in zip(F.conv2d_layer_bwd(dY, X, Z, W, s, pad, d, act, p0), O.conv2d_layer_bwd(dY, X, Z, W, s, pad, d, act, p0)):
        assert a.shape == r.shape and rel_err(a, r) <= TOL


@pytest.mark.parametrize("case", range(len(DECONV)))
def test_deconv_taps_equal_zero_stuffed_restatement(case):
    xs, oc, ks, pad, s, act = DECONV[case]
    rng = np.random.default_rng(50 + case)
    X, W, b = _mk(rng, xs, oc, ks)
    Y, Z = O.deconv2d_layer_fwd(X, W, b, s, pad, act)
    Yf, Zf = F.deconv2d_layer_fwd(X, W, b, s, pad, act)
    assert Yf.shape == Y.shape and rel_err(Yf, Y) <= TOL and rel_err(Zf, Z) <= TOL
    dY = rng.standard_normal(Y.shape)
    for a, r in zip(F.deconv2d_layer_bwd(dY, X, Z, W, s, pad, act), O.deconv2d_layer_bwd(dY, X, Z, W, s, pad, act)):
        assert a.shape == r.shape and rel_err(a, r) <= TOL


def test_fast_oracle_against_reference_golden_layers():
    """Straight against the fixtures the unmodified reference produced (tests/golden/make_golden.py)."""
    for name, fwd, bwd in (("conv2d_layer.npz", F.conv2d_layer_fwd, F.conv2d_layer_bwd),
                           ("deconv2d_layer.npz", F.deconv2d_layer_fwd, F.deconv2d_layer_bwd)):
        z, cases = load_golden(name)
        n = 0
        for c in cases:
            i, pad = c["id"], to_pad(c["pad"])
            if name.startswith("deconv") and not (isinstance(pad, int) or (isinstance(pad, tuple) and len(pad) == 2)
                                                  or (pad == "same" and c["stride"] == 1 and c["kernel_shape"][0] % 2)):
                continue                                    # outside the pinned symmetric domain
            args = (c["stride"], pad) + ((c.get("dilation", 0),) if name.startswith("conv") else ())
            Y, Z = fwd(z[i + "/X"], z[i + "/W"], z[i + "/b"], *args, c["act"], c["p0"], c["p1"])
            assert rel_err(Y, z[i + "/Y"]) <= TOL, i
            dX, dW, db = bwd(z[i + "/dLdy"], z[i + "/X"], z[i + "/Z"], z[i + "/W"], *args, c["act"], c["p0"], c["p1"])
            if not c["two_calls"]:
                assert rel_err(dW, z[i + "/dW"]) <= TOL and rel_err(db, z[i + "/db"]) <= TOL, i
            assert rel_err(dX, z[i + "/dX"]) <= TOL, i
            n += 1
        assert n >= 5


def test_fast_skip_module_equals_restatement():
    rng = np.random.default_rng(9)
    X = rng.standard_normal((3, 8, 8, 4))
    P = {}
    for t, (k, ci, co) in (("1", ((3, 3), 4, 6)), ("2", ((3, 3), 6, 6)), ("s", ((1, 1), 4, 6))):
        P["W" + t] = rng.standard_normal(k + (ci, co)) / 5
        P["b" + t] = 0.1 * rng.standard_normal((1, 1, 1, co))
        P["g" + t], P["h" + t] = rng.uniform(0.5, 1.5, co), 0.1 * rng.standard_normal(co)
        P["rm" + t], P["rv" + t] = np.zeros(co), np.ones(co)
    dY = rng.standard_normal((3, 8, 8, 6))
    a = F.skip_conv_module_fwd_bwd(X, dY, P, (3, 3), (3, 3), (1, 1), 1, 1, act="relu")
    r = O.skip_conv_module_fwd_bwd(X, dY, P, (3, 3), (3, 3), (1, 1), 1, 1, act="relu")
    assert O.conv2d_layer_fwd.__module__ == "oracle.conv_oracle"        # the temporary swap was undone
    for k in r:
        if k in ("db2", "dbs"):              # an affine conv in front of a BatchNorm: the true gradient is 0 (pure cancellation)
            assert np.abs(a[k]).max() <= 1e-12 and np.abs(r[k]).max() <= 1e-12, k
        else:
            assert rel_err(a[k], r[k]) <= 1e-10, k
