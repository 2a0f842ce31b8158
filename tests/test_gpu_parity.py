"""GPU parity tests: the CUDA path (through the layer API -> ctypes -> libnpmlconv.so) against
 (a) the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
 (b) the pinned CPU oracle (oracle/conv_oracle.py) on seeded inputs,
in the three arithmetic modes with the tolerances BASELINE.json states (normwise relative error
<= 1e-5 fp32, <= 5e-3 TF32, <= 2e-2 BF16), plus size-independent properties at the full
benchmark shapes.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest

from conftest import TOL, load_golden, to_pad, rel_err, npml
from oracle import conv_oracle as O

pytestmark = pytest.mark.gpu

MODES = ["fp32", "tf32", "bf16"]
# BatchNorm's dX and everything downstream of it subtracts nearly equal quantities; the reference's own
# tests use decimal=2..3 for those tensors (tests/test_nn.py:1049, 1918-1965).  fp32-mode bound used here:
TOL_BN = {"fp32": 5e-5, "tf32": 1e-2, "bf16": 4e-2}
# Activations whose derivative jumps at a point.  In the reduced-precision modes a pre-activation that the
# float64 reference puts within rounding distance of the kink can land on the other side, which switches
# that element's gradient on or off (the reference's own tests carry the same caveat for ReLU,
# tests/test_nn.py:1973-1978).  For these activations the reduced-precision backward is therefore checked
# against the oracle evaluated at the pre-activations the device produced (whose own parity is asserted
# first), so that the comparison measures the gradient arithmetic and not the mask flips.
KINKED = {"relu", "leaky_relu", "hard_sigmoid", "elu", "selu"}


def close(a, ref, tol, what=""):
    e = rel_err(npml().to_numpy(a) if not isinstance(a, np.ndarray) else a, ref)
    assert e <= tol, "{}: rel err {:.3e} > {:.1e}".format(what, e, tol)


def make_act(pkg, name, p0, p1):
    A = pkg.activations
    return {"identity": A.Identity, "affine": lambda: A.Affine(p0, p1), "relu": A.ReLU,
            "leaky_relu": lambda: A.LeakyReLU(p0), "tanh": A.Tanh, "sigmoid": A.Sigmoid, "elu": lambda: A.ELU(p0),
            "selu": A.SELU, "hard_sigmoid": A.HardSigmoid, "softplus": A.SoftPlus, "gelu": A.GELU,
            "exponential": A.Exponential}[name]()


# ------------------------------------------------------------------------------------------------
#  golden vectors from the reference
# ------------------------------------------------------------------------------------------------
def test_utils_golden(cuda_device):
    pkg = npml()
    z, cases = load_golden("utils.npz")
    for c in cases:
        i = c.get("id")
        if c["kind"] == "pad2D":
            ks = tuple(c["kernel_shape"]) if c["kernel_shape"] else None
            Xp, p4 = pkg.pad2D(z[i + "/X"], to_pad(c["pad"]), ks, c["stride"], c["dilation"])
            assert list(p4) == c["p4"] and np.array_equal(Xp, z[i + "/out"])
        elif c["kind"] == "dilate":
            assert np.array_equal(pkg.dilate(z[i + "/X"], c["d"]), z[i + "/out"])
        elif c["kind"] == "conv":
            X, W, pad = z[i + "/X"], z[i + "/W"], to_pad(c["pad"])
            X_col, p4 = pkg.im2col(X, W.shape, pad, c["stride"], c["dilation"])
            assert list(p4) == c["p4"]
            assert np.array_equal(X_col, z[i + "/X_col"])                      # a pure gather: bit-exact
            close(pkg.conv2D(X, W, c["stride"], pad, c["dilation"]), z[i + "/Z"], TOL["fp32"], i)
            close(pkg.conv2D_naive(X, W, c["stride"], pad, c["dilation"]), z[i + "/Z"], TOL["fp32"], i)
            back = pkg.col2im(z[i + "/C"], X.shape, W.shape, tuple(c["p4"]), c["stride"], c["dilation"])
            assert back.shape == z[i + "/col2im"].shape                        # NCHW, like the reference
            close(back, z[i + "/col2im"], TOL["fp32"], i)
        elif c["kind"] == "deconv":
            close(pkg.deconv2D_naive(z[i + "/X"], z[i + "/W"], c["stride"], to_pad(c["pad"]), 0), z[i + "/Z"],
                  TOL["fp32"], i)


def test_activations_golden(cuda_device):
    pkg = npml()
    z, cases = load_golden("activations.npz")
    x = z["x"]
    for c in cases:
        a = make_act(pkg, c["act"], c["p0"], c["p1"])
        assert str(a) == c["str"]
        close(a.fn(x), z[c["act"] + "/fn"], 2e-6, c["act"])
        close(a.grad(x), z[c["act"] + "/grad"], 2e-6, c["act"] + "'")


def _run_layer_case(pkg, cls, z, c, mode):
    i, pad = c["id"], to_pad(c["pad"])
    kw = dict(out_ch=c["out_ch"], kernel_shape=tuple(c["kernel_shape"]), pad=pad, stride=c["stride"],
              act_fn=make_act(pkg, c["act"], c["p0"], c["p1"]), mode=mode)
    if "dilation" in c:
        kw["dilation"] = c["dilation"]
    layer = cls(**kw)
    X = z[i + "/X"]
    layer.forward(X, retain_derived=False)
    layer.parameters["W"], layer.parameters["b"] = z[i + "/W"], z[i + "/b"]    # NumPy assignment, like upstream
    tol = TOL[mode]
    Y = layer.forward(X)
    close(Y, z[i + "/Y"], tol, i + " Y")
    close(layer.derived_variables["Z"][0], z[i + "/Z"], tol, i + " Z")
    assert layer.derived_variables["out_rows"][0] == z[i + "/Z"].shape[1]
    if c["two_calls"]:
        Y2 = layer.forward(z[i + "/X2"])
        close(Y2, z[i + "/Y2"], tol, i + " Y2")
        dXs = layer.backward([z[i + "/dLdy"], z[i + "/dLdy2"]])
        assert isinstance(dXs, list) and len(dXs) == 2
        close(dXs[0], z[i + "/dX"], tol, i + " dX")
        close(dXs[1], z[i + "/dX2"], tol, i + " dX2")
    else:
        close(layer.backward(z[i + "/dLdy"]), z[i + "/dX"], tol, i + " dX")
    close(layer.gradients["W"], z[i + "/dW"], tol, i + " dW")
    close(layer.gradients["b"], z[i + "/db"], tol, i + " db")


@pytest.mark.parametrize("mode", MODES)
def test_conv2d_layer_golden(cuda_device, mode):
    pkg = npml()
    z, cases = load_golden("conv2d_layer.npz")
    for c in cases:
        _run_layer_case(pkg, pkg.Conv2D, z, c, mode)


@pytest.mark.parametrize("mode", MODES)
def test_deconv2d_layer_golden(cuda_device, mode):
    pkg = npml()
    z, cases = load_golden("deconv2d_layer.npz")
    for c in cases:
        _run_layer_case(pkg, pkg.Deconv2D, z, c, mode)


def test_bn_add_golden(cuda_device):
    pkg = npml()
    z, cases = load_golden("bn_add.npz")
    for c in cases:
        i = c["id"]
        if c["kind"] == "bn":
            L = pkg.BatchNorm2D(momentum=c["momentum"], epsilon=c["epsilon"])
            L.forward(z[i + "/X"], retain_derived=False)
            L.parameters["scaler"], L.parameters["intercept"] = z[i + "/scaler"], z[i + "/intercept"]
            y = L.forward(z[i + "/X"])
            close(y, z[i + "/y"], 1e-5, i + " y")
            close(L.parameters["running_mean"], z[i + "/running_mean"], 1e-5, i + " rm")
            close(L.parameters["running_var"], z[i + "/running_var"], 1e-5, i + " rv")   # biased, layers.py:1147
            dX = L.backward(z[i + "/dLdy"])
            close(dX, z[i + "/dX"], TOL_BN["fp32"], i + " dX")
            close(L.gradients["scaler"], z[i + "/dScaler"], 1e-5, i + " dScaler")
            close(L.gradients["intercept"], z[i + "/dIntercept"], 1e-5, i + " dIntercept")
            y2 = L.forward(z[i + "/X2"], retain_derived=False)                           # running-stats path
            close(y2, z[i + "/y_infer"], 1e-5, i + " y_infer")
        else:
            A = pkg.Add(act_fn=make_act(pkg, c["act"], 0, 0))
            Xs = [z["{}/X{}".format(i, j)] for j in range(c["n_in"])]
            close(A.forward(Xs), z[i + "/Y"], 1e-6, i + " Y")
            g = A.backward(z[i + "/dLdY"])
            assert len(g) == c["n_in"]
            close(g[-1], z[i + "/dX"], 1e-6, i + " dX")


@pytest.mark.parametrize("fuse", [True, False], ids=["fused-tail", "layer-by-layer"])
@pytest.mark.parametrize("mode", MODES)
def test_skip_conv_module_golden(cuda_device, mode, fuse):
    pkg = npml()
    z, cases = load_golden("skip_conv_module.npz")
    tol = TOL_BN[mode]
    for c in cases:
        i = c["id"]
        act = None if c["act"] == "identity" else make_act(pkg, c["act"], c["p0"], c["p1"])
        M = pkg.SkipConnectionConvModule(out_ch1=c["out_ch1"], out_ch2=c["out_ch2"],
                                         kernel_shape1=tuple(c["kernel_shape1"]), kernel_shape2=tuple(c["kernel_shape2"]),
                                         kernel_shape_skip=tuple(c["kernel_shape_skip"]), pad1=to_pad(c["pad1"]),
                                         pad2=to_pad(c["pad2"]), stride1=c["stride1"], stride2=c["stride2"],
                                         stride_skip=c["stride_skip"], act_fn=act, mode=mode)
        M.fuse_training = fuse          # one pass each way over the residual tail (default) / the seven layers one by one
        X = z[i + "/X"]
        M.forward(X, retain_derived=False)
        assert list(M.pad_skip) == list(z[i + "/pad_skip"])
        for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
            conv.parameters["W"], conv.parameters["b"] = z["{}/W{}".format(i, t)], z["{}/b{}".format(i, t)]
            bn.parameters["scaler"], bn.parameters["intercept"] = z["{}/g{}".format(i, t)], z["{}/h{}".format(i, t)]
            bn.reset_running_stats()
        M.flush_gradients()
        Y = M.forward(X)
        z1_dev = pkg.to_numpy(M.conv1.derived_variables["Z"][0])
        sum_dev = pkg.to_numpy(M._tail[0] if M._tail is not None else M.add3.derived_variables["sum"][0])
        dX = M.backward(z[i + "/dLdY"])
        ref = {k[len(i) + 1:]: z[k] for k in z.files if k.startswith(i + "/")}
        if mode != "fp32" and c["act"] in KINKED:
            # backward through a kinked activation in a reduced-precision mode: a pre-activation within rounding distance
            # of the kink may land on the other side, and ONE flipped mask among these ~1 000 elements is a 3 % normwise
            # error.  The backward is therefore compared with the oracle evaluated at the device's own masks: the oracle's
            # forward is first pinned to the reference's fixture, then its backward runs with z1 / sum from the device.
            P = {}
            for t in ("1", "2", "s"):
                P["W" + t], P["b" + t], P["g" + t], P["h" + t] = ref["W" + t], ref["b" + t], ref["g" + t], ref["h" + t]
                P["rm" + t], P["rv" + t] = np.zeros_like(ref["g" + t]), np.ones_like(ref["g" + t])
            o = O.skip_conv_module_fwd(X, P, tuple(c["kernel_shape1"]), tuple(c["kernel_shape2"]), tuple(c["kernel_shape_skip"]),
                                       to_pad(c["pad1"]), to_pad(c["pad2"]), c["stride1"], c["stride2"], c["stride_skip"],
                                       c["act"], c["p0"], c["p1"])
            for k in ("Y", "conv1_out", "batchnorm2_out"):
                assert rel_err(o[k], ref[k]) <= 1e-12, k                       # the oracle IS the reference here
            close(z1_dev, o["_z1"], tol, i + " Z1")
            ref = dict(ref)
            ref.update(O.skip_conv_module_bwd(o, X, ref["dLdY"], P, c["stride1"], c["stride2"], c["stride_skip"], c["act"],
                                              c["p0"], c["p1"], z1=z1_dev, total=sum_dev))
        close(Y, ref["Y"], tol, i + " Y")
        close(dX, ref["dX"], tol, i + " dX")
        dv = M.derived_variables
        for k in ["conv1_out", "conv2_out", "batchnorm1_out", "batchnorm2_out", "conv_skip_out", "batchnorm_skip_out",
                  "dLdBn1", "dLdBn2", "dLdConv1", "dLdConv2", "dLdBnSkip", "dLdConvSkip"]:
            close(dv[k], ref[k], tol, i + " " + k)
        for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
            close(conv.gradients["W"], ref["dW" + t], tol, i + " dW" + t)
            # conv biases feed straight into a BatchNorm, so their true gradient is ~0 (pure cancellation):
            # compare absolutely, scaled by the weight-gradient magnitude
            db = pkg.to_numpy(conv.gradients["b"])
            scale = np.linalg.norm(ref["dW" + t].ravel())
            assert np.linalg.norm((db - ref["db" + t]).ravel()) <= tol * max(scale, 1.0), i + " db" + t
            close(bn.gradients["scaler"], ref["dg" + t], tol, i + " dg" + t)
            close(bn.gradients["intercept"], ref["dh" + t], tol, i + " dh" + t)
            close(bn.parameters["running_mean"], ref["rm" + t], tol, i + " rm" + t)
            close(bn.parameters["running_var"], ref["rv" + t], tol, i + " rv" + t)


# ------------------------------------------------------------------------------------------------
#  seeded inputs vs the pinned oracle, shapes that take the tensor-core kernels
# ------------------------------------------------------------------------------------------------
def _f32(rng, shape, scale=1.0):
    return (scale * rng.standard_normal(shape, dtype=np.float32)).astype(np.float64)


CONV_CASES = [
    # xs, out_ch, kernel, pad, stride, dilation, act
    ((4, 16, 16, 64), 64, (3, 3), "same", 1, 0, "identity"),        # cfg2 miniature
    ((3, 28, 28, 128), 256, (3, 3), 0, 2, 2, "identity"),           # cfg3 at N=3
    ((2, 14, 14, 128), 64, (3, 3), "same", 2, 2, "relu"),           # 'same' with stride 2 / dilation 2
    ((4, 12, 12, 64), 128, (1, 1), 0, 1, 0, "tanh"),                # 1x1 projection (cfg4 skip)
    ((2, 9, 11, 72), 40, (3, 2), (1, 2), 1, 0, "sigmoid"),          # channel tails, non-square, ragged tiles
    ((5, 7, 7, 32), 96, (3, 3), 1, 1, 1, "leaky_relu"),
    ((2, 10, 10, 16), 24, (2, 2), (1, 0, 2, 1), 3, 0, "identity"),  # stride 3, leftovers, asymmetric pad
    ((2, 8, 8, 8), 8, (3, 3), 1, 1, 0, "identity"),                 # minimum aligned channels
    ((3, 6, 6, 5), 17, (3, 3), 1, 1, 0, "identity"),                # odd channels -> CUDA-core path
    # slab kernels (stride 1): weight streaming, two channel blocks, several N tiles, generic tap count, units that
    # span images, and the shapes that must fall back to the per-tile kernels
    ((3, 12, 12, 128), 128, (3, 3), 1, 1, 0, "identity"),           # cfg4 conv2 miniature: nkb = 2, streamed weights
    ((5, 10, 10, 64), 128, (3, 3), 1, 1, 0, "relu"),                # cfg4 conv1 miniature: Z and Y both written
    ((2, 6, 6, 64), 192, (3, 3), "same", 1, 0, "identity"),         # out_ch > 128: two N tiles per unit
    ((2, 9, 9, 32), 32, (5, 5), 2, 1, 0, "tanh"),                   # 25 taps: run-time tap loop
    ((1, 3, 3, 16), 16, (3, 3), 0, 1, 0, "identity"),               # one output pixel per image
    ((1, 2, 260, 8), 8, (3, 3), "same", 1, 0, "identity"),          # padded row wider than a TMA box: per-tile path
    ((9, 5, 5, 64), 64, (3, 3), "same", 1, 3, "identity"),          # dilation 3 at stride 1, units span images
    # thin shapes (cfg1): in_ch = 1 and out_ch <= 4 take the CUDA-core thin kernels in every mode
    ((3, 32, 32, 64), 128, (3, 3), 1, 2, 0, "identity"),            # strided conv: dX by output-parity classes on the slab
    ((2, 30, 30, 64), 64, (4, 4), 1, 2, 0, "tanh"),                 # 4x4 s2 (the Deconv2D adjoint shape)
    ((4, 12, 12, 1), 32, (3, 3), "same", 1, 0, "identity"),         # cfg1 miniature
    ((3, 7, 9, 24), 2, (3, 3), 1, 1, 0, "sigmoid"),                 # out_ch = 2 forward
    # one-launch backward of thin stride-1 'same' shapes (csrc/thin_bwd.cu): 256 blocks = 16 reduction groups; 25 taps
    # with dilation 2; an even kernel ('same' pads asymmetrically), 16 lanes per pixel; one lane per pixel
    ((64, 12, 12, 2), 16, (3, 3), "same", 1, 0, "relu"),
    ((3, 9, 11, 3), 8, (5, 5), "same", 1, 1, "tanh"),
    ((2, 10, 7, 4), 64, (3, 2), "same", 1, 0, "identity"),
    ((5, 8, 8, 1), 4, (3, 3), "same", 1, 0, "sigmoid"),
    # wide outputs on the slab kernel: one N = 144 / 256 tile spanning two TMEM accumulators, 320 = 128 + 128 + 64
    ((2, 8, 8, 64), 144, (3, 3), 1, 1, 0, "relu"),
    ((3, 9, 9, 64), 256, (3, 3), 1, 1, 0, "identity"),
    ((2, 8, 8, 32), 320, (3, 3), 1, 1, 0, "tanh"),
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", range(len(CONV_CASES)))
def test_conv2d_vs_oracle(cuda_device, mode, case):
    pkg = npml()
    xs, oc, ks, pad, s, d, act = CONV_CASES[case]
    rng = np.random.default_rng(100 + case)
    p0 = 0.25 if act == "leaky_relu" else 0.0
    L = pkg.Conv2D(out_ch=oc, kernel_shape=ks, pad=pad, stride=s, dilation=d, act_fn=make_act(pkg, act, p0, 0.0),
                   mode=mode)
    X = _f32(rng, xs)
    W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
    b = _f32(rng, (1, 1, 1, oc), 0.1)                                # non-zero bias (never exercised upstream)
    L.forward(X, retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    Y = L.forward(X)
    Yr, Zr = O.conv2d_layer_fwd(X, W, b, s, pad, d, act, p0, 0.0)
    tol = TOL[mode]
    close(Y, Yr, tol, "Y")
    dY = _f32(rng, Yr.shape)                                         # random dLdy (upstream uses ones)
    dX = L.backward(dY)
    if mode != "fp32" and act in KINKED:
        close(L.derived_variables["Z"][0], Zr, tol, "Z")
        Zr = pkg.to_numpy(L.derived_variables["Z"][0])
    dXr, dWr, dbr = O.conv2d_layer_bwd(dY, X, Zr, W, s, pad, d, act, p0, 0.0)
    close(dX, dXr, tol, "dX")
    close(L.gradients["W"], dWr, tol, "dW")
    close(L.gradients["b"], dbr, tol, "db")


@pytest.mark.parametrize("mode", MODES)
def test_fused_thin_backward_equals_the_three_kernel_path(cuda_device, mode, monkeypatch):
    """npml_conv2d_bwd (one launch: dX, dW, db with the in-kernel fixed-order reduction) against the same backward run
    as wgrad + reduce + dgrad (NPML_DISABLE_THIN_BWD=1), gradient accumulation over two backward calls, and bitwise
    run-to-run determinism of the in-kernel reduction (the grid barrier must be reusable launch after launch)."""
    import torch
    pkg = npml()
    rng = np.random.default_rng(7)
    xs, oc = (40, 28, 28, 1), 32                                       # cfg1 shape
    L = pkg.Conv2D(out_ch=oc, kernel_shape=(3, 3), pad="same", mode=mode)
    X, dY = _f32(rng, xs), _f32(rng, xs[:3] + (oc,))
    L.forward(X, retain_derived=False)
    L.parameters["W"], L.parameters["b"] = _f32(rng, (3, 3, 1, oc), 1 / 3.0), _f32(rng, (1, 1, 1, oc), 0.1)

    def run(times=1):
        L.flush_gradients()
        for _ in range(times):
            L.forward(X)
        dX = L.backward([dY] * times) if times > 1 else L.backward(dY)
        torch.cuda.synchronize()
        dX = dX[-1] if isinstance(dX, list) else dX
        return pkg.to_numpy(dX), pkg.to_numpy(L.gradients["W"]).copy(), pkg.to_numpy(L.gradients["b"]).copy()

    d = L._desc((40, 28, 28, 1))
    assert pkg._lib.load().npml_conv2d_bwd_is_fused(d) == 1
    a = run()
    b = run()
    for u, v in zip(a, b):
        assert np.array_equal(u, v)                                     # bitwise, run after run
    a2 = run(times=2)                                                   # two retained forwards: gradients add up
    close(a2[1], 2 * a[1], 1e-6, "dW accumulated")
    close(a2[2], 2 * a[2], 1e-6, "db accumulated")
    monkeypatch.setenv("NPML_MAX_CTAS", "3")                            # 12 resident blocks walk 560 tiles
    m = run()
    monkeypatch.delenv("NPML_MAX_CTAS")
    for u, v, w in zip(a, m, ("dX", "dW", "db")):
        close(u, v, 1e-6 if mode != "bf16" else 2e-3, w + " (few blocks, many tiles each)")
    monkeypatch.setenv("NPML_DISABLE_THIN_BWD", "1")
    assert pkg._lib.load().npml_conv2d_bwd_is_fused(d) == 0
    c = run()
    tol = 1e-6 if mode != "bf16" else 2e-3                             # same operands, different summation order
    for u, v, w in zip(a, c, ("dX", "dW", "db")):
        close(u, v, tol, w)


def test_conv2d_bwd_entry_point_on_a_shape_that_is_not_fused(cuda_device):
    """npml_conv2d_bwd documents a fallback (wgrad_db, then dgrad) for shapes the one-launch kernel does not take; the
    host layer never calls it there, so the C ABI is driven directly and compared with the two separate calls."""
    import torch
    pkg = npml()
    L = pkg._lib
    lib = L.load()
    rng = np.random.default_rng(11)
    layer = pkg.Conv2D(out_ch=32, kernel_shape=(3, 3), pad=1, stride=2, mode="bf16")          # stride 2: never fused
    X = L.to_device(_f32(rng, (3, 12, 12, 16)), torch.bfloat16)
    layer.forward(X, retain_derived=False)
    W = layer.parameters["W"]
    d = layer._desc(tuple(X.shape), accumulate=0)
    assert lib.npml_conv2d_bwd_is_fused(d) == 0
    dZ = L.to_device(_f32(rng, (3, d.OH, d.OW, 32)), torch.bfloat16)
    outs = []
    for fused_entry in (True, False):
        dW, db, dX = torch.zeros_like(W), torch.zeros(32, device=W.device), torch.empty_like(X)
        if fused_entry:
            ws = L.workspace(lib.npml_workspace_bytes(d, L.OP_CONV_BWD))
            L.check(lib.npml_conv2d_bwd(d, L.ptr(X), L.ptr(dZ), L.ptr(W), L.ptr(dW), L.ptr(db), L.ptr(dX), L.ptr(ws),
                                        ws.numel(), L.stream()), "npml_conv2d_bwd")
        else:
            ws = L.workspace(lib.npml_workspace_bytes(d, L.OP_CONV_WGRAD))
            L.check(lib.npml_conv2d_wgrad_db(d, L.ptr(X), L.ptr(dZ), L.ptr(dW), L.ptr(db), L.ptr(ws), ws.numel(),
                                             L.stream()), "npml_conv2d_wgrad_db")
            ws = L.workspace(lib.npml_workspace_bytes(d, L.OP_CONV_DGRAD))
            L.check(lib.npml_conv2d_dgrad(d, L.ptr(dZ), L.ptr(W), L.ptr(dX), L.ptr(ws), ws.numel(), L.stream()),
                    "npml_conv2d_dgrad")
        torch.cuda.synchronize()
        outs.append([pkg.to_numpy(t) for t in (dW, db, dX)])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)                                     # the same kernels in the same order
    assert np.abs(outs[0][0]).max() > 0 and np.abs(outs[0][2]).max() > 0


DECONV_CASES = [
    ((2, 8, 8, 256), 128, (4, 4), 1, 2, "identity"),                 # cfg5 miniature
    ((3, 6, 5, 64), 64, (3, 3), 1, 1, "relu"),
    ((2, 5, 5, 32), 48, (2, 3), (1, 0), 3, "tanh"),
    ((2, 7, 7, 16), 8, (3, 3), "same", 1, "identity"),
    ((2, 4, 4, 64), 32, (5, 5), 2, 2, "identity"),
    ((2, 5, 5, 3), 6, (4, 4), 1, 2, "identity"),                     # odd channels -> CUDA-core path
    # output-parity classes on the slab kernel (forward): classes with 4/2/2/1 taps, unequal class sizes, units that
    # span images and classes
    ((3, 16, 16, 64), 64, (3, 3), 1, 2, "relu"),                     # 3x3 s2: classes of 16x16 / 16x15 / 15x16 / 15x15
    ((5, 12, 9, 128), 32, (4, 4), 1, 2, "tanh"),                     # two channel blocks, non-square
    ((2, 20, 20, 32), 160, (2, 2), 0, 2, "identity"),                # one tap per class, two N tiles
]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", range(len(DECONV_CASES)))
def test_deconv2d_vs_oracle(cuda_device, mode, case, monkeypatch):
    pkg = npml()
    if case % 2 == 0:
        # opt-in variant of the strided wgrad: one slab group per input parity plane (csrc/tc_slab_wgrad.cu)
        monkeypatch.setenv("NPML_SLAB_WGRAD_PLANES", "1")
    xs, oc, ks, pad, s, act = DECONV_CASES[case]
    rng = np.random.default_rng(200 + case)
    L = pkg.Deconv2D(out_ch=oc, kernel_shape=ks, pad=pad, stride=s, act_fn=make_act(pkg, act, 0, 0), mode=mode)
    X = _f32(rng, xs)
    W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
    b = _f32(rng, (1, 1, 1, oc), 0.1)
    L.forward(X, retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    Y = L.forward(X)
    Yr, Zr = O.deconv2d_layer_fwd(X, W, b, s, pad, act)
    tol = TOL[mode]
    close(Y, Yr, tol, "Y")
    dY = _f32(rng, Yr.shape)
    dX = L.backward(dY)
    if mode != "fp32" and act in KINKED:
        close(L.derived_variables["Z"][0], Zr, tol, "Z")
        Zr = pkg.to_numpy(L.derived_variables["Z"][0])
    dXr, dWr, dbr = O.deconv2d_layer_bwd(dY, X, Zr, W, s, pad, act)
    close(dX, dXr, tol, "dX")
    close(L.gradients["W"], dWr, tol, "dW")
    close(L.gradients["b"], dbr, tol, "db")


# ------------------------------------------------------------------------------------------------
#  frozen BatchNorm2D folded into the convolution epilogue (npml_conv2d_fprop_bn / npml_deconv2d_fprop_bn)
# ------------------------------------------------------------------------------------------------
def _random_bn(pkg, rng, ch):
    bn = pkg.BatchNorm2D()
    P = {"scaler": rng.uniform(0.5, 1.5, ch), "intercept": 0.1 * rng.standard_normal(ch),
         "running_mean": 0.2 * rng.standard_normal(ch), "running_var": rng.uniform(0.5, 2.0, ch)}
    P = {k: v.astype(np.float32).astype(np.float64) for k, v in P.items()}
    bn.in_ch = bn.out_ch = ch
    bn._init_params()
    for k, v in P.items():
        bn.parameters[k] = v
    return bn, P


# (kind, index into CONV_CASES / DECONV_CASES): slab (resident / streamed weights, Z and Y both written, N = 144 / 256
# tiles), per-tile gather kernel, CUDA-core and thin kernels, Deconv2D on the class slab and on the CUDA cores
FUSED_BN_CASES = [("conv", 0), ("conv", 1), ("conv", 4), ("conv", 8), ("conv", 9), ("conv", 10), ("conv", 11),
                  ("conv", 19), ("conv", 20), ("conv", 21), ("conv", 22), ("deconv", 0), ("deconv", 5), ("deconv", 6)]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", range(len(FUSED_BN_CASES)))
def test_conv_fused_frozen_bn_epilogue(cuda_device, mode, case):
    """Y = BN_running(act(conv(X) + b)) from ONE kernel equals the oracle's conv -> BatchNorm2D(inference) chain
    (layers.py:3037-3038 then 1141-1156), and the retained Z is still the pre-activation."""
    pkg = npml()
    kind, idx = FUSED_BN_CASES[case]
    rng = np.random.default_rng(700 + case)
    if kind == "conv":
        xs, oc, ks, pad, s, d, act = CONV_CASES[idx]
        p0 = 0.25 if act == "leaky_relu" else 0.0
        L = pkg.Conv2D(out_ch=oc, kernel_shape=ks, pad=pad, stride=s, dilation=d, act_fn=make_act(pkg, act, p0, 0.0),
                       mode=mode)
    else:
        xs, oc, ks, pad, s, act = DECONV_CASES[idx]
        p0, d = 0.0, 0
        L = pkg.Deconv2D(out_ch=oc, kernel_shape=ks, pad=pad, stride=s, act_fn=make_act(pkg, act, 0, 0), mode=mode)
    X = _f32(rng, xs)
    W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
    b = _f32(rng, (1, 1, 1, oc), 0.1)
    L.forward(X, retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    bn, P = _random_bn(pkg, rng, oc)
    if kind == "conv":
        Yr, Zr = O.conv2d_layer_fwd(X, W, b, s, pad, d, act, p0, 0.0)
    else:
        Yr, Zr = O.deconv2d_layer_fwd(X, W, b, s, pad, act)
    ref, _, _ = O.bn2d_fwd(Yr, P["scaler"], P["intercept"], P["running_mean"], P["running_var"], use_batch_stats=False)
    tol = TOL[mode]
    # inference call: nothing retained
    close(L.forward(X, retain_derived=False, fused_bn=bn), ref, tol, "fused Y")
    assert len(L.X) == 0 and len(L.derived_variables["Z"]) == 0
    # the two-kernel chain gives the same numbers (same arithmetic, one rounding more in bf16)
    close(bn.forward(L.forward(X, retain_derived=False), retain_derived=False), ref, tol, "unfused Y")
    # frozen BatchNorm2D with a retaining forward: Z is kept and is the pre-activation
    bn.freeze()
    close(L.forward(X, retain_derived=True, fused_bn=bn), ref, tol, "fused Y (frozen)")
    close(L.derived_variables["Z"][0], Zr, tol, "Z")
    bn.unfreeze()
    with pytest.raises(ValueError):
        L.forward(X, retain_derived=True, fused_bn=bn)                # a training-mode BatchNorm2D cannot be folded
    for k in ("running_mean", "running_var"):                         # inference never touches the running statistics
        assert np.array_equal(pkg.to_numpy(bn.parameters[k]), P[k])


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("module", ["conv", "identity"])
def test_skip_modules_fused_inference(cuda_device, mode, module):
    """retain_derived=False forward of the two skip-connection modules: each conv -> BatchNorm2D pair is one kernel;
    equals the oracle's layer-by-layer chain on the running statistics and the unfused product path."""
    pkg = npml()
    rng = np.random.default_rng(800)
    act = "relu"
    if module == "conv":
        cin, c1, c2 = 64, 128, 128
        M = pkg.SkipConnectionConvModule(out_ch1=c1, out_ch2=c2, kernel_shape1=(3, 3), kernel_shape2=(3, 3),
                                         kernel_shape_skip=(1, 1), pad1=1, pad2=1, act_fn=make_act(pkg, act, 0, 0),
                                         mode=mode)
    else:
        cin = c1 = c2 = 64
        M = pkg.SkipConnectionIdentityModule(out_ch=c1, kernel_shape1=(3, 3), kernel_shape2=(3, 3),
                                             act_fn=make_act(pkg, act, 0, 0), mode=mode)
    X = _f32(rng, (3, 12, 12, cin))
    M.forward(X, retain_derived=False)
    triples = [("1", M.conv1, M.batchnorm1, cin, c1), ("2", M.conv2, M.batchnorm2, c1, c2)]
    if module == "conv":
        triples.append(("s", M.conv_skip, M.batchnorm_skip, cin, c2))
    P = {}
    for t, conv, bn, ci, co in triples:
        k = conv.kernel_shape
        P["W" + t] = _f32(rng, k + (ci, co), 1.0 / np.sqrt(k[0] * k[1] * ci))
        P["b" + t] = _f32(rng, (1, 1, 1, co), 0.1)
        conv.parameters["W"], conv.parameters["b"] = P["W" + t], P["b" + t]
        _, q = _random_bn(pkg, rng, co)
        for name, v in q.items():
            bn.parameters[name] = v
        P["bn" + t] = q

    def bn_ref(x, q):
        return O.bn2d_fwd(x, q["scaler"], q["intercept"], q["running_mean"], q["running_var"], use_batch_stats=False)[0]

    pad1 = 1 if module == "conv" else "same"
    a1, _ = O.conv2d_layer_fwd(X, P["W1"], P["b1"], 1, pad1, 0, act)
    h1 = bn_ref(a1, P["bn1"])
    a2, _ = O.conv2d_layer_fwd(h1, P["W2"], P["b2"], 1, pad1, 0, "identity")
    h2 = bn_ref(a2, P["bn2"])
    if module == "conv":
        a_s, _ = O.conv2d_layer_fwd(X, P["Ws"], P["bs"], 1, 0, 0, "identity")
        short = bn_ref(a_s, P["bns"])
    else:
        short = X
    ref, _ = O.add_fwd([short, h2], act)
    tol = TOL_BN[mode]
    M.forward(X, retain_derived=False)              # the convolutions pack the newly assigned weights once
    pkg.launch_count(reset=True)
    Y = M.forward(X, retain_derived=False)
    fused_launches = pkg.launch_count(reset=True)
    close(Y, ref, tol, "fused module Y")
    M.fuse_inference = False
    Y2 = M.forward(X, retain_derived=False)
    plain_launches = pkg.launch_count(reset=True)
    close(Y2, ref, tol, "layer-by-layer module Y")
    assert fused_launches < plain_launches
    assert M.derived_variables["conv1_out"] is None and len(M.conv1.X) == 0


@pytest.mark.parametrize("act", ["relu", "tanh"])
@pytest.mark.parametrize("mode", MODES)
def test_skip_conv_module_vs_oracle(cuda_device, mode, act):
    """cfg4 at N=4, 16x16: 64 -> 128 channels, 3x3 / 3x3 / 1x1 (tensor-core kernels in tf32 / bf16)."""
    pkg = npml()
    rng = np.random.default_rng(300)
    X = _f32(rng, (4, 16, 16, 64))
    M = pkg.SkipConnectionConvModule(out_ch1=128, out_ch2=128, kernel_shape1=(3, 3), kernel_shape2=(3, 3),
                                     kernel_shape_skip=(1, 1), pad1=1, pad2=1, act_fn=make_act(pkg, act, 0, 0), mode=mode)
    M.forward(X, retain_derived=False)
    P = {}
    for t, conv, bn, cin in (("1", M.conv1, M.batchnorm1, 64), ("2", M.conv2, M.batchnorm2, 128),
                             ("s", M.conv_skip, M.batchnorm_skip, 64)):
        k = conv.kernel_shape
        P["W" + t] = _f32(rng, k + (cin, 128), 1.0 / np.sqrt(k[0] * k[1] * cin))
        P["b" + t] = _f32(rng, (1, 1, 1, 128), 0.1)
        P["g" + t] = rng.uniform(0.5, 1.5, 128).astype(np.float32).astype(np.float64)
        P["h" + t] = _f32(rng, (128,), 0.1)
        P["rm" + t], P["rv" + t] = np.zeros(128), np.ones(128)
        conv.parameters["W"], conv.parameters["b"] = P["W" + t], P["b" + t]
        bn.parameters["scaler"], bn.parameters["intercept"] = P["g" + t], P["h" + t]
        bn.reset_running_stats()
    M.flush_gradients()
    Y = M.forward(X)
    dY = _f32(rng, Y.shape)
    dX = M.backward(dY)
    o = O.skip_conv_module_fwd_bwd(X, dY, P, (3, 3), (3, 3), (1, 1), 1, 1, 1, 1, 1, act)
    tol = TOL_BN[mode]
    close(Y, o["Y"], tol, "Y")
    for k in ("conv1_out", "batchnorm1_out", "conv2_out", "batchnorm2_out", "conv_skip_out", "batchnorm_skip_out"):
        close(M.derived_variables[k], o[k], tol, k)
    # backward: with ReLU in a reduced-precision mode a fraction ~err(Z)/std(Z) of the two ReLU masks flips
    # (see KINKED above), which bounds the normwise gradient error at roughly sqrt(2 * that fraction)
    tol_b = tol if (mode == "fp32" or act not in KINKED) else {"tf32": 5e-2, "bf16": 1.5e-1}[mode]
    close(dX, o["dX"], tol_b, "dX")
    for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
        close(conv.gradients["W"], o["dW" + t], tol_b, "dW" + t)
        close(bn.gradients["scaler"], o["dg" + t], tol_b, "dg" + t)
        close(bn.gradients["intercept"], o["dh" + t], tol_b, "dh" + t)
        close(bn.parameters["running_var"], o["rv" + t], tol, "rv" + t)


@pytest.mark.gpu
def test_sync_batchnorm_split_equals_whole_batch(cuda_device):
    """sync-BN: per-shard statistics summed over two emulated ranks, then applied per shard, reproduce
    the whole-batch BatchNorm2D forward and backward of the oracle (layers.py:1146-1156, 1192-1215)."""
    import ctypes as C
    import torch
    pkg = npml()
    L = pkg._lib
    lib = L.load()
    rng = np.random.default_rng(77)
    N, H, W, ch = 8, 9, 7, 24
    X = _f32(rng, (N, H, W, ch)) * 2.0 + 0.5
    dY = _f32(rng, (N, H, W, ch))
    g = rng.uniform(0.5, 1.5, ch).astype(np.float32).astype(np.float64)
    h = _f32(rng, (ch,), 0.1)
    Yr, rm_r, rv_r = O.bn2d_fwd(X, g, h, np.zeros(ch), np.ones(ch), 0.9, 1e-5)
    dXr, dgr, dhr = O.bn2d_bwd(dY, X, g, 1e-5)
    halves = [(0, N // 2), (N // 2, N)]
    dev = torch.device("cuda:0")
    f32 = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
    gd, hd = f32(g), f32(h)
    xs = [f32(X[a:b]) for a, b in halves]
    dys = [f32(dY[a:b]) for a, b in halves]
    rows = [x.numel() // ch for x in xs]
    ws = L.workspace(1 << 22)
    sums = [torch.empty(2 * ch, dtype=torch.float64, device=dev) for _ in halves]
    for x, r, s_ in zip(xs, rows, sums):
        L.check(lib.npml_bn2d_stats(r, ch, L.ptr(x), L.F32, L.ptr(s_), L.ptr(ws), ws.numel(), L.stream()), "stats")
    tot = sums[0] + sums[1]                       # what the all-reduce produces on every rank
    ys, means, rstds, rms = [], [], [], []
    for x, r in zip(xs, rows):
        y = torch.empty_like(x)
        mean, rstd = torch.empty(ch, device=dev), torch.empty(ch, device=dev)
        rm, rv = torch.zeros(ch, device=dev), torch.ones(ch, device=dev)
        L.check(lib.npml_bn2d_apply(r, sum(rows), ch, L.ptr(x), L.ptr(y), L.F32, L.ptr(gd), L.ptr(hd), L.ptr(rm), L.ptr(rv),
                                    0.9, 1e-5, L.ptr(tot), L.ptr(mean), L.ptr(rstd), L.ptr(ws), ws.numel(), L.stream()),
                "apply")
        ys.append(y); means.append(mean); rstds.append(rstd); rms.append((rm, rv))
    Y = np.concatenate([pkg.to_numpy(y) for y in ys], axis=0)
    close(Y, Yr, 1e-5, "sync-BN Y")
    close(pkg.to_numpy(rms[0][0]), rm_r, 1e-5, "running_mean")
    close(pkg.to_numpy(rms[1][1]), rv_r, 1e-5, "running_var")
    bs = [torch.empty(2 * ch, dtype=torch.float64, device=dev) for _ in halves]
    for x, dy, r, s_, m, rs in zip(xs, dys, rows, bs, means, rstds):
        L.check(lib.npml_bn2d_bwd_stats(r, ch, L.ptr(dy), L.ptr(x), L.F32, L.ptr(m), L.ptr(rs), L.ptr(s_), L.ptr(ws),
                                        ws.numel(), L.stream()), "bwd_stats")
    btot = bs[0] + bs[1]
    dxs = []
    dg, dh = torch.zeros(ch, device=dev), torch.zeros(ch, device=dev)    # "bucket": both ranks accumulate, i.e. SUM
    for x, dy, r, s_, m, rs in zip(xs, dys, rows, bs, means, rstds):
        dx = torch.empty_like(x)
        L.check(lib.npml_bn2d_bwd_apply(r, sum(rows), ch, L.ptr(dy), L.ptr(x), L.ptr(dx), L.F32, L.ptr(gd), L.ptr(m),
                                        L.ptr(rs), L.ptr(s_), L.ptr(btot), L.ptr(dg), L.ptr(dh), 1, L.ptr(ws), ws.numel(),
                                        L.stream()), "bwd_apply")
        dxs.append(dx)
    close(np.concatenate([pkg.to_numpy(d) for d in dxs], axis=0), dXr, 2e-5, "sync-BN dX")
    close(pkg.to_numpy(dg), dgr, 2e-5, "dScaler")
    close(pkg.to_numpy(dh), dhr, 2e-5, "dIntercept")
    # the layer-level switch (one rank: the split path must equal the fused path)
    A, B = pkg.BatchNorm2D(sync=True), pkg.BatchNorm2D(sync=False)
    outs = []
    for layer in (A, B):
        layer.forward(X, retain_derived=False)
        layer.parameters["scaler"], layer.parameters["intercept"] = g, h
        layer.reset_running_stats()
        y = layer.forward(X)
        dx = layer.backward(dY)
        outs.append((y, dx, pkg.to_numpy(layer.gradients["scaler"])))
    for a, b in zip(*outs):
        close(a, b, 1e-6, "sync vs fused")
    close(outs[0][0], Yr, 1e-5, "layer sync Y")


# ------------------------------------------------------------------------------------------------
#  layer state semantics (layers.py:3040-3044, 3076-3091, 66-86)
# ------------------------------------------------------------------------------------------------
def test_layer_state_semantics(cuda_device):
    pkg = npml()
    rng = np.random.default_rng(5)
    L = pkg.Conv2D(4, (3, 3), pad=1)
    X = _f32(rng, (2, 6, 6, 3))
    L.forward(X, retain_derived=False)
    assert L.X == [] and L.derived_variables["Z"] == []
    Y = L.forward(X)
    assert isinstance(Y, np.ndarray) and Y.dtype == np.float64 and len(L.X) == 1
    dY = _f32(rng, Y.shape)
    L.backward(dY, retain_grads=False)
    assert float(pkg.to_numpy(L.gradients["W"]).std()) == 0.0            # untouched
    L.backward(dY)
    g1 = pkg.to_numpy(L.gradients["W"]).copy()
    L.backward(dY)                                                       # accumulates (+=)
    close(L.gradients["W"], 2 * g1, 1e-6, "accumulate")
    W0 = pkg.to_numpy(L.parameters["W"]).copy()
    L.update()                                                           # SGD(lr=0.01) step, then flush
    close(L.parameters["W"], W0 - 0.01 * 2 * g1, 1e-6, "sgd")
    assert L.X == [] and float(np.abs(pkg.to_numpy(L.gradients["W"])).max()) == 0.0
    L.freeze()
    L.forward(X)
    with pytest.raises(AssertionError):
        L.backward(dY)
    with pytest.raises(AssertionError):
        L.update()
    s = L.summary()
    assert s["layer"] == "Conv2D" and s["hyperparameters"]["kernel_shape"] == (3, 3)
    with pytest.raises(ValueError):
        pkg.Conv2D(2, (5, 5)).forward(np.zeros((1, 3, 3, 1)))            # kernel > input (utils.py:463-467)


@pytest.mark.parametrize("mode", MODES)
def test_conv_epilogue_emits_batchnorm_statistics(cuda_device, mode):
    """npml_conv2d_fprop_stats: the per-channel [sum Y, sum Y^2] a training-mode BatchNorm2D reduces (layers.py:1146-1147)
    come out of the convolution call itself -- from the slab kernel's epilogue (64 / 128 / 256-channel shapes, several
    work units per CTA, ragged last tile) or from the column pass that follows the other kernels -- and BatchNorm2D fed
    with them equals BatchNorm2D that reads the tensor itself."""
    import torch
    pkg = npml()
    shapes = [((9, 20, 20, 64), 64, (3, 3), 1, 1, "relu"), ((5, 13, 11, 64), 128, (3, 3), 1, 1, "identity"),
              ((3, 9, 9, 64), 256, (1, 1), 0, 1, "tanh"), ((4, 12, 12, 128), 128, (3, 3), 1, 1, "relu"),
              ((2, 28, 28, 128), 64, (3, 3), 0, 2, "identity"), ((3, 7, 7, 5), 6, (3, 3), 1, 1, "sigmoid")]
    for ctas in (None, 2):
        for xs, oc, ks, pad, s, act in shapes:
            if ctas is not None:
                os.environ["NPML_MAX_CTAS"] = str(ctas)
            try:
                rng = np.random.default_rng(77)
                L_ = pkg.Conv2D(oc, ks, pad=pad, stride=s, act_fn=make_act(pkg, act, 0, 0), mode=mode)
                X = _f32(rng, xs)
                sums = torch.zeros(2 * oc, dtype=torch.float64, device="cuda")
                L_.forward(X[:1], retain_derived=False)
                L_.parameters["b"] = _f32(rng, (1, 1, 1, oc), 0.2)
                Xd = pkg._lib.to_device(X, pkg._lib.torch_dtype(mode))
                Y = L_.forward(Xd, stats_out=sums)
                Yd = Y.double().reshape(-1, oc)
                # the kernel sums the fp32 values before they are rounded to the storage type: compare at the mode's tolerance
                tol = {"fp32": 1e-6, "tf32": 1e-6, "bf16": 4e-3}[mode]
                got = sums.cpu().numpy()
                ref1, ref2 = Yd.sum(0).cpu().numpy(), (Yd * Yd).sum(0).cpu().numpy()
                scale = float(Yd.abs().sum(0).max())
                assert np.abs(got[:oc] - ref1).max() <= tol * scale, (mode, xs, oc, "sum")
                assert np.abs(got[oc:] - ref2).max() <= 2 * tol * float(ref2.max()), (mode, xs, oc, "sumsq")
                bn_a, bn_b = pkg.BatchNorm2D(), pkg.BatchNorm2D()
                np.random.seed(3); ya = bn_a.forward(Y, sums=sums)
                np.random.seed(3); yb = bn_b.forward(Y)
                close(ya.float(), pkg.to_numpy(yb), {"fp32": 1e-5, "tf32": 1e-5, "bf16": 1e-2}[mode], "bn from conv stats")
                close(bn_a.parameters["running_var"], pkg.to_numpy(bn_b.parameters["running_var"]), 1e-3, "running var")
            finally:
                os.environ.pop("NPML_MAX_CTAS", None)


GUARD_CASES = [0, 1, 2, 3, 4, 6, 8, 9, 11, 12, 15, 16, 18, 19, 22]      # indices into CONV_CASES: every kernel family


@pytest.mark.parametrize("mode", MODES)
def test_no_writes_outside_the_output_buffers(cuda_device, mode):
    """compute-sanitizer is closed on this pool (profiles/r2s_sanitizer_closed.md), so out-of-bounds WRITES are hunted
    directly: every convolution entry point is called through the C ABI with each output tensor and the workspace placed
    between two 64 KB guard bands; the bands must come back untouched, and the results must still match the oracle."""
    import ctypes as C
    import torch
    pkg = npml()
    L_, U_ = pkg._lib, pkg.utils
    lib = L_.load()
    G = 65536
    dt = L_.torch_dtype(mode)

    class Guarded(object):
        def __init__(self, nbytes, fill=0xA5):
            self.n = int(nbytes)
            self.raw = torch.full((self.n + 2 * G + 512,), fill, dtype=torch.uint8, device="cuda")
            off = (-self.raw.data_ptr() - G) % 256                      # 256-byte aligned payload
            self.lo = G + off
            self.payload = self.raw[self.lo:self.lo + self.n]

        def view(self, shape, dtype):
            return self.payload.view(dtype).view(shape)

        def intact(self):
            a, b = self.raw[:self.lo], self.raw[self.lo + self.n:]
            return bool((a == 0xA5).all().item()) and bool((b == 0xA5).all().item())

    def ptr(t):
        return C.c_void_p(t.data_ptr())

    for case in GUARD_CASES:
        xs, oc, ks, pad, s, d, act = CONV_CASES[case]
        rng = np.random.default_rng(500 + case)
        X = _f32(rng, xs)
        W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
        b = _f32(rng, (1, 1, 1, oc), 0.1)
        p4 = U_.resolve_pad_2D(xs, pad, ks, s, d)
        oh, ow = U_.conv_out_hw(xs[1], xs[2], ks[0], ks[1], p4, s, d)
        desc = U_.make_desc(xs[0], xs[1], xs[2], xs[3], oc, ks[0], ks[1], s, d + 1, p4, oh, ow, mode)
        Xd, Wd, bd = L_.to_device(X, dt), L_.to_device(W, torch.float32), L_.to_device(b, torch.float32)
        es = 2 if mode == "bf16" else 4
        n_out, n_in = xs[0] * oh * ow * oc, int(np.prod(xs))
        gZ, gdX = Guarded(n_out * es), Guarded(n_in * es)
        gdW, gdb = Guarded(W.size * 4, 0xA5), Guarded(oc * 4, 0xA5)
        ws = [Guarded(max(256, lib.npml_workspace_bytes(desc, op))) for op in (0, 1, 2)]
        Z = gZ.view((xs[0], oh, ow, oc), dt)
        st = L_.stream()
        L_.check(lib.npml_conv2d_fprop(desc, ptr(Xd), ptr(Wd), ptr(bd), ptr(Z), None, ptr(ws[0].payload), ws[0].n, st), "fprop")
        Zr = O.conv2D(X, W, s, pad, d) + b
        close(Z, Zr, TOL[mode], "guarded Z")
        dZ = L_.to_device(_f32(rng, Zr.shape), dt)
        dX = gdX.view(xs, dt)
        L_.check(lib.npml_conv2d_dgrad(desc, ptr(dZ), ptr(Wd), ptr(dX), ptr(ws[1].payload), ws[1].n, st), "dgrad")
        dW, db = gdW.view(W.shape, torch.float32), gdb.view((1, 1, 1, oc), torch.float32)
        desc.accumulate = 0
        L_.check(lib.npml_conv2d_wgrad_db(desc, ptr(Xd), ptr(dZ), ptr(dW), ptr(db), ptr(ws[2].payload), ws[2].n, st), "wgrad")
        torch.cuda.synchronize()
        dXr, dWr, dbr = O.conv2d_layer_bwd(pkg.to_numpy(dZ), pkg.to_numpy(Xd), Zr, W, s, pad, d)
        close(dX, dXr, TOL[mode], "guarded dX")
        close(dW, dWr, TOL[mode], "guarded dW")
        close(db, dbr, TOL[mode], "guarded db")
        for name, gbuf in (("Z", gZ), ("dX", gdX), ("dW", gdW), ("db", gdb), ("ws fprop", ws[0]), ("ws dgrad", ws[1]),
                           ("ws wgrad", ws[2])):
            assert gbuf.intact(), "case {} ({}): a kernel wrote outside {}".format(case, mode, name)


def test_backward_naive_cross_check(cuda_device):
    """Conv2D._backward_naive (layers.py:3118-3178): the loop-form backward agrees with the oracle at fp32 tolerance
    from a layer that runs in bf16 mode, and accumulates into the same gradients."""
    pkg = npml()
    rng = np.random.default_rng(12)
    L = pkg.Conv2D(64, (3, 3), pad=1, stride=2, dilation=1, act_fn=pkg.activations.Tanh(), mode="bf16")
    X, dY = _f32(rng, (3, 14, 14, 64)), None
    W, b = _f32(rng, (3, 3, 64, 64), 1 / 24.0), _f32(rng, (1, 1, 1, 64), 0.1)
    L.forward(X[:1], retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    Y = L.forward(X)
    dY = _f32(rng, Y.shape)
    Zd = pkg.to_numpy(L.derived_variables["Z"][0])
    dXr, dWr, dbr = O.conv2d_layer_bwd(dY, pkg.to_numpy(L.X[0]), Zd, W, 2, 1, 1, "tanh")
    dX = L._backward_naive(dY)
    close(dX, dXr, 1e-5, "naive dX")
    close(L.gradients["W"], dWr, 1e-5, "naive dW")
    close(L.gradients["b"], dbr, 1e-5, "naive db")
    dX2 = L.backward(dY)                                   # tensor-core backward: same values at bf16 tolerance, += semantics
    close(dX2, dXr, TOL["bf16"], "tc dX")
    close(L.gradients["W"], 2 * dWr, TOL["bf16"], "accumulated dW")
    assert L.mode == "bf16"


def test_set_params_and_summary_round_trip(cuda_device):
    """LayerBase.summary / set_params (layers.py:88-136) and the module-level versions (modules.py:76-116): a layer
    rebuilt from another one's summary computes the same function."""
    pkg = npml()
    rng = np.random.default_rng(21)
    X = _f32(rng, (2, 8, 8, 8))
    A = pkg.Conv2D(16, (3, 3), pad=1, stride=1, act_fn=pkg.activations.LeakyReLU(0.2), optimizer="SGD(lr=0.1, momentum=0.5)")
    Ya = A.forward(X)
    S = A.summary()
    assert S["layer"] == "Conv2D" and set(S["parameters"]) == {"W", "b"}
    assert S["hyperparameters"]["act_fn"] == "Leaky ReLU(alpha=0.2)" and S["hyperparameters"]["optimizer"]["hyperparameters"]["momentum"] == 0.5
    B = pkg.Conv2D(16, (3, 3), pad=0, stride=1)            # different pad / activation / optimizer
    B.forward(X)
    B.flush_gradients()
    B.set_params({"parameters": {k: pkg.to_numpy(v) for k, v in S["parameters"].items()},
                  "hyperparameters": S["hyperparameters"]})
    assert B.pad == 1 and str(B.act_fn) == "Leaky ReLU(alpha=0.2)" and str(B.optimizer).startswith("SGD(lr=0.1, momentum=0.5")
    close(B.forward(X), Ya, 1e-6, "round trip Y")
    assert B.summary()["hyperparameters"]["act_fn"] == S["hyperparameters"]["act_fn"]
    # BatchNorm2D: parameters incl. the running statistics travel
    bn = pkg.BatchNorm2D(momentum=0.8)
    bn.forward(Ya)
    bn2 = pkg.BatchNorm2D()
    bn2.forward(Ya, retain_derived=False)
    s = bn.summary()
    bn2.set_params({"parameters": {k: pkg.to_numpy(v) for k, v in s["parameters"].items()}, "hyperparameters": s["hyperparameters"]})
    assert bn2.momentum == 0.8
    close(bn2.forward(Ya, retain_derived=False), bn.forward(Ya, retain_derived=False), 1e-6, "bn round trip")
    # module: nested component dicts
    M = pkg.SkipConnectionConvModule(8, 8, (3, 3), (3, 3), (1, 1), pad1=1, pad2=1, act_fn=pkg.activations.ReLU())
    Ym = M.forward(X, retain_derived=False)
    sm = M.summary()
    assert sm["layer"] == "SkipConnectionConvModule" and set(sm["parameters"]["components"]) >= {"conv1", "conv2", "conv_skip", "batchnorm1"}
    M2 = pkg.SkipConnectionConvModule(8, 8, (3, 3), (3, 3), (1, 1), pad1=1, pad2=1, act_fn=pkg.activations.ReLU())
    M2.forward(X, retain_derived=False)
    M2.set_params(sm)
    close(M2.forward(X, retain_derived=False), Ym, 1e-6, "module round trip")


def test_empty_batch_and_device_tensors(cuda_device):
    import torch
    pkg = npml()
    L = pkg.Conv2D(8, (3, 3), pad=1, mode="bf16")
    X = torch.randn(2, 8, 8, 8, device="cuda")
    Y = L.forward(X)
    assert isinstance(Y, torch.Tensor) and Y.dtype == torch.bfloat16 and Y.shape == (2, 8, 8, 8)
    E = pkg.Conv2D(8, (3, 3), pad=1)
    E.forward(np.zeros((1, 4, 4, 2)))
    E.flush_gradients()
    Ye = E.forward(np.zeros((0, 4, 4, 2)))
    assert Ye.shape == (0, 4, 4, 8)


# ------------------------------------------------------------------------------------------------
#  full benchmark shapes (BASELINE.json configs): every output tensor against the float64 oracle
# ------------------------------------------------------------------------------------------------
def _dot(a, b):
    return float((a.double() * b.double()).sum().item())


FULL_CFGS = {
    # name: (layer kind, ctor kwargs, X shape, seed)           -- the objects and shapes of SURVEY.md 8d
    "cfg1": ("conv", dict(out_ch=32, kernel_shape=(3, 3), pad="same"), (32, 28, 28, 1), 1001),
    "cfg2": ("conv", dict(out_ch=64, kernel_shape=(3, 3), pad="same"), (256, 56, 56, 64), 1002),
    "cfg3": ("conv", dict(out_ch=256, kernel_shape=(3, 3), pad=0, stride=2, dilation=2), (128, 28, 28, 128), 1003),
    "cfg3-same": ("conv", dict(out_ch=256, kernel_shape=(3, 3), pad="same", stride=2, dilation=2), (16, 28, 28, 128), 1013),
    "cfg4-conv1": ("conv", dict(out_ch=128, kernel_shape=(3, 3), pad=1), (256, 32, 32, 64), 1004),
    "cfg4-conv2": ("conv", dict(out_ch=128, kernel_shape=(3, 3), pad=1), (256, 32, 32, 128), 1014),
    "cfg4-skip": ("conv", dict(out_ch=128, kernel_shape=(1, 1), pad=0), (256, 32, 32, 64), 1024),
    "cfg5": ("deconv", dict(out_ch=128, kernel_shape=(4, 4), pad=1, stride=2), (128, 16, 16, 256), 1005),
}
_full_cache = {}


def _full_case(cfg):
    """Seeded inputs of a BASELINE config and the float64 oracle's (Y, dX, dW, db) for them.  oracle/fast_oracle.py
    evaluates the reference's sums tap by tap (pinned to the literal restatement and to the reference's fixtures by
    tests/test_fast_oracle_cpu.py), so a whole configuration takes seconds; one config is cached across the modes."""
    from oracle import fast_oracle as F
    if cfg in _full_cache:
        return _full_cache[cfg]
    _full_cache.clear()
    kind, kw, xs, seed = FULL_CFGS[cfg]
    rng = np.random.default_rng(seed)
    X = rng.standard_normal(xs, dtype=np.float32)
    ks, oc = kw["kernel_shape"], kw["out_ch"]
    W = O.glorot_uniform(ks + (xs[3], oc), rng=np.random.RandomState(seed)).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, 1, 1, oc))).astype(np.float32)
    if kind == "conv":
        args = (kw.get("stride", 1), kw["pad"], kw.get("dilation", 0))
        Y, Z = F.conv2d_layer_fwd(X.astype(np.float64), W.astype(np.float64), b.astype(np.float64), *args)
        dY = (Y + 0.5 * Y.std() * rng.standard_normal(Y.shape)).astype(np.float32)       # correlated with Y: see the adjoint check
        dX, dW, db = F.conv2d_layer_bwd(dY.astype(np.float64), X.astype(np.float64), Z, W.astype(np.float64), *args)
    else:
        args = (kw["stride"], kw["pad"])
        Y, Z = F.deconv2d_layer_fwd(X.astype(np.float64), W.astype(np.float64), b.astype(np.float64), *args)
        dY = (Y + 0.5 * Y.std() * rng.standard_normal(Y.shape)).astype(np.float32)
        dX, dW, db = F.deconv2d_layer_bwd(dY.astype(np.float64), X.astype(np.float64), Z, W.astype(np.float64), *args)
    out = dict(X=X, W=W, b=b, dY=dY, Y=Y.astype(np.float32), dX=dX.astype(np.float32), dW=dW, db=db)
    _full_cache[cfg] = out
    return out


def _dev_err(a, ref):
    """Normwise relative error of a device tensor against a host reference, computed on the device in float64."""
    import torch
    r = torch.from_numpy(np.ascontiguousarray(ref)).to(a.device).double().reshape(a.shape)
    return float(((a.double() - r).norm() / r.norm()).item())


@pytest.mark.parametrize("cfg", list(FULL_CFGS))
@pytest.mark.parametrize("mode", MODES)
def test_full_size_against_oracle(cuda_device, mode, cfg):
    """At the BASELINE.json sizes themselves: the WHOLE of Y, dX, dW and db against the float64 oracle (not a sample),
    so every work unit of the persistent kernels -- a CTA's 2nd .. 16th unit, ring wrap-arounds, the last partial
    wave -- is compared; plus the adjoint identities <Y, dY> = <X, dX> = <W, dW> + <b, db> with a tolerance relative
    to the inner product itself (dY is correlated with Y so that the inner product is O(|Y| |dY|))."""
    import torch
    pkg = npml()
    kind, kw, xs, _ = FULL_CFGS[cfg]
    c = _full_case(cfg)
    L = (pkg.Conv2D if kind == "conv" else pkg.Deconv2D)(mode=mode, **kw)
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    X = torch.from_numpy(c["X"]).cuda().to(dt)
    dY = torch.from_numpy(c["dY"]).cuda().to(dt)
    L.forward(X[:1], retain_derived=False)
    L.parameters["W"], L.parameters["b"] = c["W"], c["b"]
    Y = L.forward(X)
    dX = L.backward(dY)
    tol = TOL[mode]
    errs = {"Y": _dev_err(Y, c["Y"]), "dX": _dev_err(dX, c["dX"]), "dW": _dev_err(L.gradients["W"], c["dW"]),
            "db": _dev_err(L.gradients["b"], c["db"])}
    assert all(v <= tol for v in errs.values()), (cfg, mode, errs)
    # the checks have teeth: a zeroed (or half-zeroed) tensor is a relative error of 1 (0.7)
    half = dX.clone()
    half.view(-1)[half.numel() // 2:] = 0
    assert _dev_err(half, c["dX"]) > 0.5
    # adjoint identities, relative to the inner product
    lhs = _dot(Y, dY)
    rhs_x = _dot(X, dX) + _dot(L.parameters["b"].expand_as(Y), dY)
    rhs_w = _dot(L.parameters["W"], L.gradients["W"]) + _dot(L.parameters["b"], L.gradients["b"])
    assert abs(lhs) > 0.3 * float(Y.double().norm() * dY.double().norm())
    assert abs(lhs - rhs_x) <= tol * abs(lhs), (lhs, rhs_x)
    assert abs(lhs - rhs_w) <= tol * abs(lhs), (lhs, rhs_w)


@pytest.mark.parametrize("mode", MODES)
def test_full_size_skip_conv_module_against_oracle(cuda_device, mode):
    """cfg4 (SkipConnectionConvModule, N=256, 32x32x64->128, ReLU) at full size: Y, dX, every parameter gradient and
    the running statistics against the float64 oracle (modules/modules.py:905-984).  The two ReLU derivatives of the
    backward are evaluated at the device's own pre-activations (whose parity is asserted first): with 33.5 M elements
    even fp32 rounding flips a few dozen masks, and a flipped mask is an O(1) error in that element (see KINKED)."""
    import torch
    from oracle import fast_oracle as F
    pkg = npml()
    key = "cfg4-module"
    if key not in _full_cache:
        _full_cache.clear()
        rng = np.random.default_rng(1004)
        rs = np.random.RandomState(1004)
        X = rng.standard_normal((256, 32, 32, 64), dtype=np.float32)
        dY = rng.standard_normal((256, 32, 32, 128), dtype=np.float32)
        P = {}
        for t, (k, ci) in (("1", ((3, 3), 64)), ("2", ((3, 3), 128)), ("s", ((1, 1), 64))):
            P["W" + t] = O.glorot_uniform(k + (ci, 128), rng=rs).astype(np.float32).astype(np.float64)
            P["b" + t] = (0.1 * rng.standard_normal((1, 1, 1, 128))).astype(np.float32).astype(np.float64)
            P["g" + t] = rs.rand(128).astype(np.float32).astype(np.float64) + 0.25
            P["h" + t] = (0.1 * rng.standard_normal(128)).astype(np.float32).astype(np.float64)
            P["rm" + t], P["rv" + t] = np.zeros(128), np.ones(128)
        fwd = F.skip_conv_module_fwd(X.astype(np.float64), P, (3, 3), (3, 3), (1, 1), 1, 1, act="relu")
        _full_cache[key] = (X, dY, P, fwd)
    X, dY, P, fwd = _full_cache[key]
    M = pkg.SkipConnectionConvModule(out_ch1=128, out_ch2=128, kernel_shape1=(3, 3), kernel_shape2=(3, 3),
                                     kernel_shape_skip=(1, 1), pad1=1, pad2=1, act_fn=pkg.activations.ReLU(), mode=mode)
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    Xd, dYd = torch.from_numpy(X).cuda().to(dt), torch.from_numpy(dY).cuda().to(dt)
    M.forward(Xd[:2], retain_derived=False)
    for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
        conv.parameters["W"], conv.parameters["b"] = P["W" + t], P["b" + t]
        bn.parameters["scaler"], bn.parameters["intercept"] = P["g" + t], P["h" + t]
        bn.reset_running_stats()
    M.flush_gradients()
    Y = M.forward(Xd)
    # the fused tail keeps Y instead of the sum for ReLU (Y > 0 <=> sum > 0): either works as the mask's argument
    z1_dev, total_dev = M.conv1.derived_variables["Z"][0], M._tail[0]
    dX = M.backward(dYd)
    tol = TOL_BN[mode]
    bad = {}

    def chk(name, got, ref):
        e = _dev_err(got if isinstance(got, torch.Tensor) else torch.as_tensor(np.asarray(got)).cuda(), ref)
        if not e <= tol:
            bad[name] = e

    chk("Y", Y, fwd["Y"])
    chk("Z1", z1_dev, fwd["_z1"])
    chk("relu(sum)", total_dev, np.maximum(fwd["_total"], 0))
    for t, bn in (("1", M.batchnorm1), ("2", M.batchnorm2), ("s", M.batchnorm_skip)):
        chk("rm" + t, bn.parameters["running_mean"], fwd["rm" + t])
        chk("rv" + t, bn.parameters["running_var"], fwd["rv" + t])
    assert not bad, (mode, bad)
    ref = F.skip_conv_module_bwd(dict(fwd), X.astype(np.float64), dY.astype(np.float64), P, act="relu",
                                 z1=z1_dev.double().cpu().numpy(), total=total_dev.double().cpu().numpy())
    chk("dX", dX, ref["dX"])
    for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
        chk("dW" + t, conv.gradients["W"], ref["dW" + t])
        chk("dg" + t, bn.gradients["scaler"], ref["dg" + t])
    chk("dh2", M.batchnorm2.gradients["intercept"], ref["dh2"])
    chk("dhs", M.batchnorm_skip.gradients["intercept"], ref["dhs"])
    chk("db1", M.conv1.gradients["b"], ref["db1"])
    # dIntercept of batchnorm1 = sum over 262 144 pixels of dLdBn1 = dgrad(conv2) of a BatchNorm backward, whose
    # per-channel sums are zero by construction: only the image borders survive, the rest is cancellation (measured:
    # |dh1| is 0.2 of the summands' norm).  The rounding noise of such a sum scales with the norm of the SUMMANDS, so
    # the mode's tolerance is applied relative to ||dLdBn1||_F; a zeroed dh1 is still 10x over that bound in bf16.
    eps = TOL[mode]
    dh1 = M.batchnorm1.gradients["intercept"].double().cpu().numpy().ravel()
    assert np.linalg.norm(dh1 - ref["dh1"].ravel()) <= eps * np.linalg.norm(ref["dLdBn1"].ravel()), (mode, "dh1")
    assert not bad, (mode, bad)


# a persistent CTA's steady state (2nd .. 10th work unit: slab-ring phase flips, accumulator-ring wrap-around, the
# partial last unit) on a miniature, by capping the grid: NPML_MAX_CTAS is read when a call is planned
STEADY_CASES = [
    ("conv", (8, 28, 28, 64), dict(out_ch=64, kernel_shape=(3, 3), pad="same"), "relu"),             # cfg2 kernel, resident weights
    ("conv", (6, 20, 20, 128), dict(out_ch=128, kernel_shape=(3, 3), pad=1), "identity"),            # streamed weights, nkb = 2
    ("conv", (6, 24, 24, 64), dict(out_ch=128, kernel_shape=(3, 3), pad=1, stride=2), "tanh"),       # planes fwd / classes dX
    ("conv", (5, 28, 28, 128), dict(out_ch=256, kernel_shape=(3, 3), pad=0, stride=2, dilation=2), "identity"),   # cfg3 kernels
    ("deconv", (6, 16, 16, 256), dict(out_ch=128, kernel_shape=(4, 4), pad=1, stride=2), "identity"),  # cfg5 kernels
]


@pytest.mark.parametrize("ctas", [1, 3])
@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("case", range(len(STEADY_CASES)))
def test_persistent_steady_state_vs_oracle(cuda_device, monkeypatch, case, mode, ctas):
    pkg = npml()
    monkeypatch.setenv("NPML_MAX_CTAS", str(ctas))
    kind, xs, kw, act = STEADY_CASES[case]
    rng = np.random.default_rng(300 + case)
    L = (pkg.Conv2D if kind == "conv" else pkg.Deconv2D)(act_fn=make_act(pkg, act, 0, 0), mode=mode, **kw)
    ks, oc = kw["kernel_shape"], kw["out_ch"]
    X = _f32(rng, xs)
    W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
    b = _f32(rng, (1, 1, 1, oc), 0.1)
    L.forward(X[:1], retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    Y = L.forward(X)
    if kind == "conv":
        a = (kw.get("stride", 1), kw["pad"], kw.get("dilation", 0), act)
        Yr, Zr = O.conv2d_layer_fwd(X, W, b, *a)
    else:
        a = (kw["stride"], kw["pad"], act)
        Yr, Zr = O.deconv2d_layer_fwd(X, W, b, *a)
    tol = TOL[mode]
    close(Y, Yr, tol, "Y")
    dY = _f32(rng, Yr.shape)
    dX = L.backward(dY)
    if act in KINKED:
        close(L.derived_variables["Z"][0], Zr, tol, "Z")
        Zr = pkg.to_numpy(L.derived_variables["Z"][0])
    dXr, dWr, dbr = (O.conv2d_layer_bwd if kind == "conv" else O.deconv2d_layer_bwd)(dY, X, Zr, W, *a)
    close(dX, dXr, tol, "dX")
    close(L.gradients["W"], dWr, tol, "dW")
    close(L.gradients["b"], dbr, tol, "db")


ALL_ACTS = [("identity", 0, 0), ("affine", 0.7, -0.2), ("relu", 0, 0), ("leaky_relu", 0.25, 0), ("tanh", 0, 0),
            ("sigmoid", 0, 0), ("elu", 0.9, 0), ("selu", 0, 0), ("hard_sigmoid", 0, 0), ("softplus", 0, 0), ("gelu", 0, 0),
            ("exponential", 0, 0)]
# one shape per tensor-core epilogue: slab (stride 1), per-tile gather kernel (cfg3 geometry), Deconv2D on output classes
ACT_EPILOGUES = [("conv", (3, 10, 10, 64), dict(out_ch=64, kernel_shape=(3, 3), pad=1)),
                 ("conv", (2, 28, 28, 128), dict(out_ch=64, kernel_shape=(3, 3), pad=0, stride=2, dilation=2)),
                 ("deconv", (2, 8, 8, 64), dict(out_ch=64, kernel_shape=(4, 4), pad=1, stride=2))]


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("epi", range(len(ACT_EPILOGUES)))
@pytest.mark.parametrize("act", ALL_ACTS, ids=[a[0] for a in ALL_ACTS])
def test_every_activation_on_tensor_core_epilogues(cuda_device, act, epi, mode):
    """All twelve activation functors as the epilogue (fn) and the dZ prologue (grad) of every tcgen05 kernel family,
    on 64-channel shapes (the reference fixtures have <= 5 channels and never reach these kernels)."""
    pkg = npml()
    name, p0, p1 = act
    kind, xs, kw = ACT_EPILOGUES[epi]
    rng = np.random.default_rng(400 + epi)
    L = (pkg.Conv2D if kind == "conv" else pkg.Deconv2D)(act_fn=make_act(pkg, name, p0, p1), mode=mode, **kw)
    ks, oc = kw["kernel_shape"], kw["out_ch"]
    X = _f32(rng, xs)
    W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
    b = _f32(rng, (1, 1, 1, oc), 0.1)
    L.forward(X[:1], retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    assert L._uses_tc(xs)
    Y = L.forward(X)
    if kind == "conv":
        a = (kw.get("stride", 1), kw["pad"], kw.get("dilation", 0), name, p0, p1)
        Yr, Zr = O.conv2d_layer_fwd(X, W, b, *a)
    else:
        a = (kw["stride"], kw["pad"], name, p0, p1)
        Yr, Zr = O.deconv2d_layer_fwd(X, W, b, *a)
    tol = TOL[mode]
    close(Y, Yr, tol, name + " Y")
    close(L.derived_variables["Z"][0], Zr, tol, name + " Z")
    dY = _f32(rng, Yr.shape)
    dX = L.backward(dY)
    Zd = pkg.to_numpy(L.derived_variables["Z"][0])          # backward evaluated at the device's own pre-activations
    dXr, dWr, dbr = (O.conv2d_layer_bwd if kind == "conv" else O.deconv2d_layer_bwd)(dY, X, Zd, W, *a)
    close(dX, dXr, tol, name + " dX")
    close(L.gradients["W"], dWr, tol, name + " dW")
    close(L.gradients["b"], dbr, tol, name + " db")


def test_wgrad_is_bitwise_deterministic(cuda_device):
    """dW / db come from a fixed-order split-K reduction (no atomics): two runs on the same inputs are bit-identical,
    in every mode and on every wgrad kernel family (layers/layers.py:3106-3110)."""
    import torch
    pkg = npml()
    shapes = [("conv", (64, 56, 56, 64), dict(out_ch=64, kernel_shape=(3, 3), pad="same")),
              ("conv", (32, 28, 28, 128), dict(out_ch=256, kernel_shape=(3, 3), pad=0, stride=2, dilation=2)),
              ("deconv", (32, 16, 16, 256), dict(out_ch=128, kernel_shape=(4, 4), pad=1, stride=2)),
              ("conv", (32, 28, 28, 1), dict(out_ch=32, kernel_shape=(3, 3), pad="same"))]
    for mode in MODES:
        for kind, xs, kw in shapes:
            torch.manual_seed(3)
            dt = torch.bfloat16 if mode == "bf16" else torch.float32
            X = torch.randn(xs, device="cuda").to(dt)
            L = (pkg.Conv2D if kind == "conv" else pkg.Deconv2D)(mode=mode, **kw)
            np.random.seed(5)
            Y = L.forward(X)
            dY = torch.randn(Y.shape, device="cuda").to(dt)
            runs = []
            for _ in range(3):
                L.gradients["W"].zero_(); L.gradients["b"].zero_()
                dX = L.backward(dY)
                runs.append((L.gradients["W"].clone(), L.gradients["b"].clone(), dX.clone()))
            for r in runs[1:]:
                for a, b_ in zip(runs[0], r):
                    assert torch.equal(a, b_), (mode, kind, xs)


# ------------------------------------------------------------------------------------------------
#  optimizer step after the path: device kernels vs the reference's golden vectors
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_optimizers_match_reference_golden(cuda_device):
    import torch
    pkg = npml()
    Op, Sc = pkg.optimizers, pkg.schedulers
    z, cases = load_golden("optimizers.npz")
    dev = torch.device("cuda:0")
    for c in cases:
        cid = c["id"]
        kwargs = dict(c["kwargs"])
        if c["sched"] is not None:
            sk = {k: v for k, v in c["sched"].items() if k != "kind"}
            kwargs["lr_scheduler"] = (Sc.ExponentialScheduler if c["sched"]["kind"] == "exponential"
                                      else Sc.NoamScheduler)(**sk)
        opt = getattr(Op, c["kind"])(**kwargs)
        param = torch.tensor(z[cid + "/param0"], dtype=torch.float32, device=dev)
        for step in range(4):
            grad = torch.tensor(z["{}/grad{}".format(cid, step)], dtype=torch.float32, device=dev)
            opt.step()
            out = opt.update(param, grad, "W")
            assert out.data_ptr() == param.data_ptr()                 # in place
            close(param, z["{}/param{}".format(cid, step + 1)], 2e-6, "{} param step {}".format(cid, step + 1))
            cache = opt.cache["W"]
            if c["kind"] == "Adam":
                assert cache["t"] == step + 1
                close(cache["mean"], z["{}/mean{}".format(cid, step + 1)], 2e-6, cid + " mean")
                close(cache["var"], z["{}/var{}".format(cid, step + 1)], 2e-6, cid + " var")
            else:
                close(cache, z["{}/cache{}".format(cid, step + 1)], 2e-6, cid + " cache")


@pytest.mark.gpu
def test_layer_update_with_adam_matches_oracle(cuda_device):
    """Conv2D.update(): optimizer.step(), one Adam update per parameter, gradients flushed (layers.py:76-86)."""
    from oracle import opt_oracle as P
    pkg = npml()
    rng = np.random.default_rng(31)
    X, dY = _f32(rng, (2, 6, 6, 8)), _f32(rng, (2, 6, 6, 8))
    L = pkg.Conv2D(8, (3, 3), pad="same", optimizer="Adam(lr=0.01, decay1=0.9, decay2=0.999, eps=1e-07, clip_norm=None)")
    L.forward(X)
    W0, b0 = pkg.to_numpy(L.parameters["W"]), pkg.to_numpy(L.parameters["b"])
    L.backward(dY)
    gW, gb = pkg.to_numpy(L.gradients["W"]), pkg.to_numpy(L.gradients["b"])
    L.update()
    Wr, _, _, _ = P.adam_step(W0, gW, 0 * W0, 0 * W0, 0, 0.01, 0.9, 0.999, 1e-7)
    br, _, _, _ = P.adam_step(b0, gb, 0 * b0, 0 * b0, 0, 0.01, 0.9, 0.999, 1e-7)
    close(L.parameters["W"], Wr, 1e-5, "W after update")
    close(L.parameters["b"], br, 1e-5, "b after update")
    assert float(L.gradients["W"].abs().sum()) == 0.0 and L.X == []
    assert L.hyperparameters["optimizer"]["hyperparameters"]["id"] == "Adam"


# ------------------------------------------------------------------------------------------------
#  callers either side of the path: Conv1D, Pool2D / Flatten, identity skip module
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", MODES)
def test_conv1d_layer_golden(cuda_device, mode):
    pkg = npml()
    z, cases = load_golden("conv1d_layer.npz")
    tol = TOL[mode]
    for c in cases:
        if mode != "fp32" and c["act"] in KINKED:
            continue                                   # mask flips next to the kink, see KINKED above
        i, pad = c["id"], to_pad(c["pad"])
        layer = pkg.Conv1D(out_ch=c["out_ch"], kernel_width=c["kernel_width"], pad=pad, stride=c["stride"],
                           dilation=c["dilation"], act_fn=make_act(pkg, c["act"], c["p0"], c["p1"]), mode=mode)
        X = z[i + "/X"]
        layer.forward(X, retain_derived=False)
        assert tuple(layer.parameters["W"].shape) == z[i + "/W"].shape and tuple(layer.parameters["b"].shape) == z[i + "/b"].shape
        layer.parameters["W"], layer.parameters["b"] = z[i + "/W"], z[i + "/b"]
        Y = layer.forward(X)
        close(Y, z[i + "/Y"], tol, i + " Y")
        close(layer.derived_variables["Z"][0], z[i + "/Z"], tol, i + " Z")
        assert tuple(layer.X[0].shape) == X.shape
        assert layer.derived_variables["out_rows"][0] == z[i + "/Z"].shape[1]       # l_out (layers.py:2749)
        assert layer.derived_variables["out_cols"][0] == z[i + "/Z"].shape[2]       # out_ch (layers.py:2750)
        if c["two_calls"]:
            close(layer.forward(z[i + "/X2"]), z[i + "/Y2"], tol, i + " Y2")
            dXs = layer.backward([z[i + "/dLdy"], z[i + "/dLdy2"]])
            close(dXs[0], z[i + "/dX"], tol, i + " dX")
            close(dXs[1], z[i + "/dX2"], tol, i + " dX2")
        else:
            close(layer.backward(z[i + "/dLdy"]), z[i + "/dX"], tol, i + " dX")
        close(layer.gradients["W"], z[i + "/dW"], tol, i + " dW")
        close(layer.gradients["b"], z[i + "/db"], tol, i + " db")
        if mode == "fp32":
            close(pkg.conv1D(X, z[i + "/W"], c["stride"], pad, c["dilation"]), z[i + "/conv1D"], tol, i + " conv1D")
            Xp, p2 = pkg.pad1D(X, pad, c["kernel_width"], c["stride"], c["dilation"])
            assert list(p2) == list(z[i + "/pad"]) and Xp.shape[1] == X.shape[1] + sum(p2)


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("case", [
    # n, l, c_in, c_out, kw, pad, stride, dilation
    (4, 200, 64, 64, 3, "causal", 1, 3),           # Wavenet residual shape: slab kernels on a one-row volume
    (2, 1000, 32, 48, 2, "causal", 1, 7),          # long sequence (row wider than one slab)
    (3, 128, 64, 128, 5, "same", 2, 0),            # strided
])
def test_conv1d_tensor_core_vs_oracle(cuda_device, mode, case):
    from oracle import next_oracle as NO
    pkg = npml()
    n, l, ci, co, kw, pad, s, d = case
    rng = np.random.default_rng(7 + l)
    X, W, b = _f32(rng, (n, l, ci)), _f32(rng, (kw, ci, co), (kw * ci) ** -0.5), _f32(rng, (1, 1, co), 0.1)
    layer = pkg.Conv1D(co, kw, pad=pad, stride=s, dilation=d, act_fn="tanh", mode=mode)
    layer.forward(X, retain_derived=False)
    layer.parameters["W"], layer.parameters["b"] = W, b
    Y = layer.forward(X)
    Yr, Zr = NO.conv1d_layer_fwd(X, W, b, s, pad, d, "tanh")
    dY = _f32(rng, Yr.shape)
    dX = layer.backward(dY)
    dXr, dWr, dbr = NO.conv1d_layer_bwd(dY, X, Zr, W, s, pad, d, "tanh")
    tol = TOL[mode]
    close(Y, Yr, tol, "Y")
    close(dX, dXr, tol, "dX")
    close(layer.gradients["W"], dWr, tol, "dW")
    close(layer.gradients["b"], dbr, tol, "db")
    assert layer._uses_tc((n, 1, l, ci))


def test_pool2d_flatten_golden(cuda_device):
    import torch
    pkg = npml()
    z, cases = load_golden("pool2d.npz")
    for c in cases:
        i = c["id"]
        if c.get("flatten"):
            F = pkg.Flatten(keep_dim=c["keep_dim"])
            Xd = torch.tensor(z[i + "/X"], dtype=torch.float32, device="cuda")
            Y = F.forward(Xd)
            assert Y.data_ptr() == Xd.data_ptr() and np.array_equal(Y.cpu().numpy(), z[i + "/Y"].astype(np.float32))
            assert np.array_equal(F.backward(z[i + "/dLdy"]), z[i + "/dX"])
            continue
        P = pkg.Pool2D(kernel_shape=tuple(c["kernel_shape"]), stride=c["stride"], pad=to_pad(c["pad"]), mode=c["mode"])
        Y = P.forward(z[i + "/X"])
        close(Y, z[i + "/Y"], 0.0 if c["mode"] == "max" else 1e-6, i + " Y")       # max is a selection: exact
        assert P.derived_variables["out_rows"][0] == z[i + "/Y"].shape[1]
        if c["backward"]:
            close(P.backward(z[i + "/dLdY"]), z[i + "/dX"], 1e-6, i + " dX")


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("mode", ["max", "average"])
def test_pool2d_vs_oracle_vae_shape(cuda_device, dtype, mode):
    """The pooling step of the reference's VAE encoder (models/vae.py:142-189) on device tensors, both element types."""
    import torch
    from oracle import next_oracle as NO
    pkg = npml()
    rng = np.random.default_rng(41)
    td = torch.float32 if dtype == "fp32" else torch.bfloat16
    Xd = torch.tensor(_f32(rng, (8, 28, 28, 32)), dtype=torch.float32, device="cuda").to(td)
    X = Xd.float().cpu().numpy().astype(np.float64)                  # the oracle sees the stored (rounded) values
    for ks, s, pad in (((2, 2), 2, 0), ((3, 3), 2, 1), ((3, 2), 1, (1, 0))):
        P = pkg.Pool2D(kernel_shape=ks, stride=s, pad=pad, mode=mode)
        Y = P.forward(Xd)
        Yr = NO.pool2d_fwd(X, ks, s, pad, mode)
        assert Y.dtype == td
        close(Y.float(), Yr, 0.0 if mode == "max" else (1e-6 if dtype == "fp32" else 4e-3), "Y")
        dYd = torch.tensor(_f32(rng, Yr.shape), dtype=torch.float32, device="cuda").to(td)
        dXr = NO.pool2d_bwd(dYd.float().cpu().numpy().astype(np.float64), X, ks, s, pad, mode)
        close(P.backward(dYd).float(), dXr, 1e-6 if dtype == "fp32" else 4e-3, "dX")


@pytest.mark.parametrize("fuse", [True, False], ids=["fused-tail", "layer-by-layer"])
@pytest.mark.parametrize("mode", MODES)
def test_skip_identity_module_golden(cuda_device, mode, fuse):
    pkg = npml()
    z, cases = load_golden("skip_identity_module.npz")
    tol = TOL_BN[mode]
    for c in cases:
        if mode != "fp32" and c["act"] in KINKED:
            continue
        i = c["id"]
        act = None if c["act"] == "identity" else make_act(pkg, c["act"], c["p0"], c["p1"])
        M = pkg.SkipConnectionIdentityModule(out_ch=c["out_ch"], kernel_shape1=tuple(c["kernel_shape1"]),
                                             kernel_shape2=tuple(c["kernel_shape2"]), act_fn=act, mode=mode)
        M.fuse_training = fuse
        X = z[i + "/X"]
        M.forward(X, retain_derived=False)
        for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2)):
            conv.parameters["W"], conv.parameters["b"] = z["{}/W{}".format(i, t)], z["{}/b{}".format(i, t)]
            bn.parameters["scaler"], bn.parameters["intercept"] = z["{}/g{}".format(i, t)], z["{}/h{}".format(i, t)]
            bn.reset_running_stats()
        M.flush_gradients()
        Y = M.forward(X)
        dX = M.backward(z[i + "/dLdY"])
        close(Y, z[i + "/Y"], tol, i + " Y")
        close(dX, z[i + "/dX"], tol, i + " dX")
        dv = M.derived_variables
        for k in ["conv1_out", "conv2_out", "batchnorm1_out", "batchnorm2_out", "dLdBn1", "dLdBn2", "dLdConv1", "dLdConv2"]:
            close(dv[k], z["{}/{}".format(i, k)], tol, i + " " + k)
        close(dv["dLdAdd3_X"], z[i + "/dX"], tol, i + " dLdAdd3_X")
        for t, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2)):
            close(conv.gradients["W"], z["{}/dW{}".format(i, t)], tol, i + " dW" + t)
            close(bn.gradients["scaler"], z["{}/dg{}".format(i, t)], tol, i + " dg" + t)
            # BN1's intercept feeds a 1x1 conv2 + BN2 in case m1: its true gradient is 0 (pure cancellation), so it is
            # compared absolutely, on the scale of the scaler gradient
            dh_ref = z["{}/dh{}".format(i, t)]
            dh = pkg.to_numpy(bn.gradients["intercept"])
            scale = max(np.linalg.norm(z["{}/dg{}".format(i, t)].ravel()), np.linalg.norm(dh_ref.ravel()), 1.0)
            assert np.linalg.norm((dh - dh_ref).ravel()) <= tol * scale, i + " dh" + t
            close(bn.parameters["running_mean"], z["{}/rm{}".format(i, t)], tol, i + " rm" + t)
            close(bn.parameters["running_var"], z["{}/rv{}".format(i, t)], tol, i + " rv" + t)
        assert M.hyperparameters["component_ids"] == ["conv1", "batchnorm1", "conv2", "batchnorm2", "add3"]


@pytest.mark.parametrize("fold", ["2", "3"])
@pytest.mark.parametrize("mode", ["bf16", "tf32"])
def test_row_of_taps_slab_kernel_optin(cuda_device, monkeypatch, mode, fold):
    """csrc/tc_slab3_conv.cu (NPML_SLAB3=2|3: column taps of a filter row folded into one MMA, epilogue row shifts):
    forward and dX of a 3x3 / 64-channel layer, of a dilated one, and of a causal Conv1D against the oracle."""
    from oracle import next_oracle as NO
    monkeypatch.setenv("NPML_SLAB3", fold)
    pkg = npml()
    rng = np.random.default_rng(5)
    tol = TOL[mode]
    for (n, h, w, ci, co, pad, d) in ((3, 20, 23, 64, 64, "same", 0), (2, 17, 30, 32, 48, 2, 1)):
        X, W = _f32(rng, (n, h, w, ci)), _f32(rng, (3, 3, ci, co), (9 * ci) ** -0.5)
        b = _f32(rng, (1, 1, 1, co), 0.1)
        L = pkg.Conv2D(co, (3, 3), pad=pad, dilation=d, act_fn="tanh", mode=mode)
        L.forward(X, retain_derived=False)
        L.parameters["W"], L.parameters["b"] = W, b
        Y = L.forward(X)
        Yr, Zr = O.conv2d_layer_fwd(X, W, b, 1, pad, d, "tanh")
        dY = _f32(rng, Yr.shape)
        dX = L.backward(dY)
        dXr, dWr, dbr = O.conv2d_layer_bwd(dY, X, Zr, W, 1, pad, d, "tanh")
        close(Y, Yr, tol, "Y")
        close(dX, dXr, tol, "dX")
        close(L.gradients["W"], dWr, tol, "dW")
    X, W, b = _f32(rng, (2, 300, 64)), _f32(rng, (3, 64, 64), 192 ** -0.5), _f32(rng, (1, 1, 64), 0.1)
    L = pkg.Conv1D(64, 3, pad="causal", dilation=1, mode=mode)
    L.forward(X, retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    Y = L.forward(X)
    Yr, Zr = NO.conv1d_layer_fwd(X, W, b, 1, "causal", 1)
    dY = _f32(rng, Yr.shape)
    close(Y, Yr, tol, "conv1d Y")
    close(L.backward(dY), NO.conv1d_layer_bwd(dY, X, Zr, W, 1, "causal", 1)[0], tol, "conv1d dX")


# stride-1 shapes csrc/tc_ws_conv.cu takes in bf16 mode: pair mode (out_ch <= 64: two taps per MMA, even slot distance),
# single mode (out_ch in 65..128), two K blocks, all-paired 4x4 rows, 25 taps, dilation 3 (delta = 4), 1x1, channel tails
WS_CASES = [0, 3, 4, 5, 7, 9, 10, 12, 13, 15, 21]


@pytest.mark.parametrize("case", WS_CASES)
def test_weight_stationary_kernel_optin(cuda_device, monkeypatch, case):
    """csrc/tc_ws_conv.cu (NPML_WS=1: weights resident in tensor memory, pixels along N, tap pairs stacked along M):
    forward (with the folded frozen BatchNorm2D too) and dX against the oracle, and bit-for-bit the same dW path."""
    monkeypatch.setenv("NPML_WS", "1")
    pkg = npml()
    xs, oc, ks, pad, s, d, act = CONV_CASES[case]
    assert s == 1
    rng = np.random.default_rng(900 + case)
    p0 = 0.25 if act == "leaky_relu" else 0.0
    L = pkg.Conv2D(out_ch=oc, kernel_shape=ks, pad=pad, stride=s, dilation=d, act_fn=make_act(pkg, act, p0, 0.0),
                   mode="bf16")
    X = _f32(rng, xs)
    W = _f32(rng, ks + (xs[3], oc), 1.0 / np.sqrt(ks[0] * ks[1] * xs[3]))
    b = _f32(rng, (1, 1, 1, oc), 0.1)
    L.forward(X, retain_derived=False)
    L.parameters["W"], L.parameters["b"] = W, b
    tol = TOL["bf16"]
    Y = L.forward(X)
    Yr, Zr = O.conv2d_layer_fwd(X, W, b, s, pad, d, act, p0, 0.0)
    close(Y, Yr, tol, "Y")
    close(L.derived_variables["Z"][0], Zr, tol, "Z")
    dY = _f32(rng, Yr.shape)
    dX = L.backward(dY)
    Zd = pkg.to_numpy(L.derived_variables["Z"][0]) if act in KINKED else Zr
    dXr, dWr, dbr = O.conv2d_layer_bwd(dY, X, Zd, W, s, pad, d, act, p0, 0.0)
    close(dX, dXr, tol, "dX")
    close(L.gradients["W"], dWr, tol, "dW")
    bn, P = _random_bn(pkg, rng, oc)
    ref, _, _ = O.bn2d_fwd(Yr, P["scaler"], P["intercept"], P["running_mean"], P["running_var"], use_batch_stats=False)
    close(L.forward(X, retain_derived=False, fused_bn=bn), ref, tol, "fused BN")


def test_multiply_fc_golden(cuda_device):
    pkg = npml()
    z, cases = load_golden("wavenet_vae.npz")
    for c in cases:
        i = c["id"]
        if c["kind"] == "multiply":
            M = pkg.Multiply(act_fn=make_act(pkg, c["act"], 0, 0))
            Xs = [z["{}/X{}".format(i, j)] for j in range(c["n_in"])]
            close(M.forward(Xs), z[i + "/Y"], 2e-6, i + " Y")
            g = M.backward(z[i + "/dLdY"])
            assert len(g) == c["n_in"]
            for j in range(c["n_in"]):
                close(g[j], z["{}/dX{}".format(i, j)], 2e-6, "{} dX{}".format(i, j))
        elif c["kind"] == "fc":
            F = pkg.FullyConnected(n_out=c["n_out"], act_fn=make_act(pkg, c["act"], 0, 0))
            X = z[i + "/X"]
            F.forward(X, retain_derived=False)
            assert tuple(F.parameters["W"].shape) == z[i + "/W"].shape and tuple(F.parameters["b"].shape) == z[i + "/b"].shape
            F.parameters["W"], F.parameters["b"] = z[i + "/W"], z[i + "/b"]
            close(F.forward(X), z[i + "/Y"], TOL["fp32"], i + " Y")
            assert list(F.derived_variables) == ["Z"] and tuple(F.X[0].shape) == X.shape
            close(F.backward(z[i + "/dLdy"]), z[i + "/dX"], TOL["fp32"], i + " dX")
            close(F.gradients["W"], z[i + "/dW"], TOL["fp32"], i + " dW")
            close(F.gradients["b"], z[i + "/db"], TOL["fp32"], i + " db")


@pytest.mark.parametrize("mode", MODES)
def test_wavenet_module_golden(cuda_device, mode):
    pkg = npml()
    z, cases = load_golden("wavenet_vae.npz")
    tol = TOL[mode] * (1 if mode == "fp32" else 2)     # two chained convolutions and a gate product
    for c in cases:
        if c["kind"] != "wavenet":
            continue
        i = c["id"]
        M = pkg.WavenetResidualModule(ch_residual=c["ch_residual"], ch_dilation=c["ch_dilation"], dilation=c["dilation"],
                                      kernel_width=2, mode=mode)
        M.forward(z[i + "/X_main"], z[i + "/X_skip"])
        for t, conv in (("d", M.conv_dilation), ("1", M.conv_1x1)):
            conv.parameters["W"], conv.parameters["b"] = z["{}/W{}".format(i, t)], z["{}/b{}".format(i, t)]
        M.flush_gradients()
        Ym, Ys = M.forward(z[i + "/X_main"], z[i + "/X_skip"])
        close(Ym, z[i + "/Y_main"], tol, i + " Y_main")
        close(Ys, z[i + "/Y_skip"], tol, i + " Y_skip")
        dXm, dXs = M.backward(z[i + "/dY_skip"], z[i + "/dY_main"] if c["with_main"] else None)
        close(dXm, z[i + "/dX_main"], tol, i + " dX_main")
        close(dXs, z[i + "/dX_skip"], tol, i + " dX_skip")
        dv = M.derived_variables
        for k in ("conv_dilation_out", "tanh_out", "sigm_out", "multiply_gate_out", "conv_1x1_out", "dLdTanh", "dLdSigmoid",
                  "dLdConv_1x1", "dLdMultiply", "dLdConv_dilation"):
            close(dv[k], z["{}/{}".format(i, k)], tol, i + " " + k)
        close(M.conv_dilation.gradients["W"], z[i + "/dWd"], tol, i + " dWd")
        close(M.conv_dilation.gradients["b"], z[i + "/dbd"], tol, i + " dbd")
        close(M.conv_1x1.gradients["W"], z[i + "/dW1"], tol, i + " dW1")
        close(M.conv_1x1.gradients["b"], z[i + "/db1"], tol, i + " db1")


def test_vae_encoder_chain_golden(cuda_device):
    """Conv -> MaxPool -> Conv -> MaxPool -> Flatten -> FC -> FC (models/vae.py:142-189) end to end on the device,
    forward and backward, against the reference's own layers (fixture) -- device tensors flow between the layers."""
    import torch
    pkg = npml()
    z, _ = load_golden("wavenet_vae.npz")
    enc = [pkg.Conv2D(act_fn="relu", pad="same", out_ch=8, stride=1, kernel_shape=(5, 5)),
           pkg.Pool2D(mode="max", pad=0, stride=2, kernel_shape=(2, 2)),
           pkg.Conv2D(act_fn="relu", pad="same", out_ch=16, stride=1, kernel_shape=(5, 5)),
           pkg.Pool2D(mode="max", pad=0, stride=2, kernel_shape=(2, 2)),
           pkg.Flatten(), pkg.FullyConnected(n_out=24, act_fn="relu"), pkg.FullyConnected(n_out=10)]
    X = torch.tensor(z["vae/X"], dtype=torch.float32, device="cuda")
    out = X
    for L in enc:
        out = L.forward(out, retain_derived=False)
    k = 0
    for L in enc:
        if L.parameters:
            L.parameters["W"], L.parameters["b"] = z["vae/W{}".format(k)], z["vae/b{}".format(k)]
            k += 1
    out = X
    for L in enc:
        out = L.forward(out)
    assert out.is_cuda
    close(out, z["vae/Y"], 1e-5, "Y")
    g = torch.tensor(z["vae/dLdY"], dtype=torch.float32, device="cuda")
    for L in reversed(enc):
        g = L.backward(g)
    close(g, z["vae/dX"], 1e-5, "dX")
    k = 0
    for L in enc:
        if L.parameters:
            close(L.gradients["W"], z["vae/dW{}".format(k)], 1e-5, "dW{}".format(k))
            close(L.gradients["b"], z["vae/db{}".format(k)], 1e-5, "db{}".format(k))
            k += 1
