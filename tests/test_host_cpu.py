"""CPU-only checks of the host side: geometry / error conventions against the reference's golden
vectors, and that libnpmlconv.so loads and exports every symbol include/npml_conv.h declares.
No kernels are launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, to_pad, npml


def test_library_exports_every_declared_symbol():
    pkg = npml()
    pkg.build()
    lib = pkg._lib.load()
    hdr = open(os.path.join(ROOT, "include", "npml_conv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(npml_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(pkg._lib.SIGNATURES), declared ^ set(pkg._lib.SIGNATURES)
    assert lib.npml_version() >= 100


def test_desc_struct_matches_header():
    pkg = npml()
    assert ctypes.sizeof(pkg._lib.ConvDesc) == 21 * 4
    names = [f[0] for f in pkg._lib.ConvDesc._fields_]
    hdr = open(os.path.join(ROOT, "include", "npml_conv.h")).read()
    body = hdr[hdr.index("typedef struct npml_conv_desc"):hdr.index("} npml_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body):
        fields += [v.strip() for v in decl.split(",")]
    assert fields == names


def test_workspace_query_and_tc_predicate_without_gpu():
    pkg = npml()
    U, L = pkg.utils, pkg._lib
    d = U.make_desc(256, 56, 56, 64, 64, 3, 3, 1, 1, (1, 1, 1, 1), 56, 56, "bf16")
    lib = L.load()
    for op in range(9):
        assert lib.npml_workspace_bytes(d, op) >= 256
    assert lib.npml_uses_tensor_cores(d, L.OP_CONV_FPROP) in (0, 1)
    d1 = U.make_desc(32, 28, 28, 1, 32, 3, 3, 1, 1, (1, 1, 1, 1), 28, 28, "bf16")
    assert lib.npml_uses_tensor_cores(d1, L.OP_CONV_FPROP) == 0      # in_ch = 1 -> CUDA-core path
    # the one-launch backward (csrc/thin_bwd.cu) takes thin stride-1 'same' shapes only; its planner runs without a GPU
    assert lib.npml_conv2d_bwd_is_fused(d1) == 1                      # cfg1
    assert lib.npml_workspace_bytes(d1, L.OP_CONV_BWD) >= 256 + (3 * 3 * 1 + 1) * 32 * 4
    assert lib.npml_conv2d_bwd_is_fused(d) == 0                       # cfg2: 64 input channels
    for bad in (U.make_desc(32, 28, 28, 1, 32, 3, 3, 2, 1, (1, 1, 1, 1), 14, 14, "bf16"),      # stride 2
                U.make_desc(32, 28, 28, 1, 32, 3, 3, 1, 1, (0, 0, 0, 0), 26, 26, "bf16"),      # 'valid': output != input size
                U.make_desc(32, 28, 28, 1, 24, 3, 3, 1, 1, (1, 1, 1, 1), 28, 28, "bf16"),      # 6 lanes per pixel: not a power of two
                U.make_desc(32, 28, 28, 8, 32, 3, 3, 1, 1, (1, 1, 1, 1), 28, 28, "bf16")):     # in_ch > 4
        assert lib.npml_conv2d_bwd_is_fused(bad) == 0


def test_geometry_against_reference_golden():
    U = npml().utils
    z, cases = load_golden("utils.npz")
    for c in cases:
        if c["kind"] == "calc_pad_dims_2D":
            for it in c["items"]:
                got = U.calc_pad_dims_2D(tuple(it["X_shape"]), tuple(it["out_dim"]), tuple(it["kernel_shape"]),
                                         it["stride"], it["dilation"])
                assert list(got) == it["out"]
        elif c["kind"] == "pad2D":
            ks = tuple(c["kernel_shape"]) if c["kernel_shape"] else None
            p4 = U.resolve_pad_2D(z[c["id"] + "/X"].shape, to_pad(c["pad"]), ks, c["stride"], c["dilation"])
            assert list(p4) == c["p4"]
        elif c["kind"] == "conv":
            X, W = z[c["id"] + "/X"], z[c["id"] + "/W"]
            p4 = U.resolve_pad_2D(X.shape, to_pad(c["pad"]), W.shape[:2], c["stride"], c["dilation"])
            assert list(p4) == c["p4"]
            assert U.calc_conv_out_dims(X.shape, W.shape, c["stride"], p4, c["dilation"]) == z[c["id"] + "/Z"].shape
            oh, ow = U.conv_out_hw(X.shape[1], X.shape[2], W.shape[0], W.shape[1], p4, c["stride"], c["dilation"])
            assert (oh, ow) == z[c["id"] + "/Z"].shape[1:3]
        elif c["kind"] == "deconv":
            X, W = z[c["id"] + "/X"], z[c["id"] + "/W"]
            p4, oh, ow = U.deconv_geometry(X.shape[1], X.shape[2], W.shape[0], W.shape[1], to_pad(c["pad"]), c["stride"])
            assert (oh, ow) == z[c["id"] + "/Z"].shape[1:3]
            assert oh == c["stride"] * (X.shape[1] - 1) + W.shape[0] - p4[0] - p4[1]


def test_benchmark_config_geometry():
    U = npml().utils
    assert U.resolve_pad_2D((256, 56, 56, 64), "same", (3, 3), 1, 0) == (1, 1, 1, 1)              # cfg2
    assert U.resolve_pad_2D((128, 28, 28, 128), "same", (3, 3), 2, 2) == (16, 17, 16, 17)         # cfg3 'same'
    assert U.conv_out_hw(28, 28, 3, 3, (0, 0, 0, 0), 2, 2) == (11, 11)                            # cfg3
    assert U.deconv_geometry(16, 16, 4, 4, 1, 2) == ((1, 1, 1, 1), 32, 32)                        # cfg5


def test_error_conventions():
    U = npml().utils
    with pytest.raises(ValueError):
        U.calc_pad_dims_2D([1, 3, 3, 1], (3, 3), (3, 3), 1)          # utils.py:77-87
    with pytest.raises(ValueError):
        U.calc_pad_dims_2D((1, 9, 9, 1), (3, 3), (3, 3), 1)          # negative pad, utils.py:116-119
    with pytest.raises(AssertionError):
        U.calc_pad_dims_2D((1, 8, 8, 1), (3, 3), (3, 3), 1)          # unattainable, utils.py:107-108
    with pytest.raises(ValueError):
        U.conv_out_hw(3, 3, 5, 5, (0, 0, 0, 0), 1, 0)                # kernel > padded input, utils.py:463-467
    with pytest.raises(TypeError):
        U.col2im(np.zeros((9, 1)), (1, 3, 3, 1), (3, 3, 1, 1), 1, 1)  # pad must be a 4-tuple, utils.py:577-578


def test_activation_strings_and_initializer():
    A = npml().activations
    assert str(A.ReLU()) == "ReLU" and str(A.LeakyReLU(0.3)) == "Leaky ReLU(alpha=0.3)"
    assert str(A.Affine(2, 1)) == "Affine(slope=2, intercept=1)"
    init = A.ActivationInitializer
    assert isinstance(init(None)(), A.Identity)
    assert isinstance(init("relu")(), A.ReLU)
    a = init("Affine(slope=0.5, intercept=2.0)")()
    assert (a.p0, a.p1) == (0.5, 2.0)
    assert init("Leaky ReLU(alpha=0.25)")().p0 == 0.25
    with pytest.raises(ValueError):
        init("swish")()


def test_weight_init_consumes_np_random_like_reference():
    pkg = npml()
    from oracle import conv_oracle as O
    np.random.seed(1002)
    ours = pkg.layers.WeightInitializer("Identity", "glorot_uniform")((3, 3, 64, 64))
    np.random.seed(1002)
    ref = O.glorot_uniform((3, 3, 64, 64))
    assert np.array_equal(ours, ref)


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg = npml()
    with pytest.raises(RuntimeError):
        pkg.conv2D(np.zeros((1, 4, 4, 1)), np.zeros((3, 3, 1, 1)), 1, 0)
    with pytest.raises(RuntimeError):
        pkg.Conv2D(2, (3, 3)).forward(np.zeros((1, 4, 4, 1)))


def test_shard_batch():
    D = npml().dist
    for n, g in [(256, 8), (128, 4), (32, 2), (10, 4)]:
        spans = [D.shard_batch(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_1d_geometry_and_next_layer_host_logic():
    """Host side of the SURVEY 8f rows: 1-D pads against the reference's values, Pool2D / FullyConnected / Flatten
    bookkeeping that needs no kernel, and that the drop-ins fail loudly without a GPU."""
    import torch
    pkg = npml()
    U = pkg.utils
    z, cases = load_golden("conv1d_layer.npz")
    for c in cases:
        X = z[c["id"] + "/X"]
        p = U.resolve_pad_1D(X.shape, to_pad(c["pad"]), c["kernel_width"], c["stride"], c["dilation"])
        assert list(p) == list(z[c["id"] + "/pad"])
        assert U.calc_conv_out_dims(X.shape, z[c["id"] + "/W"].shape, c["stride"], p, c["dilation"]) == z[c["id"] + "/Z"].shape
    assert U.calc_pad_dims_1D((2, 16, 4), 16, 2, 1, dilation=3, causal=True) == (4, 0)       # all of it on the left
    with pytest.raises(ValueError):
        U.calc_pad_dims_1D([2, 16, 4], 16, 2, 1)                                            # utils.py:153-163
    with pytest.raises(ValueError):
        U.calc_pad_dims_1D((2, 16, 4), 3, 2, 1)                                             # negative pad
    P = pkg.Pool2D(kernel_shape=(3, 3), stride=2, pad=1, mode="average")
    d, pm = P._desc((4, 9, 7, 5))
    assert (d.OH, d.OW, d.pr1, d.pc2, pm) == (5, 4, 1, 1, 1)
    assert P.hyperparameters["layer"] == "Pool2D" and P.hyperparameters["mode"] == "average"
    with pytest.raises(ValueError):
        pkg.Pool2D(kernel_shape=(2, 2), mode="median")._desc((1, 4, 4, 1))
    F = pkg.Flatten(keep_dim="first")
    Y = F.forward(np.zeros((3, 4, 5, 2)))
    assert Y.shape == (3, 40) and F.backward(np.ones((3, 40))).shape == (3, 4, 5, 2)
    assert pkg.Flatten(keep_dim=-1).forward(np.zeros((2, 3))).shape == (1, 6)
    fc = pkg.FullyConnected(n_out=7, act_fn="relu")
    assert fc.hyperparameters == {"layer": "FullyConnected", "init": "glorot_uniform", "n_in": None, "n_out": 7,
                                  "act_fn": "ReLU", "optimizer": fc._opt_hparams()}
    W = pkg.WavenetResidualModule(ch_residual=4, ch_dilation=6, dilation=2, kernel_width=3)
    assert W.conv_dilation.kernel_width == 2 and W.conv_dilation.pad == "causal"            # modules.py:161-169
    assert W.hyperparameters["component_ids"] == ["conv_1x1", "add_skip", "add_residual", "conv_dilation", "multiply_gate"]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            pkg.Conv1D(4, 3).forward(np.zeros((1, 8, 2)))
        with pytest.raises(RuntimeError):
            P.forward(np.zeros((1, 4, 4, 1)))


def test_bench_algorithmic_work_matches_survey_table():
    """bench.py's roofline numerator (fwd + bwd = 3 passes of 2 * MACs, padded taps counted) against SURVEY.md 8d:
    cfg1 0.0434, cfg2 177.57, cfg3 27.41, cfg4 360.78 (convolutions only), cfg5 103.08 GFLOP per step."""
    import importlib
    bench = importlib.import_module("bench")
    want = {"cfg1": 0.0434, "cfg2": 177.57, "cfg3": 27.41, "cfg4": 360.78, "cfg5": 103.08}
    for name, gf in want.items():
        cfg = dict(bench.CONFIGS[name])
        got = 3.0 * bench.flops_per_pass(cfg, cfg["N"]) / 1e9
        assert abs(got - gf) / gf < 2e-3, (name, got, gf)
    # both arms print the same `config` object (the driver compares them), and bench.py's own geometry agrees with the package
    pkg = npml()
    for name in want:
        cfg = dict(bench.CONFIGS[name]); cfg["_id"] = name
        assert bench.config_dict(cfg, "bf16", 2, cfg["N"])["global_batch"] == 2 * cfg["N"]
        if cfg["kind"] == "conv":
            p4 = pkg.utils.resolve_pad_2D((1, cfg["H"], cfg["W"], cfg["C"]), cfg["pad"], cfg["k"], cfg["stride"], cfg["dilation"])
            assert bench.out_hw(cfg) == tuple(pkg.utils.conv_out_hw(cfg["H"], cfg["W"], cfg["k"][0], cfg["k"][1], p4, cfg["stride"], cfg["dilation"]))
        elif cfg["kind"] == "deconv":
            assert bench.out_hw(cfg) == tuple(pkg.utils.deconv_geometry(cfg["H"], cfg["W"], cfg["k"][0], cfg["k"][1], cfg["pad"], cfg["stride"])[1:])


def test_reference_arm_runs_the_unmodified_reference():
    """bench.py --impl reference: the CPU arm is the reference itself when baseline/_ref is staged (kind "reference"),
    the oracle port otherwise (kind "port"); one tiny timed step of cfg1 through it."""
    import importlib
    bench = importlib.import_module("bench")
    cfg = dict(bench.CONFIGS["cfg1"]); cfg["_id"] = "cfg1"
    arm = bench.CpuArm(cfg, 1001)
    from baseline import reference_arm as R
    assert arm.kind == ("reference" if R.available() else "port")
    X, dY = arm.inputs(2)
    out = arm.step(X, dY)
    assert out[0].shape[0] == 2
    n, ts, sample = arm.time(0.5, 1, 0)
    assert 2 <= n <= cfg["N"] and len(ts) == 1 and "host threads" in sample
