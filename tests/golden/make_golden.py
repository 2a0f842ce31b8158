"""
Generate the golden input/output vectors under tests/golden/ by importing the
UNMODIFIED reference (ddbourgin/numpy-ml at /root/reference) in the build container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed here as small
.npz fixtures.  Every array is float64 exactly as the reference returned it; inputs are
float32-representable (drawn as float32, upcast) so the CUDA paths can consume them
without a representation error.  Case hyper-parameters are stored as a JSON string under
the key "__cases__" of each file.
"""
import collections
import collections.abc
import json
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
collections.Hashable = collections.abc.Hashable            # numpy_ml/utils/data_structures.py:3
sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
sys.path.insert(0, "/root/reference")

from numpy_ml.neural_nets import utils as U                      # noqa: E402
from numpy_ml.neural_nets import activations as A                # noqa: E402
from numpy_ml.neural_nets.layers import (Conv2D, Deconv2D, BatchNorm2D, Add, Conv1D, Pool2D, Flatten,   # noqa: E402
                                         Multiply, FullyConnected)
from numpy_ml.neural_nets.modules import (SkipConnectionConvModule, SkipConnectionIdentityModule,    # noqa: E402
                                          WavenetResidualModule)

HERE = os.path.dirname(os.path.abspath(__file__))


def f32(rng, shape, scale=1.0):
    return (scale * rng.standard_normal(shape, dtype=np.float32)).astype(np.float64)


ACTS = {
    "identity": lambda: A.Identity(),
    "affine": lambda: A.Affine(slope=0.75, intercept=0.125),
    "relu": lambda: A.ReLU(),
    "leaky_relu": lambda: A.LeakyReLU(alpha=0.25),
    "tanh": lambda: A.Tanh(),
    "sigmoid": lambda: A.Sigmoid(),
    "elu": lambda: A.ELU(alpha=0.5),
    "selu": lambda: A.SELU(),
    "hard_sigmoid": lambda: A.HardSigmoid(),
    "softplus": lambda: A.SoftPlus(),
    "gelu": lambda: A.GELU(),
    "exponential": lambda: A.Exponential(),
}
ACT_PARAMS = {"affine": (0.75, 0.125), "leaky_relu": (0.25, 0.0), "elu": (0.5, 0.0)}


def save(name, arrays, cases):
    arrays = {k: np.ascontiguousarray(v) for k, v in arrays.items()}
    arrays["__cases__"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(HERE, name), **arrays)
    print(name, len(cases), "cases,", sum(v.nbytes for v in arrays.values()) // 1024, "KiB")


def to_list(p):
    return list(p) if isinstance(p, tuple) else p


def to_pad(p):
    return tuple(p) if isinstance(p, list) else p


# --------------------------------------------------------------------------------------------
def gen_utils():
    rng = np.random.default_rng(7)
    arrays, cases = {}, []

    # calc_pad_dims_2D
    pads = []
    for (xs, od, ks, s, d) in [((2, 56, 56, 3), (56, 56), (3, 3), 1, 0), ((1, 28, 28, 2), (28, 28), (3, 3), 2, 2),
                               ((1, 9, 7, 2), (9, 7), (2, 4), 1, 0), ((1, 8, 8, 1), (8, 8), (4, 4), 1, 0),
                               ((1, 32, 32, 4), (32, 32), (1, 1), 1, 0), ((3, 11, 13, 2), (6, 7), (3, 5), 2, 1),
                               ((1, 5, 5, 1), (5, 5), (3, 3), 1, 3)]:
        pads.append({"X_shape": list(xs), "out_dim": list(od), "kernel_shape": list(ks), "stride": s,
                     "dilation": d, "out": list(int(v) for v in U.calc_pad_dims_2D(xs, od, ks, s, d))})
    cases.append({"kind": "calc_pad_dims_2D", "items": pads})

    # pad2D / dilate
    X = f32(rng, (2, 5, 4, 3))
    for i, (pad, ks, s, d) in enumerate([(2, None, None, 0), ((1, 2), None, None, 0), ((0, 3, 2, 1), None, None, 0),
                                         ("same", (3, 3), 1, 0), ("same", (2, 3), 2, 1)]):
        Xp, p4 = U.pad2D(X, pad, ks, s, d)
        arrays[f"pad2D{i}/X"], arrays[f"pad2D{i}/out"] = X, Xp
        cases.append({"kind": "pad2D", "id": f"pad2D{i}", "pad": to_list(pad), "kernel_shape": to_list(ks),
                      "stride": s, "dilation": d, "p4": [int(v) for v in p4]})
    for i, d in enumerate([1, 2]):
        arrays[f"dilate{i}/X"], arrays[f"dilate{i}/out"] = X, U.dilate(X, d)
        cases.append({"kind": "dilate", "id": f"dilate{i}", "d": d})

    # im2col / col2im / conv2D / conv2D_naive
    confs = [((2, 6, 5, 3), (3, 3, 3, 4), 1, 1, 0), ((2, 7, 7, 2), (2, 3, 2, 3), 2, (1, 2), 0),
             ((1, 9, 8, 3), (3, 2, 3, 2), 2, (2, 0, 1, 3), 1), ((3, 8, 8, 2), (3, 3, 2, 5), 1, "same", 0),
             ((2, 10, 10, 2), (3, 3, 2, 2), 2, "same", 2), ((2, 4, 4, 1), (4, 4, 1, 3), 1, 0, 0),
             ((2, 5, 6, 4), (1, 1, 4, 3), 1, 0, 0), ((1, 12, 11, 2), (2, 2, 2, 2), 3, 1, 4)]
    for i, (xs, ws, s, pad, d) in enumerate(confs):
        X, W = f32(rng, xs), f32(rng, ws)
        X_col, p4 = U.im2col(X, ws, pad, s, d)
        Z = U.conv2D(X, W, s, pad, d)
        Zn = U.conv2D_naive(X, W, s, pad, d)
        assert np.allclose(Z, Zn)
        C = f32(rng, X_col.shape)
        back = U.col2im(C, xs, ws, tuple(int(v) for v in p4), s, d)
        arrays.update({f"conv{i}/X": X, f"conv{i}/W": W, f"conv{i}/X_col": X_col, f"conv{i}/Z": Z,
                       f"conv{i}/C": C, f"conv{i}/col2im": back})
        cases.append({"kind": "conv", "id": f"conv{i}", "stride": s, "pad": to_list(pad), "dilation": d,
                      "p4": [int(v) for v in p4]})

    # deconv2D_naive (pinned domain: symmetric pads)
    dconfs = [((2, 4, 4, 3), (4, 4, 3, 2), 2, 1), ((2, 5, 3, 2), (3, 3, 2, 4), 1, 1), ((1, 3, 4, 2), (2, 3, 2, 3), 3, (1, 0)),
              ((2, 4, 5, 3), (3, 3, 3, 2), 2, 0), ((1, 6, 6, 2), (3, 3, 2, 2), 1, "same"), ((2, 3, 3, 2), (5, 5, 2, 3), 2, 2)]
    for i, (xs, ws, s, pad) in enumerate(dconfs):
        X, W = f32(rng, xs), f32(rng, ws)
        arrays.update({f"deconv{i}/X": X, f"deconv{i}/W": W, f"deconv{i}/Z": U.deconv2D_naive(X, W, s, pad, 0)})
        cases.append({"kind": "deconv", "id": f"deconv{i}", "stride": s, "pad": to_list(pad)})
    save("utils.npz", arrays, cases)


# --------------------------------------------------------------------------------------------
def gen_activations():
    x = np.concatenate([np.linspace(-4, 4, 41), [0.0, -0.0, 2.5, -2.5, 1e-3, -1e-3]]).astype(np.float32).astype(np.float64)
    arrays, cases = {"x": x}, []
    for name, mk in ACTS.items():
        a = mk()
        arrays[f"{name}/fn"] = np.asarray(a.fn(x), dtype=np.float64)
        arrays[f"{name}/grad"] = np.asarray(a.grad(x), dtype=np.float64)
        p0, p1 = ACT_PARAMS.get(name, (0.0, 0.0))
        cases.append({"act": name, "p0": p0, "p1": p1, "str": str(a)})
    save("activations.npz", arrays, cases)


# --------------------------------------------------------------------------------------------
def run_conv_layer(rng, cls, xs, out_ch, ks, pad, s, d, act, two_calls=False):
    kw = dict(out_ch=out_ch, kernel_shape=ks, pad=pad, stride=s, act_fn=ACTS[act]())
    if cls is Conv2D:
        kw["dilation"] = d
    L = cls(**kw)
    X = f32(rng, xs)
    L.forward(X, retain_derived=False)           # lazy init only
    W = f32(rng, L.parameters["W"].shape, 0.4)
    b = f32(rng, L.parameters["b"].shape, 0.3)
    L.parameters["W"], L.parameters["b"] = W, b
    L.flush_gradients()
    Y = L.forward(X)
    out = {"X": X, "W": W, "b": b, "Y": Y, "Z": L.derived_variables["Z"][0]}
    dY = f32(rng, Y.shape)
    out["dLdy"] = dY
    if two_calls:
        X2 = f32(rng, xs)
        Y2 = L.forward(X2)
        dY2 = f32(rng, Y2.shape)
        dXs = L.backward([dY, dY2])
        out.update({"X2": X2, "Y2": Y2, "dLdy2": dY2, "dX": dXs[0], "dX2": dXs[1]})
    else:
        out["dX"] = L.backward(dY)
    out["dW"], out["db"] = L.gradients["W"], L.gradients["b"]
    return out


def gen_conv2d():
    rng = np.random.default_rng(11)
    arrays, cases = {}, []
    confs = [
        # xs, out_ch, ks, pad, s, d, act, two_calls
        ((2, 7, 6, 3), 4, (3, 3), 1, 1, 0, "identity", False),
        ((2, 8, 8, 2), 3, (3, 3), "same", 1, 0, "relu", False),
        ((3, 9, 7, 1), 5, (2, 3), (1, 2), 2, 0, "tanh", False),
        ((2, 10, 9, 3), 2, (3, 2), (2, 0, 1, 3), 2, 1, "sigmoid", False),
        ((2, 12, 12, 2), 4, (3, 3), 0, 2, 2, "identity", False),       # cfg3 miniature
        ((2, 12, 12, 2), 3, (3, 3), "same", 2, 2, "affine", False),    # 'same' with stride 2, dilation 2
        ((2, 5, 5, 4), 6, (1, 1), 0, 1, 0, "leaky_relu", False),
        ((1, 4, 4, 2), 2, (4, 4), 0, 1, 0, "elu", False),              # kernel == input
        ((2, 6, 6, 5), 3, (3, 3), 1, 1, 0, "identity", True),          # two forwards, list backward
        ((2, 11, 10, 3), 17, (3, 3), 1, 3, 0, "softplus", False),      # stride leftovers, odd channels
        ((2, 6, 7, 2), 2, (2, 2), 1, 1, 4, "selu", False),
        ((2, 7, 7, 3), 4, (3, 3), 1, 1, 0, "gelu", False),
        ((2, 7, 7, 3), 4, (3, 3), 1, 1, 0, "hard_sigmoid", False),
        ((2, 6, 6, 2), 3, (3, 3), 1, 1, 0, "exponential", False),
        ((4, 28, 28, 1), 32, (3, 3), "same", 1, 0, "identity", False),  # cfg1 at N=4
    ]
    for i, (xs, oc, ks, pad, s, d, act, two) in enumerate(confs):
        out = run_conv_layer(rng, Conv2D, xs, oc, ks, pad, s, d, act, two)
        arrays.update({f"c{i}/{k}": v for k, v in out.items()})
        p0, p1 = ACT_PARAMS.get(act, (0.0, 0.0))
        cases.append({"id": f"c{i}", "out_ch": oc, "kernel_shape": list(ks), "pad": to_list(pad), "stride": s,
                      "dilation": d, "act": act, "p0": p0, "p1": p1, "two_calls": two})
    save("conv2d_layer.npz", arrays, cases)


def gen_deconv2d():
    rng = np.random.default_rng(13)
    arrays, cases = {}, []
    confs = [
        ((2, 4, 4, 3), 2, (4, 4), 1, 2, "identity", False),            # cfg5 miniature
        ((2, 5, 4, 2), 3, (3, 3), 1, 1, "relu", False),
        ((2, 3, 4, 2), 3, (2, 3), (1, 0), 3, "tanh", False),
        ((1, 4, 5, 3), 2, (3, 3), 0, 2, "sigmoid", False),
        ((2, 6, 6, 2), 4, (3, 3), "same", 1, "affine", False),
        ((2, 3, 3, 2), 5, (5, 5), 2, 2, "leaky_relu", True),
        ((2, 4, 4, 4), 3, (1, 1), 0, 1, "identity", False),
        ((2, 4, 3, 2), 3, (2, 2), 0, 2, "identity", False),
    ]
    for i, (xs, oc, ks, pad, s, act, two) in enumerate(confs):
        out = run_conv_layer(rng, Deconv2D, xs, oc, ks, pad, s, 0, act, two)
        arrays.update({f"d{i}/{k}": v for k, v in out.items()})
        p0, p1 = ACT_PARAMS.get(act, (0.0, 0.0))
        cases.append({"id": f"d{i}", "out_ch": oc, "kernel_shape": list(ks), "pad": to_list(pad), "stride": s,
                      "act": act, "p0": p0, "p1": p1, "two_calls": two})
    save("deconv2d_layer.npz", arrays, cases)


def gen_bn_add():
    rng = np.random.default_rng(17)
    arrays, cases = {}, []
    for i, (xs, mm, eps) in enumerate([((3, 5, 4, 6), 0.9, 1e-5), ((2, 7, 7, 3), 0.5, 1e-3), ((4, 2, 3, 1), 0.99, 1e-5)]):
        L = BatchNorm2D(momentum=mm, epsilon=eps)
        X = f32(rng, xs) * 2.0 + 0.5
        L.forward(X, retain_derived=False)
        L.parameters["scaler"] = f32(rng, (xs[3],)) * 0.5 + 1.0
        L.parameters["intercept"] = f32(rng, (xs[3],)) * 0.1
        g, h = L.parameters["scaler"].copy(), L.parameters["intercept"].copy()
        y = L.forward(X)
        dy = f32(rng, xs)
        dX = L.backward(dy)
        rm1, rv1 = L.parameters["running_mean"].copy(), L.parameters["running_var"].copy()
        X2 = f32(rng, xs)
        y_infer = L.forward(X2, retain_derived=False)       # running-stats path
        arrays.update({f"bn{i}/X": X, f"bn{i}/scaler": g, f"bn{i}/intercept": h, f"bn{i}/y": y, f"bn{i}/dLdy": dy,
                       f"bn{i}/dX": dX, f"bn{i}/dScaler": L.gradients["scaler"], f"bn{i}/dIntercept": L.gradients["intercept"],
                       f"bn{i}/running_mean": rm1, f"bn{i}/running_var": rv1, f"bn{i}/X2": X2, f"bn{i}/y_infer": y_infer})
        cases.append({"kind": "bn", "id": f"bn{i}", "momentum": mm, "epsilon": eps})
    for i, (n_in, act) in enumerate([(2, "relu"), (3, "tanh"), (2, "identity")]):
        L = Add(act_fn=ACTS[act]())
        Xs = [f32(rng, (2, 4, 3, 5)) for _ in range(n_in)]
        Y = L.forward(Xs)
        dY = f32(rng, Y.shape)
        dXs = L.backward(dY)
        for j, x in enumerate(Xs):
            arrays[f"add{i}/X{j}"] = x
        arrays.update({f"add{i}/Y": Y, f"add{i}/dLdY": dY, f"add{i}/dX": np.asarray(dXs[0], dtype=np.float64)})
        cases.append({"kind": "add", "id": f"add{i}", "n_in": n_in, "act": act})
    save("bn_add.npz", arrays, cases)


def gen_module():
    rng = np.random.default_rng(19)
    arrays, cases = {}, []
    confs = [
        ((3, 8, 8, 3), 4, 5, (3, 3), (3, 3), (1, 1), 1, 1, 1, 1, 1, "relu"),          # cfg4 miniature
        ((2, 9, 7, 2), 3, 4, (3, 2), (2, 3), (2, 2), (1, 2), (2, 1), 1, 1, 1, "identity"),
        ((2, 10, 10, 2), 3, 3, (3, 3), (3, 3), (3, 3), 1, 1, 2, 1, 2, "tanh"),
    ]
    for i, (xs, oc1, oc2, k1, k2, ks, pad1, pad2, s1, s2, ss, act) in enumerate(confs):
        M = SkipConnectionConvModule(out_ch1=oc1, out_ch2=oc2, kernel_shape1=k1, kernel_shape2=k2, kernel_shape_skip=ks,
                                     pad1=pad1, pad2=pad2, stride1=s1, stride2=s2, stride_skip=ss,
                                     act_fn=None if act == "identity" else ACTS[act]())
        X = f32(rng, xs)
        M.forward(X, retain_derived=False)          # lazy init of every component
        P = {}
        for tag, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
            conv.parameters["W"] = f32(rng, conv.parameters["W"].shape, 0.4)
            conv.parameters["b"] = f32(rng, conv.parameters["b"].shape, 0.3)
            bn.parameters["scaler"] = f32(rng, bn.parameters["scaler"].shape) * 0.5 + 1.0
            bn.parameters["intercept"] = f32(rng, bn.parameters["intercept"].shape) * 0.1
            bn.reset_running_stats()
            P["W" + tag], P["b" + tag] = conv.parameters["W"].copy(), conv.parameters["b"].copy()
            P["g" + tag], P["h" + tag] = bn.parameters["scaler"].copy(), bn.parameters["intercept"].copy()
        for c in M.components:                      # per-component flush (module.update() breaks the next forward)
            c.flush_gradients()
        Y = M.forward(X)
        dY = f32(rng, Y.shape)
        dX = M.backward(dY).copy()
        out = {"X": X, "dLdY": dY, "Y": Y, "dX": dX, "pad_skip": np.asarray(M.pad_skip)}
        out.update(P)
        dv = M.derived_variables
        for k in ["conv1_out", "conv2_out", "batchnorm1_out", "batchnorm2_out", "conv_skip_out", "batchnorm_skip_out",
                  "dLdBn1", "dLdBn2", "dLdConv1", "dLdConv2", "dLdBnSkip", "dLdConvSkip"]:
            out[k] = np.asarray(dv[k], dtype=np.float64)
        for tag, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2), ("s", M.conv_skip, M.batchnorm_skip)):
            out["dW" + tag], out["db" + tag] = conv.gradients["W"], conv.gradients["b"]
            out["dg" + tag], out["dh" + tag] = bn.gradients["scaler"], bn.gradients["intercept"]
            out["rm" + tag], out["rv" + tag] = bn.parameters["running_mean"], bn.parameters["running_var"]
        arrays.update({f"m{i}/{k}": v for k, v in out.items()})
        p0, p1 = ACT_PARAMS.get(act, (0.0, 0.0))
        cases.append({"id": f"m{i}", "out_ch1": oc1, "out_ch2": oc2, "kernel_shape1": list(k1), "kernel_shape2": list(k2),
                      "kernel_shape_skip": list(ks), "pad1": to_list(pad1), "pad2": to_list(pad2), "stride1": s1,
                      "stride2": s2, "stride_skip": ss, "act": act, "p0": p0, "p1": p1})
    save("skip_conv_module.npz", arrays, cases)


def gen_optimizers():
    """four steps of every reference optimizer (with and without clip-norm, and with the
    Exponential / Noam schedules) on one (3, 3, 4, 5) parameter; every intermediate parameter and cache is stored."""
    from numpy_ml.neural_nets.optimizers import SGD, AdaGrad, RMSProp, Adam
    from numpy_ml.neural_nets.schedulers import ExponentialScheduler, NoamScheduler
    rng = np.random.default_rng(9)
    arrays, cases = {}, []
    specs = [
        ("sgd", SGD, dict(lr=0.05, momentum=0.9), None),
        ("sgd_clip", SGD, dict(lr=0.05, momentum=0.5, clip_norm=1.5), None),
        ("adagrad", AdaGrad, dict(lr=0.1, eps=1e-7), None),
        ("adagrad_clip", AdaGrad, dict(lr=0.1, eps=1e-7, clip_norm=2.0), None),
        ("rmsprop", RMSProp, dict(lr=0.01, decay=0.9, eps=1e-7), None),
        ("rmsprop_exp", RMSProp, dict(lr=0.01, decay=0.8, eps=1e-7),
         dict(kind="exponential", initial_lr=0.02, stage_length=2, staircase=True, decay=0.5)),
        ("adam", Adam, dict(lr=0.01, decay1=0.9, decay2=0.999, eps=1e-7), None),
        ("adam_clip_noam", Adam, dict(lr=0.01, decay1=0.8, decay2=0.99, eps=1e-7, clip_norm=3.0),
         dict(kind="noam", model_dim=64, scale_factor=2.0, warmup_steps=3)),
    ]
    for cid, cls, kw, sched in specs:
        kwargs = dict(kw)
        if sched is not None:
            sk = {k: v for k, v in sched.items() if k != "kind"}
            kwargs["lr_scheduler"] = ExponentialScheduler(**sk) if sched["kind"] == "exponential" else NoamScheduler(**sk)
        opt = cls(**kwargs)
        param = f32(rng, (3, 3, 4, 5))
        arrays[cid + "/param0"] = param
        lrs = []
        for step in range(4):
            grad = f32(rng, (3, 3, 4, 5), 2.0)
            arrays["{}/grad{}".format(cid, step)] = grad
            opt.step()                                   # LayerBase.update steps the optimizer first (layers.py:82)
            lrs.append(float(opt.lr_scheduler(opt.cur_step, None)))
            param = opt.update(param, grad, "W")
            arrays["{}/param{}".format(cid, step + 1)] = np.array(param)
            c = opt.cache["W"]                           # copies: AdaGrad updates its cache in place
            if isinstance(c, dict):
                arrays["{}/mean{}".format(cid, step + 1)] = np.array(c["mean"])
                arrays["{}/var{}".format(cid, step + 1)] = np.array(c["var"])
            else:
                arrays["{}/cache{}".format(cid, step + 1)] = np.array(c)
        cases.append(dict(id=cid, kind=cls.__name__, kwargs=kw, sched=sched, lrs=lrs, string=str(opt)))
    save("optimizers.npz", arrays, cases)


def gen_conv1d():
    """Conv1D through the reference's own expand-dims route, incl. 'same' / 'causal' pads."""
    rng = np.random.default_rng(23)
    arrays, cases = {}, []
    confs = [
        # xs (n, l, c), out_ch, kw, pad, s, d, act, two_calls
        ((2, 9, 3), 4, 3, 1, 1, 0, "identity", False),
        ((3, 12, 2), 5, 3, "same", 1, 0, "relu", False),
        ((2, 16, 4), 3, 2, "causal", 1, 3, "tanh", False),            # Wavenet-style dilated causal conv
        ((2, 11, 3), 2, 4, (2, 1), 2, 0, "sigmoid", False),
        ((2, 14, 2), 6, 3, "same", 1, 2, "affine", False),
        ((2, 10, 5), 7, 1, 0, 1, 0, "leaky_relu", False),
        ((2, 8, 3), 3, 3, 1, 1, 0, "identity", True),
        ((1, 13, 1), 8, 5, "causal", 1, 1, "elu", False),
    ]
    for i, (xs, oc, kw, pad, s, d, act, two) in enumerate(confs):
        L = Conv1D(out_ch=oc, kernel_width=kw, pad=pad, stride=s, dilation=d, act_fn=ACTS[act]())
        X = f32(rng, xs)
        L.forward(X, retain_derived=False)
        W, b = f32(rng, L.parameters["W"].shape, 0.4), f32(rng, L.parameters["b"].shape, 0.3)
        L.parameters["W"], L.parameters["b"] = W, b
        L.flush_gradients()
        Y = L.forward(X)
        out = {"X": X, "W": W, "b": b, "Y": Y, "Z": L.derived_variables["Z"][0]}
        out["pad"] = np.asarray(U.pad1D(X, pad, kw, s, dilation=d)[1])
        dY = f32(rng, Y.shape)
        out["dLdy"] = dY
        if two:
            X2 = f32(rng, xs)
            Y2 = L.forward(X2)
            dY2 = f32(rng, Y2.shape)
            dXs = L.backward([dY, dY2])
            out.update({"X2": X2, "Y2": Y2, "dLdy2": dY2, "dX": dXs[0], "dX2": dXs[1]})
        else:
            out["dX"] = L.backward(dY)
        out["dW"], out["db"] = L.gradients["W"], L.gradients["b"]
        out["conv1D"] = U.conv1D(X, W, s, pad, d)                          # the functional form (utils.py:668-717)
        arrays.update({f"c{i}/{k}": v for k, v in out.items()})
        p0, p1 = ACT_PARAMS.get(act, (0.0, 0.0))
        cases.append({"id": f"c{i}", "out_ch": oc, "kernel_width": kw, "pad": to_list(pad), "stride": s, "dilation": d,
                      "act": act, "p0": p0, "p1": p1, "two_calls": two})
    save("conv1d_layer.npz", arrays, cases)


def gen_pool2d():
    """Pool2D max / average forward + backward, Flatten.  Max-mode backward is generated only
    for pad == 0: with padding the reference reads its windows from the unpadded input (layers.py:3326)."""
    rng = np.random.default_rng(29)
    arrays, cases = {}, []
    confs = [
        # xs, kernel, stride, pad, mode
        ((2, 8, 8, 3), (2, 2), 2, 0, "max"),
        ((2, 9, 7, 5), (3, 3), 2, 0, "max"),
        ((3, 6, 6, 2), (2, 3), 1, 0, "max"),
        ((2, 8, 8, 3), (2, 2), 2, 0, "average"),
        ((2, 9, 7, 4), (3, 3), 1, 1, "average"),
        ((2, 7, 10, 2), (3, 2), 2, (1, 2), "average"),
        ((2, 6, 6, 8), (3, 3), 1, "same", "average"),
        ((2, 7, 7, 3), (3, 3), 2, 1, "max"),          # forward only (backward outside the pinned domain)
        ((1, 5, 5, 1), (5, 5), 1, 0, "max"),          # kernel == input
    ]
    for i, (xs, ks, s, pad, mode) in enumerate(confs):
        L = Pool2D(kernel_shape=ks, stride=s, pad=pad, mode=mode)
        X = f32(rng, xs)
        if mode == "max":                             # ties: the first maximal pixel (row-major) takes the gradient
            X[0, :2, :2, 0] = X.max() + 1.0
            X = np.round(X * 4) / 4
        Y = L.forward(X)
        out = {"X": X, "Y": Y}
        bwd = not (mode == "max" and pad != 0)
        if bwd:
            dY = f32(rng, Y.shape)
            out["dLdY"], out["dX"] = dY, L.backward(dY)
        arrays.update({f"p{i}/{k}": v for k, v in out.items()})
        cases.append({"id": f"p{i}", "kernel_shape": list(ks), "stride": s, "pad": to_list(pad), "mode": mode,
                      "backward": bwd})
    for j, keep in enumerate(["first", "last", -1]):
        F = Flatten(keep_dim=keep)
        X = f32(rng, (3, 4, 5, 2))
        Y = F.forward(X)
        dY = f32(rng, Y.shape)
        arrays.update({f"f{j}/X": X, f"f{j}/Y": Y, f"f{j}/dLdy": dY, f"f{j}/dX": F.backward(dY)})
        cases.append({"id": f"f{j}", "keep_dim": keep, "flatten": True})
    save("pool2d.npz", arrays, cases)


def gen_identity_module():
    """SkipConnectionIdentityModule (modules.py:360-612)."""
    rng = np.random.default_rng(37)
    arrays, cases = {}, []
    confs = [((3, 8, 8, 4), 5, (3, 3), (3, 3), "relu"),
             ((2, 7, 9, 3), 4, (3, 3), (1, 1), "identity"),
             ((2, 6, 6, 2), 3, (3, 3), (3, 3), "tanh")]
    for i, (xs, oc, k1, k2, act) in enumerate(confs):
        M = SkipConnectionIdentityModule(out_ch=oc, kernel_shape1=k1, kernel_shape2=k2,
                                         act_fn=None if act == "identity" else ACTS[act]())
        X = f32(rng, xs)
        M.forward(X, retain_derived=False)
        P = {}
        for tag, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2)):
            conv.parameters["W"] = f32(rng, conv.parameters["W"].shape, 0.4)
            conv.parameters["b"] = f32(rng, conv.parameters["b"].shape, 0.3)
            bn.parameters["scaler"] = f32(rng, bn.parameters["scaler"].shape) * 0.5 + 1.0
            bn.parameters["intercept"] = f32(rng, bn.parameters["intercept"].shape) * 0.1
            bn.reset_running_stats()
            P["W" + tag], P["b" + tag] = conv.parameters["W"].copy(), conv.parameters["b"].copy()
            P["g" + tag], P["h" + tag] = bn.parameters["scaler"].copy(), bn.parameters["intercept"].copy()
        for c in M.components:
            c.flush_gradients()
        Y = M.forward(X)
        dY = f32(rng, Y.shape)
        dX = np.array(M.backward(dY))
        out = {"X": X, "dLdY": dY, "Y": Y, "dX": dX}
        out.update(P)
        dv = M.derived_variables
        for k in ["conv1_out", "conv2_out", "batchnorm1_out", "batchnorm2_out", "dLdBn1", "dLdBn2", "dLdConv1", "dLdConv2"]:
            out[k] = np.asarray(dv[k], dtype=np.float64)
        for tag, conv, bn in (("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2)):
            out["dW" + tag], out["db" + tag] = conv.gradients["W"], conv.gradients["b"]
            out["dg" + tag], out["dh" + tag] = bn.gradients["scaler"], bn.gradients["intercept"]
            out["rm" + tag], out["rv" + tag] = bn.parameters["running_mean"], bn.parameters["running_var"]
        arrays.update({f"m{i}/{k}": v for k, v in out.items()})
        p0, p1 = ACT_PARAMS.get(act, (0.0, 0.0))
        cases.append({"id": f"m{i}", "out_ch": oc, "kernel_shape1": list(k1), "kernel_shape2": list(k2), "act": act,
                      "p0": p0, "p1": p1})
    save("skip_identity_module.npz", arrays, cases)


def gen_wavenet_vae():
    """SURVEY.md 8f ranks 2-3, the models either side of the path: Multiply, FullyConnected, WavenetResidualModule
    (modules.py:119-357) and the BernoulliVAE encoder chain Conv -> MaxPool -> Conv -> MaxPool -> Flatten -> FC -> FC
    (models/vae.py:142-189) built from the reference's own layers."""
    rng = np.random.default_rng(43)
    arrays, cases = {}, []
    # Multiply
    for i, (shape, n_in, act) in enumerate([((3, 5, 4), 2, "identity"), ((2, 4, 4, 3), 3, "tanh")]):
        M = Multiply(act_fn=ACTS[act]())
        Xs = [f32(rng, shape) for _ in range(n_in)]
        Y = M.forward(Xs)
        dY = f32(rng, Y.shape)
        g = M.backward(dY)
        for j in range(n_in):
            arrays[f"mul{i}/X{j}"], arrays[f"mul{i}/dX{j}"] = Xs[j], g[j]
        arrays[f"mul{i}/Y"], arrays[f"mul{i}/dLdY"] = Y, dY
        cases.append({"id": f"mul{i}", "kind": "multiply", "n_in": n_in, "act": act})
    # FullyConnected
    for i, (n, din, dout, act) in enumerate([(5, 12, 7, "relu"), (64, 32, 16, "identity"), (3, 9, 4, "sigmoid")]):
        F = FullyConnected(n_out=dout, act_fn=ACTS[act]())
        X = f32(rng, (n, din))
        F.forward(X, retain_derived=False)
        W, b = f32(rng, (din, dout), 0.4), f32(rng, (1, dout), 0.3)
        F.parameters["W"], F.parameters["b"] = W, b
        F.flush_gradients()
        Y = F.forward(X)
        dY = f32(rng, Y.shape)
        dX = F.backward(dY)
        arrays.update({f"fc{i}/X": X, f"fc{i}/W": W, f"fc{i}/b": b, f"fc{i}/Y": Y, f"fc{i}/dLdy": dY, f"fc{i}/dX": dX,
                       f"fc{i}/dW": F.gradients["W"], f"fc{i}/db": F.gradients["b"]})
        cases.append({"id": f"fc{i}", "kind": "fc", "n_out": dout, "act": act})
    # WavenetResidualModule
    for i, (n, l, chr_, chd, dil, with_main) in enumerate([(2, 16, 4, 6, 1, True), (3, 24, 8, 8, 3, True), (2, 12, 3, 5, 0, False)]):
        M = WavenetResidualModule(ch_residual=chr_, ch_dilation=chd, dilation=dil, kernel_width=2)
        Xm, Xs = f32(rng, (n, l, chr_)), f32(rng, (n, l, chr_))
        M.forward(Xm, Xs)
        P = {}
        for tag, conv in (("d", M.conv_dilation), ("1", M.conv_1x1)):
            conv.parameters["W"] = f32(rng, conv.parameters["W"].shape, 0.5)
            conv.parameters["b"] = f32(rng, conv.parameters["b"].shape, 0.3)
            P["W" + tag], P["b" + tag] = conv.parameters["W"].copy(), conv.parameters["b"].copy()
        for c in M.components:
            c.flush_gradients()
        Ym, Ys = M.forward(Xm, Xs)
        dYs, dYm = f32(rng, Ys.shape), (f32(rng, Ym.shape) if with_main else None)
        dXm, dXs = M.backward(dYs.copy(), None if dYm is None else dYm.copy())
        out = {"X_main": Xm, "X_skip": Xs, "Y_main": Ym, "Y_skip": Ys, "dY_skip": dYs, "dX_main": np.array(dXm), "dX_skip": np.array(dXs)}
        if dYm is not None:
            out["dY_main"] = dYm
        out.update(P)
        dv = M.derived_variables
        for k in ["conv_dilation_out", "tanh_out", "sigm_out", "multiply_gate_out", "conv_1x1_out", "dLdTanh", "dLdSigmoid",
                  "dLdConv_1x1", "dLdMultiply", "dLdConv_dilation"]:
            out[k] = np.asarray(dv[k], dtype=np.float64)
        out["dWd"], out["dbd"] = M.conv_dilation.gradients["W"], M.conv_dilation.gradients["b"]
        out["dW1"], out["db1"] = M.conv_1x1.gradients["W"], M.conv_1x1.gradients["b"]
        arrays.update({f"wn{i}/{k}": v for k, v in out.items()})
        cases.append({"id": f"wn{i}", "kind": "wavenet", "ch_residual": chr_, "ch_dilation": chd, "dilation": dil,
                      "with_main": with_main})
    # VAE encoder chain (defaults of models/vae.py at a small size): 5x5 'same' conv, 2x2 max pools, two dense layers
    enc = [Conv2D(act_fn=A.ReLU(), pad="same", out_ch=8, stride=1, kernel_shape=(5, 5)),
           Pool2D(mode="max", pad=0, stride=2, kernel_shape=(2, 2)),
           Conv2D(act_fn=A.ReLU(), pad="same", out_ch=16, stride=1, kernel_shape=(5, 5)),
           Pool2D(mode="max", pad=0, stride=2, kernel_shape=(2, 2)),
           Flatten(), FullyConnected(n_out=24, act_fn=A.ReLU()), FullyConnected(n_out=10, act_fn=A.Identity())]
    X = np.round(f32(rng, (4, 12, 12, 1)) * 8) / 8
    out = X
    for L in enc:
        out = L.forward(out)
    k = 0
    for L in enc:                       # non-trivial parameters (biases are zero at init)
        if L.parameters:
            L.parameters["W"] = np.round(f32(rng, L.parameters["W"].shape, 0.3) * 64) / 64
            L.parameters["b"] = f32(rng, L.parameters["b"].shape, 0.2)
            arrays[f"vae/W{k}"], arrays[f"vae/b{k}"] = L.parameters["W"].copy(), L.parameters["b"].copy()
            k += 1
        L.flush_gradients()
    out = X
    for L in enc:
        out = L.forward(out)
    dY = f32(rng, out.shape)
    g = dY
    for L in reversed(enc):
        g = L.backward(g)
    arrays.update({"vae/X": X, "vae/Y": out, "vae/dLdY": dY, "vae/dX": np.array(g)})
    k = 0
    for L in enc:
        if L.parameters:
            arrays[f"vae/dW{k}"], arrays[f"vae/db{k}"] = L.gradients["W"], L.gradients["b"]
            k += 1
    cases.append({"id": "vae", "kind": "vae_encoder"})
    save("wavenet_vae.npz", arrays, cases)


def gen_module_inference():
    """retain_derived=False forwards of the two skip-connection modules AFTER a training forward, so that the running
    statistics every BatchNorm2D normalises with (layers/layers.py:1141-1146) are non-trivial.  Pins the oracle chain the
    fused conv -> BatchNorm2D epilogue (npml_conv2d_fprop_bn) is tested against."""
    rng = np.random.default_rng(23)
    arrays, cases = {}, []
    confs = [("conv", (3, 8, 8, 3), 4, 5, "relu"), ("conv", (2, 9, 9, 2), 3, 3, "tanh"), ("identity", (3, 7, 7, 4), 4, 4, "relu")]
    for i, (kind, xs, oc1, oc2, act) in enumerate(confs):
        if kind == "conv":
            M = SkipConnectionConvModule(out_ch1=oc1, out_ch2=oc2, kernel_shape1=(3, 3), kernel_shape2=(3, 3),
                                         kernel_shape_skip=(1, 1), pad1=1, pad2=1, act_fn=ACTS[act]())
        else:
            M = SkipConnectionIdentityModule(out_ch=oc1, kernel_shape1=(3, 3), kernel_shape2=(3, 3), act_fn=ACTS[act]())
        X = f32(rng, xs)
        M.forward(X, retain_derived=False)
        triples = [("1", M.conv1, M.batchnorm1), ("2", M.conv2, M.batchnorm2)]
        if kind == "conv":
            triples.append(("s", M.conv_skip, M.batchnorm_skip))
        out = {}
        for tag, conv, bn in triples:
            conv.parameters["W"] = f32(rng, conv.parameters["W"].shape, 0.4)
            conv.parameters["b"] = f32(rng, conv.parameters["b"].shape, 0.3)
            bn.parameters["scaler"] = f32(rng, bn.parameters["scaler"].shape) * 0.5 + 1.0
            bn.parameters["intercept"] = f32(rng, bn.parameters["intercept"].shape) * 0.1
            bn.reset_running_stats()
        for c in M.components:
            c.flush_gradients()
        M.forward(f32(rng, xs))                     # a training forward moves the running statistics
        for c in M.components:
            c.flush_gradients()
        for tag, conv, bn in triples:
            out["W" + tag], out["b" + tag] = conv.parameters["W"].copy(), conv.parameters["b"].copy()
            out["g" + tag], out["h" + tag] = bn.parameters["scaler"].copy(), bn.parameters["intercept"].copy()
            out["rm" + tag] = np.asarray(bn.parameters["running_mean"], dtype=np.float64).copy()
            out["rv" + tag] = np.asarray(bn.parameters["running_var"], dtype=np.float64).copy()
        Y = M.forward(X, retain_derived=False)
        for tag, conv, bn in triples:               # inference must not move them
            assert np.array_equal(out["rm" + tag], bn.parameters["running_mean"])
        out.update({"X": X, "Y": np.asarray(Y, dtype=np.float64)})
        arrays.update({f"i{i}/{k}": v for k, v in out.items()})
        cases.append({"id": f"i{i}", "kind": kind, "out_ch1": oc1, "out_ch2": oc2, "act": act})
    save("skip_modules_inference.npz", arrays, cases)


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:                                         # e.g. `make_golden.py conv1d pool2d`
        for name in only:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_wavenet_vae()
    gen_conv1d()
    gen_pool2d()
    gen_identity_module()
    gen_optimizers()
    gen_utils()
    gen_activations()
    gen_conv2d()
    gen_deconv2d()
    gen_bn_add()
    gen_module()
    gen_module_inference()
