"""Pins oracle/conv_oracle.py against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, to_pad, rel_err, npml
from oracle import conv_oracle as O

TOL = 1e-12


def close(a, b, tol=TOL):
    assert rel_err(a, b) <= tol, rel_err(a, b)


def test_utils_golden():
    z, cases = load_golden("utils.npz")
    for c in cases:
        k = c["kind"]
        if k == "calc_pad_dims_2D":
            for it in c["items"]:
                got = O.calc_pad_dims_2D(tuple(it["X_shape"]), tuple(it["out_dim"]), tuple(it["kernel_shape"]),
                                         it["stride"], it["dilation"])
                assert list(got) == it["out"]
        elif k == "pad2D":
            ks = tuple(c["kernel_shape"]) if c["kernel_shape"] else None
            Xp, p4 = O.pad2D(z[c["id"] + "/X"], to_pad(c["pad"]), ks, c["stride"], c["dilation"])
            assert list(p4) == c["p4"]
            assert np.array_equal(Xp, z[c["id"] + "/out"])
        elif k == "dilate":
            assert np.array_equal(O.dilate(z[c["id"] + "/X"], c["d"]), z[c["id"] + "/out"])
        elif k == "conv":
            i = c["id"]
            X, W = z[i + "/X"], z[i + "/W"]
            pad = to_pad(c["pad"])
            X_col, p4 = O.im2col(X, W.shape, pad, c["stride"], c["dilation"])
            assert list(p4) == c["p4"]
            assert np.array_equal(X_col, z[i + "/X_col"])
            close(O.conv2D(X, W, c["stride"], pad, c["dilation"]), z[i + "/Z"])
            close(O.conv2D_loops(X, W, c["stride"], pad, c["dilation"]), z[i + "/Z"])
            back = O.col2im(z[i + "/C"], X.shape, W.shape, tuple(c["p4"]), c["stride"], c["dilation"])
            close(back, z[i + "/col2im"])
            assert O.calc_conv_out_dims(X.shape, W.shape, c["stride"], tuple(c["p4"]), c["dilation"]) == z[i + "/Z"].shape
        elif k == "deconv":
            i = c["id"]
            close(O.deconv2D_naive(z[i + "/X"], z[i + "/W"], c["stride"], to_pad(c["pad"]), 0), z[i + "/Z"])
        else:
            raise AssertionError(k)


def test_activations_golden():
    z, cases = load_golden("activations.npz")
    x = z["x"]
    for c in cases:
        close(O.act_fn(c["act"], x, c["p0"], c["p1"]), z[c["act"] + "/fn"], 1e-15)
        close(O.act_grad(c["act"], x, c["p0"], c["p1"]), z[c["act"] + "/grad"], 1e-15)


def test_conv2d_layer_golden():
    z, cases = load_golden("conv2d_layer.npz")
    for c in cases:
        i, pad = c["id"], to_pad(c["pad"])
        a = (c["act"], c["p0"], c["p1"])
        X, W, b = z[i + "/X"], z[i + "/W"], z[i + "/b"]
        Y, Z = O.conv2d_layer_fwd(X, W, b, c["stride"], pad, c["dilation"], *a)
        close(Y, z[i + "/Y"])
        close(Z, z[i + "/Z"])
        dX, dW, db = O.conv2d_layer_bwd(z[i + "/dLdy"], X, Z, W, c["stride"], pad, c["dilation"], *a)
        close(dX, z[i + "/dX"])
        if c["two_calls"]:
            X2 = z[i + "/X2"]
            Y2, Z2 = O.conv2d_layer_fwd(X2, W, b, c["stride"], pad, c["dilation"], *a)
            dX2, dW2, db2 = O.conv2d_layer_bwd(z[i + "/dLdy2"], X2, Z2, W, c["stride"], pad, c["dilation"], *a)
            close(dX2, z[i + "/dX2"])
            dW, db = dW + dW2, db + db2
        close(dW, z[i + "/dW"])
        close(db, z[i + "/db"])


def test_deconv2d_layer_golden():
    z, cases = load_golden("deconv2d_layer.npz")
    for c in cases:
        i, pad = c["id"], to_pad(c["pad"])
        a = (c["act"], c["p0"], c["p1"])
        X, W, b = z[i + "/X"], z[i + "/W"], z[i + "/b"]
        Y, Z = O.deconv2d_layer_fwd(X, W, b, c["stride"], pad, *a)
        close(Y, z[i + "/Y"])
        dX, dW, db = O.deconv2d_layer_bwd(z[i + "/dLdy"], X, Z, W, c["stride"], pad, *a)
        close(dX, z[i + "/dX"])
        if c["two_calls"]:
            X2 = z[i + "/X2"]
            Y2, Z2 = O.deconv2d_layer_fwd(X2, W, b, c["stride"], pad, *a)
            dX2, dW2, db2 = O.deconv2d_layer_bwd(z[i + "/dLdy2"], X2, Z2, W, c["stride"], pad, *a)
            close(dX2, z[i + "/dX2"])
            dW, db = dW + dW2, db + db2
        close(dW, z[i + "/dW"])
        close(db, z[i + "/db"])


def test_bn_add_golden():
    z, cases = load_golden("bn_add.npz")
    for c in cases:
        i = c["id"]
        if c["kind"] == "bn":
            ch = z[i + "/X"].shape[3]
            y, rm, rv = O.bn2d_fwd(z[i + "/X"], z[i + "/scaler"], z[i + "/intercept"], np.zeros(ch), np.ones(ch),
                                   c["momentum"], c["epsilon"])
            close(y, z[i + "/y"])
            close(rm, z[i + "/running_mean"])
            close(rv, z[i + "/running_var"])
            dX, dg, dh = O.bn2d_bwd(z[i + "/dLdy"], z[i + "/X"], z[i + "/scaler"], c["epsilon"])
            close(dX, z[i + "/dX"], 1e-10)
            close(dg, z[i + "/dScaler"])
            close(dh, z[i + "/dIntercept"])
            y2, _, _ = O.bn2d_fwd(z[i + "/X2"], z[i + "/scaler"], z[i + "/intercept"], rm, rv,
                                  c["momentum"], c["epsilon"], use_batch_stats=False)
            close(y2, z[i + "/y_infer"])
        else:
            Xs = [z[f"{i}/X{j}"] for j in range(c["n_in"])]
            Y, total = O.add_fwd(Xs, c["act"])
            close(Y, z[i + "/Y"])
            g = O.add_bwd(z[i + "/dLdY"], total, c["n_in"], c["act"])
            assert len(g) == c["n_in"]
            close(g[-1], z[i + "/dX"])


def test_skip_conv_module_golden():
    z, cases = load_golden("skip_conv_module.npz")
    for c in cases:
        i = c["id"]
        P = {}
        for t in "12s":
            ch = z[f"{i}/g{t}"].shape[0]
            P.update({"W" + t: z[f"{i}/W{t}"], "b" + t: z[f"{i}/b{t}"], "g" + t: z[f"{i}/g{t}"], "h" + t: z[f"{i}/h{t}"],
                      "rm" + t: np.zeros(ch), "rv" + t: np.ones(ch)})
        o = O.skip_conv_module_fwd_bwd(z[i + "/X"], z[i + "/dLdY"], P, tuple(c["kernel_shape1"]), tuple(c["kernel_shape2"]),
                                       tuple(c["kernel_shape_skip"]), to_pad(c["pad1"]), to_pad(c["pad2"]), c["stride1"],
                                       c["stride2"], c["stride_skip"], c["act"], c["p0"], c["p1"])
        assert list(o["pad_skip"]) == list(z[i + "/pad_skip"])
        for k in ["Y", "dX", "conv1_out", "conv2_out", "batchnorm1_out", "batchnorm2_out", "conv_skip_out",
                  "batchnorm_skip_out", "dLdBn1", "dLdBn2", "dLdConv1", "dLdConv2", "dLdBnSkip", "dLdConvSkip"]:
            close(o[k], z[f"{i}/{k}"], 1e-9)
        for t in "12s":
            for k in ("dW", "db", "dg", "dh", "rm", "rv"):
                close(np.asarray(o[k + t]).reshape(z[f"{i}/{k}{t}"].shape), z[f"{i}/{k}{t}"], 1e-9)


def test_oracle_error_conventions():
    X = np.zeros((1, 3, 3, 1))
    with pytest.raises(ValueError):
        O.im2col(X, (5, 5, 1, 1), 0, 1)                       # kernel > padded input  (utils.py:463-467)
    with pytest.raises(TypeError):
        O.col2im(np.zeros((9, 1)), X.shape, (3, 3, 1, 1), 1, 1)     # pad must be a 4-tuple (utils.py:577-578)
    with pytest.raises(ValueError):
        O.calc_pad_dims_2D([1, 3, 3, 1], (3, 3), (3, 3), 1)   # X_shape not a tuple   (utils.py:77-87)
    with pytest.raises(ValueError):
        O.calc_pad_dims_2D((1, 9, 9, 1), (3, 3), (3, 3), 1)   # negative pad          (utils.py:116-119)


# ---------------------------------------------------------------------------------------------------
#  optimizers + learning-rate schedules
# ---------------------------------------------------------------------------------------------------
def _sched_lr(sched, base_lr, step):
    from oracle import opt_oracle as P
    if sched is None:
        return P.lr_constant(base_lr, step)
    if sched["kind"] == "exponential":
        return P.lr_exponential(sched["initial_lr"], sched["stage_length"], sched["staircase"], sched["decay"], step)
    return P.lr_noam(sched["model_dim"], sched["scale_factor"], sched["warmup_steps"], step)


def test_optimizer_oracle_matches_reference_golden():
    from oracle import opt_oracle as P
    z, cases = load_golden("optimizers.npz")
    assert {c["kind"] for c in cases} == {"SGD", "AdaGrad", "RMSProp", "Adam"}
    for c in cases:
        cid, kw = c["id"], c["kwargs"]
        param = z[cid + "/param0"]
        state = {"cache": np.zeros_like(param), "mean": np.zeros_like(param), "var": np.zeros_like(param), "t": 0}
        for step in range(4):
            lr = _sched_lr(c["sched"], kw["lr"], step + 1)
            assert abs(lr - c["lrs"][step]) <= 1e-15 * max(1.0, abs(lr)), (cid, step, lr, c["lrs"][step])
            g = z["{}/grad{}".format(cid, step)]
            cn = kw.get("clip_norm")
            if c["kind"] == "SGD":
                param, state["cache"] = P.sgd_step(param, g, state["cache"], lr, kw["momentum"], cn)
            elif c["kind"] == "AdaGrad":
                param, state["cache"] = P.adagrad_step(param, g, state["cache"], lr, kw["eps"], cn)
            elif c["kind"] == "RMSProp":
                param, state["cache"] = P.rmsprop_step(param, g, state["cache"], lr, kw["decay"], kw["eps"], cn)
            else:
                param, state["mean"], state["var"], state["t"] = P.adam_step(
                    param, g, state["mean"], state["var"], state["t"], lr, kw["decay1"], kw["decay2"], kw["eps"], cn)
            assert rel_err(param, z["{}/param{}".format(cid, step + 1)]) < 1e-13, (cid, step)
            if c["kind"] == "Adam":
                assert rel_err(state["var"], z["{}/var{}".format(cid, step + 1)]) < 1e-13
                assert rel_err(state["mean"], z["{}/mean{}".format(cid, step + 1)]) < 1e-13
            else:
                assert rel_err(state["cache"], z["{}/cache{}".format(cid, step + 1)]) < 1e-13


def test_host_optimizer_strings_and_schedules_match_reference():
    """The product's optimizer / scheduler objects print and schedule exactly like the reference's (no GPU needed)."""
    pkg = npml()
    Op, Sc = pkg.optimizers, pkg.schedulers
    _, cases = load_golden("optimizers.npz")
    for c in cases:
        kwargs = dict(c["kwargs"])
        if c["sched"] is not None:
            sk = {k: v for k, v in c["sched"].items() if k != "kind"}
            kwargs["lr_scheduler"] = (Sc.ExponentialScheduler if c["sched"]["kind"] == "exponential"
                                      else Sc.NoamScheduler)(**sk)
        opt = getattr(Op, c["kind"])(**kwargs)
        assert str(opt) == c["string"]
        for step in range(4):
            opt.step()
            assert abs(opt.lr_scheduler(opt.cur_step, None) - c["lrs"][step]) <= 1e-15
        again = Op.OptimizerInitializer(str(opt))()              # round trip through the string form
        assert type(again) is type(opt)
        for k, v in opt.hyperparameters.items():
            if k != "lr_scheduler":
                assert again.hyperparameters[k] == v, (k, again.hyperparameters[k], v)
        assert str(again.lr_scheduler) == str(opt.lr_scheduler)
    assert isinstance(Op.OptimizerInitializer(None)(), Op.SGD)
    with pytest.raises(NotImplementedError):
        Op.OptimizerInitializer("Lion(lr=0.1)")()
    with pytest.raises(NotImplementedError):
        Sc.SchedulerInitializer("KingScheduler(initial_lr=0.01, patience=1000, decay=0.99)")()


# ---------------------------------------------------------------------------------------------------
#  callers either side of the path: Conv1D, Pool2D / Flatten, identity skip module
# ---------------------------------------------------------------------------------------------------
def test_conv1d_layer_golden():
    from oracle import next_oracle as N
    z, cases = load_golden("conv1d_layer.npz")
    assert {"same", "causal"} <= {c["pad"] for c in cases if isinstance(c["pad"], str)}
    for c in cases:
        i, pad = c["id"], to_pad(c["pad"])
        X, W, b = z[i + "/X"], z[i + "/W"], z[i + "/b"]
        a = (c["act"], c["p0"], c["p1"])
        assert list(N.resolve_pad_1D(X.shape, pad, c["kernel_width"], c["stride"], c["dilation"])) == list(z[i + "/pad"])
        close(N.conv1D(X, W, c["stride"], pad, c["dilation"]), z[i + "/conv1D"])
        Y, Z = N.conv1d_layer_fwd(X, W, b, c["stride"], pad, c["dilation"], *a)
        close(Y, z[i + "/Y"])
        close(Z, z[i + "/Z"])
        dX, dW, db = N.conv1d_layer_bwd(z[i + "/dLdy"], X, Z, W, c["stride"], pad, c["dilation"], *a)
        close(dX, z[i + "/dX"])
        if c["two_calls"]:
            X2 = z[i + "/X2"]
            Y2, Z2 = N.conv1d_layer_fwd(X2, W, b, c["stride"], pad, c["dilation"], *a)
            close(Y2, z[i + "/Y2"])
            dX2, dW2, db2 = N.conv1d_layer_bwd(z[i + "/dLdy2"], X2, Z2, W, c["stride"], pad, c["dilation"], *a)
            close(dX2, z[i + "/dX2"])
            dW, db = dW + dW2, db + db2
        close(dW, z[i + "/dW"])
        close(db, z[i + "/db"])
    with pytest.raises(ValueError):
        N.calc_pad_dims_1D([2, 9, 3], 9, 3, 1)                 # X_shape not a tuple (utils.py:153-163)


def test_pool2d_flatten_golden():
    from oracle import next_oracle as N
    z, cases = load_golden("pool2d.npz")
    for c in cases:
        i = c["id"]
        if c.get("flatten"):
            Y = N.flatten_fwd(z[i + "/X"], c["keep_dim"])
            assert np.array_equal(Y, z[i + "/Y"])
            assert np.array_equal(z[i + "/dLdy"].reshape(z[i + "/X"].shape), z[i + "/dX"])
            continue
        ks, pad = tuple(c["kernel_shape"]), to_pad(c["pad"])
        close(N.pool2d_fwd(z[i + "/X"], ks, c["stride"], pad, c["mode"]), z[i + "/Y"], 1e-15)
        if c["backward"]:
            close(N.pool2d_bwd(z[i + "/dLdY"], z[i + "/X"], ks, c["stride"], pad, c["mode"]), z[i + "/dX"], 1e-15)


def test_skip_identity_module_golden():
    from oracle import next_oracle as N
    z, cases = load_golden("skip_identity_module.npz")
    for c in cases:
        i = c["id"]
        P = {k: z["{}/{}".format(i, k)] for k in ("W1", "b1", "W2", "b2", "g1", "h1", "g2", "h2")}
        for t in "12":
            P["rm" + t], P["rv" + t] = np.zeros_like(P["g" + t]), np.ones_like(P["g" + t])
        o = N.skip_identity_module_fwd_bwd(z[i + "/X"], z[i + "/dLdY"], P, tuple(c["kernel_shape1"]),
                                           tuple(c["kernel_shape2"]), 1, 1, c["act"], c["p0"], c["p1"])
        for k in ("Y", "dX", "conv1_out", "conv2_out", "batchnorm1_out", "batchnorm2_out", "dLdBn1", "dLdBn2",
                  "dLdConv1", "dLdConv2", "dW1", "db1", "dW2", "db2", "dg1", "dh1", "dg2", "dh2", "rm1", "rv1", "rm2",
                  "rv2"):
            close(o[k], z["{}/{}".format(i, k)], 1e-10)


def _vae_encoder_oracle(z):
    """Conv(5x5,'same',ReLU) -> MaxPool 2x2/2 -> Conv(5x5,'same',ReLU) -> MaxPool 2x2/2 -> Flatten -> FC(ReLU) -> FC
    (models/vae.py:142-189), forward + backward with the oracle's functions."""
    from oracle import next_oracle as N
    W = [z["vae/W{}".format(k)] for k in range(4)]
    b = [z["vae/b{}".format(k)] for k in range(4)]
    X = z["vae/X"]
    a1, z1 = O.conv2d_layer_fwd(X, W[0], b[0], 1, "same", 0, "relu")
    p1 = N.pool2d_fwd(a1, (2, 2), 2, 0, "max")
    a2, z2 = O.conv2d_layer_fwd(p1, W[1], b[1], 1, "same", 0, "relu")
    p2 = N.pool2d_fwd(a2, (2, 2), 2, 0, "max")
    f = N.flatten_fwd(p2)
    h, _ = N.fc_fwd(f, W[2], b[2], "relu")
    y, _ = N.fc_fwd(h, W[3], b[3])
    g, dW3, db3 = N.fc_bwd(z["vae/dLdY"], h, W[3], b[3])
    g, dW2, db2 = N.fc_bwd(g, f, W[2], b[2], "relu")
    g = N.pool2d_bwd(g.reshape(p2.shape), a2, (2, 2), 2, 0, "max")
    g, dW1, db1 = O.conv2d_layer_bwd(g, p1, z2, W[1], 1, "same", 0, "relu")
    g = N.pool2d_bwd(g, a1, (2, 2), 2, 0, "max")
    dX, dW0, db0 = O.conv2d_layer_bwd(g, X, z1, W[0], 1, "same", 0, "relu")
    return y, dX, [dW0, dW1, dW2, dW3], [db0, db1, db2, db3]


def test_multiply_fc_wavenet_vae_golden():
    from oracle import next_oracle as N
    z, cases = load_golden("wavenet_vae.npz")
    for c in cases:
        i = c["id"]
        if c["kind"] == "multiply":
            Xs = [z["{}/X{}".format(i, j)] for j in range(c["n_in"])]
            Y, prod = N.multiply_fwd(Xs, c["act"])
            close(Y, z[i + "/Y"])
            for j, g in enumerate(N.multiply_bwd(z[i + "/dLdY"], Xs, prod, c["act"])):
                close(g, z["{}/dX{}".format(i, j)])
        elif c["kind"] == "fc":
            X, W, b = z[i + "/X"], z[i + "/W"], z[i + "/b"]
            close(N.fc_fwd(X, W, b, c["act"])[0], z[i + "/Y"])
            dX, dW, db = N.fc_bwd(z[i + "/dLdy"], X, W, b, c["act"])
            close(dX, z[i + "/dX"]); close(dW, z[i + "/dW"]); close(db, z[i + "/db"])
        elif c["kind"] == "wavenet":
            P = {k: z["{}/{}".format(i, k)] for k in ("Wd", "bd", "W1", "b1")}
            o = N.wavenet_module_fwd_bwd(z[i + "/X_main"], z[i + "/X_skip"], z[i + "/dY_skip"],
                                         z[i + "/dY_main"] if c["with_main"] else None, P, c["dilation"])
            for k in ("Y_main", "Y_skip", "dX_main", "dX_skip", "conv_dilation_out", "tanh_out", "sigm_out",
                      "multiply_gate_out", "conv_1x1_out", "dLdTanh", "dLdSigmoid", "dLdConv_1x1", "dLdMultiply",
                      "dLdConv_dilation", "dWd", "dbd", "dW1", "db1"):
                close(o[k], z["{}/{}".format(i, k)], 1e-11)
        else:
            y, dX, dWs, dbs = _vae_encoder_oracle(z)
            close(y, z["vae/Y"], 1e-11)
            close(dX, z["vae/dX"], 1e-11)
            for k in range(4):
                close(dWs[k], z["vae/dW{}".format(k)], 1e-11)
                close(dbs[k], z["vae/db{}".format(k)], 1e-11)


def test_skip_modules_inference_golden():
    """retain_derived=False forwards of the skip-connection modules (every BatchNorm2D on its running statistics,
    layers/layers.py:1141-1146; modules/modules.py:905-937, 574-594) as produced by the unmodified reference: the
    layer-by-layer oracle chain that the fused conv -> BatchNorm2D epilogue is tested against on the GPU
    (tests/test_gpu_parity.py::test_skip_modules_fused_inference) must reproduce them."""
    z, cases = load_golden("skip_modules_inference.npz")

    def bn(x, i, t):
        return O.bn2d_fwd(x, z[f"{i}/g{t}"], z[f"{i}/h{t}"], z[f"{i}/rm{t}"], z[f"{i}/rv{t}"], use_batch_stats=False)[0]

    for c in cases:
        i, act = c["id"], c["act"]
        X = z[i + "/X"]
        pad1 = 1 if c["kind"] == "conv" else "same"
        a1, _ = O.conv2d_layer_fwd(X, z[i + "/W1"], z[i + "/b1"], 1, pad1, 0, act)
        a2, _ = O.conv2d_layer_fwd(bn(a1, i, "1"), z[i + "/W2"], z[i + "/b2"], 1, pad1, 0, "identity")
        h2 = bn(a2, i, "2")
        if c["kind"] == "conv":
            a_s, _ = O.conv2d_layer_fwd(X, z[i + "/Ws"], z[i + "/bs"], 1, 0, 0, "identity")
            short = bn(a_s, i, "s")
        else:
            short = X
        Y, _ = O.add_fwd([short, h2], act)
        assert rel_err(Y, z[i + "/Y"]) < 1e-12, i
        # a training forward moved the running statistics away from their initial (0, 1)
        assert np.abs(z[i + "/rm1"]).max() > 1e-3 and np.abs(z[i + "/rv1"] - 1.0).max() > 1e-3
