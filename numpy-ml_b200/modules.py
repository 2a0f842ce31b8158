"""SkipConnectionConvModule (and SkipConnectionIdentityModule, at the end of the file), the reference's "ResNet-like convolution shortcut" block
(numpy_ml/neural_nets/modules/modules.py:615-984): conv1 -> act -> BN1 -> conv2 -> BN2, a
conv_skip -> BN_skip projection of the input, Add, act.  Note the order Conv -> act -> BN
(modules.py:646), not Conv -> BN -> act."""
import ctypes as C
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib as L
from .activations import ActivationInitializer, Affine, Tanh, Sigmoid
from .layers import Conv2D, Conv1D, BatchNorm2D, Add, Multiply
from .layers import _dz_prep
from .utils import calc_pad_dims_2D


def _device_add(a, b):
    """a + b through npml_add_act_fwd (new tensor)."""
    out = torch.empty_like(a)
    arr = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr())
    with L.profiled("npml_add_act_fwd"):
        L.check(L.load().npml_add_act_fwd(2, arr, L.code_of(a), a.numel(), 0, 0.0, 0.0, L.ptr(out), None, L.stream()),
                "npml_add_act_fwd")
    return out


class _Lazy(object):
    """A derived variable that the fused passes never write to memory (e.g. batchnorm2_out inside
    Y = act(BN2(..) + BNskip(..))): computed from what WAS retained the first time somebody reads it."""

    def __init__(self, fn):
        self.fn, self.value = fn, None

    def get(self):
        if self.value is None:
            self.value = self.fn()
        return self.value


def _resolve(dv):
    return {k: (v.get() if isinstance(v, _Lazy) else v) for k, v in dv.items()}


def _pair_add_act(bn_a, xa, saved_a, bn_b, xb, saved_b, act, want_sum):
    """Y = act(BN_a(xa) + BN_b(xb)) in one pass (bn_b None: xb is added as it is).  Returns (sum or None, Y)."""
    ch = xa.shape[3]
    Y = torch.empty_like(xa)
    S = torch.empty_like(xa) if want_sum else None
    Pa = bn_a.parameters
    args_b = [None] * 4
    if bn_b is not None:
        Pb = bn_b.parameters
        args_b = [L.ptr(Pb["scaler"]), L.ptr(Pb["intercept"]), L.ptr(saved_b[0]), L.ptr(saved_b[1])]
    with L.profiled("npml_bn2d_pair_add_act_fwd"):
        L.check(L.load().npml_bn2d_pair_add_act_fwd(xa.numel() // ch, ch, L.ptr(xa), L.ptr(xb), L.code_of(xa),
                                                    L.ptr(Pa["scaler"]), L.ptr(Pa["intercept"]), L.ptr(saved_a[0]),
                                                    L.ptr(saved_a[1]), *args_b, act.code, act.p0, act.p1, L.ptr(S), L.ptr(Y),
                                                    L.stream()), "npml_bn2d_pair_add_act_fwd")
    return S, Y


def _pair_bwd(dY, S, act, bn_a, xa, saved_a, bn_b, xb, saved_b, retain_grads, want_dxb=True):
    """(dxa, dxb) = (BN_a'(g), BN_b'(g)) with g = dY * act'(S) formed on the fly; the parameter gradients of both
    BatchNorm2D layers accumulate (layers.py:1192-1215).  bn_b None: dxb = g (identity shortcut)."""
    lib = L.load()
    ch = xa.shape[3]
    rows = xa.numel() // ch
    ws = bn_a._ws(rows, ch)
    sums = torch.empty(3 * ch, dtype=torch.float64, device=xa.device)
    mb = [None, None] if bn_b is None else [L.ptr(saved_b[0]), L.ptr(saved_b[1])]
    dxa = torch.empty_like(xa)
    dxb = torch.empty_like(xa) if want_dxb else None
    Ga, Pa = bn_a.gradients, bn_a.parameters
    total = rows
    with L.profiled("npml_bn2d_pair_bwd"):
        L.check(lib.npml_bn2d_pair_bwd_stats(rows, ch, L.ptr(dY), L.ptr(S), act.code, act.p0, act.p1, L.ptr(xa),
                                             L.ptr(xb) if bn_b is not None else None, L.code_of(xa), L.ptr(saved_a[0]),
                                             L.ptr(saved_a[1]), mb[0], mb[1], L.ptr(sums), L.ptr(ws), ws.numel(), L.stream()),
                "npml_bn2d_pair_bwd_stats")
        glob = sums
        if bn_a.sync:
            from . import dist as D
            if D.world_size() > 1:
                glob = D.allreduce_sum_f64(sums.clone())
                total = rows * D.world_size()
        gb = [None] * 5
        if bn_b is not None:
            Gb, Pb = bn_b.gradients, bn_b.parameters
            gb = [L.ptr(Pb["scaler"]), L.ptr(saved_b[0]), L.ptr(saved_b[1]),
                  L.ptr(Gb["scaler"]) if retain_grads else None, L.ptr(Gb["intercept"]) if retain_grads else None]
        L.check(lib.npml_bn2d_pair_bwd_apply(rows, total, ch, L.ptr(dY), L.ptr(S), act.code, act.p0, act.p1, L.ptr(xa),
                                             L.ptr(xb) if bn_b is not None else None, L.code_of(xa), L.ptr(Pa["scaler"]),
                                             L.ptr(saved_a[0]), L.ptr(saved_a[1]), gb[0], gb[1], gb[2], L.ptr(sums),
                                             L.ptr(glob), L.ptr(dxa), L.ptr(dxb),
                                             L.ptr(Ga["scaler"]) if retain_grads else None,
                                             L.ptr(Ga["intercept"]) if retain_grads else None, gb[3], gb[4], 1, L.ptr(ws),
                                             ws.numel(), L.stream()), "npml_bn2d_pair_bwd_apply")
    return dxa, dxb


class ModuleBase(ABC):
    """modules.py:21-116."""

    def __init__(self):
        self.X = None
        self.trainable = True
        super().__init__()

    @abstractmethod
    def forward(self, z, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def backward(self, out, **kwargs):
        raise NotImplementedError

    @property
    def components(self):
        return [getattr(self, c) for c in self.hyperparameters["component_ids"] if hasattr(self, c)]

    def freeze(self):
        self.trainable = False
        for c in self.components:
            c.freeze()

    def unfreeze(self):
        self.trainable = True
        for c in self.components:
            c.unfreeze()

    def update(self, cur_loss=None):
        assert self.trainable, "Layer is frozen"
        for c in self.components:
            c.update(cur_loss)
        self.flush_gradients()

    def flush_gradients(self):
        """Unlike modules.py:69-71 (which sets the components' derived variables to None and so breaks
        the next forward), the components are flushed to empty lists."""
        assert self.trainable, "Layer is frozen"
        self.X = []
        self._dv = {}
        self._tail = None
        comps = self.components
        buckets = {id(getattr(c, "_bucket", None)) for c in comps if c.gradients}
        shared = getattr(comps[0], "_bucket", None) if len(buckets) == 1 else None
        if shared is not None and shared.owns_only(comps):
            shared.flat.zero_()                      # every component gradient in one memset
            for c in comps:
                c.flush_gradients(zero=False)
        else:
            for c in comps:
                c.flush_gradients()

    def set_params(self, summary_dict):
        cids = self.hyperparameters["component_ids"]
        for key in ("parameters", "hyperparameters"):
            for c, cd in summary_dict.get(key, {}).get("components", {}).items():
                if c in cids and cd is not None and hasattr(self, c):
                    getattr(self, c).set_params(cd)
        return self

    def summary(self):
        return {"parameters": self.parameters, "layer": self.hyperparameters["layer"],
                "hyperparameters": self.hyperparameters}


class SkipConnectionConvModule(ModuleBase):
    # inference (retain_derived=False) forwards fold each BatchNorm2D into the epilogue of the convolution before it;
    # set to False to run the seven layers one by one as in training
    fuse_inference = True
    # training forwards / backwards run the residual tail -- batchnorm2, batchnorm_skip, add3 and its activation
    # (modules.py:931-937, 941-953) -- as ONE pass each way, and fold conv1's activation derivative into batchnorm1's
    # backward; the tensors that are then never written (batchnorm2_out, batchnorm_skip_out, dLdBn2, dLdBnSkip,
    # dLdConv1) are materialised on demand by `derived_variables`.  Set to False to run the seven layers one by one.
    fuse_training = True

    def __init__(self, out_ch1, out_ch2, kernel_shape1, kernel_shape2, kernel_shape_skip, pad1=0, pad2=0, stride1=1,
                 stride2=1, act_fn=None, epsilon=1e-5, momentum=0.9, stride_skip=1, optimizer=None,
                 init="glorot_uniform", mode="fp32", sync_bn=False):
        super().__init__()
        self.init = init
        self.pad1 = pad1
        self.pad2 = pad2
        self.in_ch = None
        self.out_ch1 = out_ch1
        self.out_ch2 = out_ch2
        self.epsilon = epsilon
        self.stride1 = stride1
        self.stride2 = stride2
        self.momentum = momentum
        self.optimizer = optimizer
        self.stride_skip = stride_skip
        self.kernel_shape1 = kernel_shape1
        self.kernel_shape2 = kernel_shape2
        self.kernel_shape_skip = kernel_shape_skip
        self.act_fn = Affine(slope=1, intercept=0) if act_fn is None else ActivationInitializer(act_fn)()
        self.mode = mode
        self.sync_bn = sync_bn      # extension: all-reduce the BatchNorm statistics over the data-parallel ranks
        self._init_params()

    def _init_params(self, X=None):
        self._dv = {}
        self.conv1 = Conv2D(pad=self.pad1, init=self.init, act_fn=self.act_fn, out_ch=self.out_ch1,
                            stride=self.stride1, optimizer=self.optimizer, kernel_shape=self.kernel_shape1,
                            mode=self.mode)
        self.conv2 = Conv2D(pad=self.pad2, init=self.init, out_ch=self.out_ch2, stride=self.stride2,
                            optimizer=self.optimizer, kernel_shape=self.kernel_shape2,
                            act_fn=Affine(slope=1, intercept=0), mode=self.mode)
        self.batchnorm1 = BatchNorm2D(epsilon=self.epsilon, momentum=self.momentum, sync=self.sync_bn)
        self.batchnorm2 = BatchNorm2D(epsilon=self.epsilon, momentum=self.momentum, sync=self.sync_bn)
        self.batchnorm_skip = BatchNorm2D(epsilon=self.epsilon, momentum=self.momentum, sync=self.sync_bn)
        self.add3 = Add(self.act_fn)
        self.add3._share_grads = True       # the module only reads the branch gradients (no in-place update)

    def _calc_skip_padding(self, X_shape):
        """Padding of the skip conv so that it matches conv2's output dims (modules.py:755-788).
        Like the reference, pad1/pad2 must be ints or 2-tuples (anything else fails here)."""
        pads = []
        for p in [self.pad1, self.pad2]:
            if isinstance(p, int):
                pads.append((p, p, p, p))
            elif isinstance(p, tuple) and len(p) == 2:
                pads.append((p[0], p[0], p[1], p[1]))
        self.pad1, self.pad2 = pads          # raises ValueError for 'same' / 4-tuples, as upstream
        _, in_rows, in_cols, _ = X_shape
        s1, (fr1, fc1) = self.stride1, self.kernel_shape1
        pr11, pr12, pc11, pc12 = self.pad1
        out_rows1 = int(np.floor(1 + (in_rows + pr11 + pr12 - fr1) / s1))
        out_cols1 = int(np.floor(1 + (in_cols + pc11 + pc12 - fc1) / s1))
        s2, (fr2, fc2) = self.stride2, self.kernel_shape2
        pr21, pr22, pc21, pc22 = self.pad2
        out_rows2 = int(np.floor(1 + (out_rows1 + pr21 + pr22 - fr2) / s2))
        out_cols2 = int(np.floor(1 + (out_cols1 + pc21 + pc22 - fc2) / s2))
        self.pad_skip = calc_pad_dims_2D(tuple(X_shape), (out_rows2, out_cols2), stride=self.stride_skip,
                                         kernel_shape=self.kernel_shape_skip)

    def _init_conv_skip(self, X_shape):
        self._calc_skip_padding(X_shape)
        self.conv_skip = Conv2D(init=self.init, pad=self.pad_skip, out_ch=self.out_ch2, stride=self.stride_skip,
                                kernel_shape=self.kernel_shape_skip, act_fn=Affine(slope=1, intercept=0),
                                optimizer=self.optimizer, mode=self.mode)

    def _by_component(self, attr):
        names = ["add3", "conv1", "conv2", "conv_skip", "batchnorm1", "batchnorm2", "batchnorm_skip"]
        return {n: (getattr(getattr(self, n), attr) if hasattr(self, n) else None) for n in names}

    @property
    def parameters(self):
        return {"components": self._by_component("parameters")}

    @property
    def gradients(self):
        return {"components": self._by_component("gradients")}

    @property
    def hyperparameters(self):
        return {"layer": "SkipConnectionConvModule", "init": self.init, "pad1": self.pad1, "pad2": self.pad2,
                "in_ch": self.in_ch, "out_ch1": self.out_ch1, "out_ch2": self.out_ch2, "epsilon": self.epsilon,
                "stride1": self.stride1, "stride2": self.stride2, "momentum": self.momentum,
                "act_fn": str(self.act_fn), "stride_skip": self.stride_skip, "kernel_shape1": self.kernel_shape1,
                "kernel_shape2": self.kernel_shape2, "kernel_shape_skip": self.kernel_shape_skip,
                "pad_skip": self.pad_skip if hasattr(self, "pad_skip") else None,
                "component_ids": ["add3", "conv1", "conv2", "conv_skip", "batchnorm1", "batchnorm2", "batchnorm_skip"],
                "components": self._by_component("hyperparameters")}

    @property
    def derived_variables(self):
        dv = {"conv1_out": None, "conv2_out": None, "conv_skip_out": None, "batchnorm1_out": None,
              "batchnorm2_out": None, "batchnorm_skip_out": None, "components": self._by_component("derived_variables")}
        dv.update(_resolve(self._dv))
        return dv

    def _fusable(self):
        comps = [self.conv1, self.conv2, self.conv_skip, self.batchnorm1, self.batchnorm2, self.batchnorm_skip]
        return self.fuse_training and self.trainable and all(c.trainable for c in comps) and len(self.conv1.X) == 0

    def forward(self, X, retain_derived=True):
        """modules.py:905-937."""
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, L.torch_dtype(self.mode))
        if not hasattr(self, "conv_skip"):
            self._init_conv_skip(tuple(Xd.shape))
            self.in_ch = Xd.shape[3]
        self._tail = None
        if retain_derived and self._fusable():
            # every convolution leaves the [sum, sum of squares] of its output for the BatchNorm2D behind it
            # (npml_conv2d_fprop_stats): no statistics pass over conv1_out / conv2_out / conv_skip_out
            q1, q2, qs = (torch.empty(2 * c.out_ch, dtype=torch.float64, device=Xd.device)
                          for c in (self.conv1, self.conv2, self.conv_skip))
            conv1_out = self.conv1.forward(Xd, True, stats_out=q1)
            bn1_out = self.batchnorm1.forward(conv1_out, True, sums=q1)
            conv2_out = self.conv2.forward(bn1_out, True, stats_out=q2)
            conv_skip_out = self.conv_skip.forward(Xd, True, stats_out=qs)
            bn2, bns, act = self.batchnorm2, self.batchnorm_skip, self.act_fn
            s2 = bn2.forward_stats(conv2_out, sums=q2)
            ss = bns.forward_stats(conv_skip_out, sums=qs)
            # ReLU / identity need no stored sum: Y > 0 <=> sum > 0
            want_sum = act.code not in (0, 2)
            S, Y = _pair_add_act(bn2, conv2_out, s2, bns, conv_skip_out, ss, act, want_sum)
            self._tail = (S if want_sum else Y, conv2_out, s2, conv_skip_out, ss)
            self.add3.X.append([None, None])                 # the Add layer's own bookkeeping (layers.py:700-703)
            self.add3.derived_variables["sum"].append(S if want_sum else None)
            self._dv.update({"conv1_out": conv1_out, "conv2_out": conv2_out, "conv_skip_out": conv_skip_out,
                             "batchnorm1_out": bn1_out,
                             "batchnorm2_out": _Lazy(lambda: bn2.normalise_with(conv2_out, s2)),
                             "batchnorm_skip_out": _Lazy(lambda: bns.normalise_with(conv_skip_out, ss))})
            return L.to_numpy(Y) if as_np else Y
        if not retain_derived and self.fuse_inference:
            # inference: every BatchNorm2D is the affine of its running statistics (layers.py:1141-1146) and nothing is
            # retained, so each conv -> BN pair is ONE kernel (bias, activation and BN affine in the epilogue)
            bn1_out = self.conv1.forward(Xd, False, fused_bn=self.batchnorm1)
            bn2_out = self.conv2.forward(bn1_out, False, fused_bn=self.batchnorm2)
            bn_skip_out = self.conv_skip.forward(Xd, False, fused_bn=self.batchnorm_skip)
            Y = self.add3.forward([bn_skip_out, bn2_out], False)
            return L.to_numpy(Y) if as_np else Y
        conv1_out = self.conv1.forward(Xd, retain_derived)
        bn1_out = self.batchnorm1.forward(conv1_out, retain_derived)
        conv2_out = self.conv2.forward(bn1_out, retain_derived)
        bn2_out = self.batchnorm2.forward(conv2_out, retain_derived)
        conv_skip_out = self.conv_skip.forward(Xd, retain_derived)
        bn_skip_out = self.batchnorm_skip.forward(conv_skip_out, retain_derived)
        Y = self.add3.forward([bn_skip_out, bn2_out], retain_derived)
        if retain_derived:
            self._dv.update({"conv1_out": conv1_out, "conv2_out": conv2_out, "conv_skip_out": conv_skip_out,
                             "batchnorm1_out": bn1_out, "batchnorm2_out": bn2_out,
                             "batchnorm_skip_out": bn_skip_out})
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdY, retain_grads=True):
        """modules.py:939-984."""
        as_np = isinstance(dLdY, np.ndarray)
        dY = L.to_device(dLdY, L.torch_dtype(self.mode))
        if getattr(self, "_tail", None) is not None:
            S, conv2_out, s2, conv_skip_out, ss = self._tail
            act = self.act_fn
            dConv2_out, dConvskip_out = _pair_bwd(dY, S, act, self.batchnorm2, conv2_out, s2, self.batchnorm_skip,
                                                  conv_skip_out, ss, retain_grads)
            dX = self.conv_skip.backward(dConvskip_out, retain_grads)
            dBn1_out = self.conv2.backward(dConv2_out, retain_grads)
            a1 = self.conv1.act_fn
            z1 = None if a1.code in (0, 2) else self.conv1.derived_variables["Z"][-1]
            dZ1 = self.batchnorm1.backward_act(dBn1_out, a1, z1, retain_grads)
            dX = _device_add(dX, self.conv1.backward_from_dz(dZ1, retain_grads))
            if retain_grads:
                bn1 = self.batchnorm1
                g = _Lazy(lambda: _dz_prep(dY, S, act, dY.dtype) if act.code != 0 else dY)
                self._dv.update({"dLdAdd3_X": dX, "dLdBn2": g, "dLdBn1": dBn1_out, "dLdConv2": dConv2_out,
                                 "dLdConv1": _Lazy(lambda: bn1._bwd(dBn1_out, bn1.X[-1], bn1._saved[-1], False)),
                                 "dLdBnSkip": g, "dLdConvSkip": dConvskip_out})
            return L.to_numpy(dX) if as_np else dX
        dBnskip_out, dBn2_out = self.add3.backward(dY)
        dConvskip_out = self.batchnorm_skip.backward(dBnskip_out)
        dX = self.conv_skip.backward(dConvskip_out)
        dConv2_out = self.batchnorm2.backward(dBn2_out)
        dBn1_out = self.conv2.backward(dConv2_out)
        dConv1_out = self.batchnorm1.backward(dBn1_out)
        dX = _device_add(dX, self.conv1.backward(dConv1_out))
        if retain_grads:
            self._dv.update({"dLdAdd3_X": dX, "dLdBn2": dBn2_out, "dLdBn1": dBn1_out, "dLdConv2": dConv2_out,
                             "dLdConv1": dConv1_out, "dLdBnSkip": dBnskip_out, "dLdConvSkip": dConvskip_out})
        return L.to_numpy(dX) if as_np else dX


class SkipConnectionIdentityModule(ModuleBase):
    """modules/modules.py:360-612: conv1('same', act) -> BN1 -> conv2('same', affine, out_ch = in_ch) -> BN2 ->
    Add([X, BN2]) -> act.  Same kernels as SkipConnectionConvModule, the input itself is the shortcut."""
    fuse_inference = True
    fuse_training = True          # see SkipConnectionConvModule: Y = act(X + BN2(conv2_out)) and its backward in one pass each

    def __init__(self, out_ch, kernel_shape1, kernel_shape2, stride1=1, stride2=1, act_fn=None, epsilon=1e-5,
                 momentum=0.9, optimizer=None, init="glorot_uniform", mode="fp32", sync_bn=False):
        super().__init__()
        self.init = init
        self.in_ch = None
        self.out_ch = out_ch
        self.epsilon = epsilon
        self.stride1 = stride1
        self.stride2 = stride2
        self.optimizer = optimizer
        self.momentum = momentum
        self.kernel_shape1 = kernel_shape1
        self.kernel_shape2 = kernel_shape2
        self.act_fn = Affine(slope=1, intercept=0) if act_fn is None else ActivationInitializer(act_fn)()
        self.mode = mode
        self.sync_bn = sync_bn
        self._init_params()

    def _init_params(self):
        self._dv = {}
        self.conv1 = Conv2D(pad="same", init=self.init, out_ch=self.out_ch, act_fn=self.act_fn, stride=self.stride1,
                            optimizer=self.optimizer, kernel_shape=self.kernel_shape1, mode=self.mode)
        # conv2 needs the input's channel count: created on the first forward, like the reference (modules.py:445-459)
        self.batchnorm1 = BatchNorm2D(epsilon=self.epsilon, momentum=self.momentum, sync=self.sync_bn)
        self.batchnorm2 = BatchNorm2D(epsilon=self.epsilon, momentum=self.momentum, sync=self.sync_bn)
        self.add3 = Add(self.act_fn)
        self.add3._share_grads = True       # the module only reads the branch gradients (no in-place update)

    def _init_conv2(self):
        self.conv2 = Conv2D(pad="same", init=self.init, out_ch=self.in_ch, stride=self.stride2,
                            optimizer=self.optimizer, kernel_shape=self.kernel_shape2,
                            act_fn=Affine(slope=1, intercept=0), mode=self.mode)

    _names = ["add3", "conv1", "conv2", "batchnorm1", "batchnorm2"]

    def _by_component(self, attr):
        return {n: (getattr(getattr(self, n), attr) if hasattr(self, n) else None) for n in self._names}

    @property
    def parameters(self):
        return {"components": self._by_component("parameters")}

    @property
    def gradients(self):
        return {"components": self._by_component("gradients")}

    @property
    def hyperparameters(self):
        return {"layer": "SkipConnectionIdentityModule", "init": self.init, "in_ch": self.in_ch, "out_ch": self.out_ch,
                "epsilon": self.epsilon, "stride1": self.stride1, "stride2": self.stride2, "momentum": self.momentum,
                "optimizer": self.optimizer, "act_fn": str(self.act_fn), "kernel_shape1": self.kernel_shape1,
                "kernel_shape2": self.kernel_shape2,
                "component_ids": ["conv1", "batchnorm1", "conv2", "batchnorm2", "add3"],
                "components": self._by_component("hyperparameters")}

    @property
    def derived_variables(self):
        dv = {"conv1_out": None, "conv2_out": None, "batchnorm1_out": None, "batchnorm2_out": None,
              "components": self._by_component("derived_variables")}
        dv.update(_resolve(self._dv))
        return dv

    def _fusable(self):
        comps = [self.conv1, self.conv2, self.batchnorm1, self.batchnorm2]
        return self.fuse_training and self.trainable and all(c.trainable for c in comps) and len(self.conv1.X) == 0

    def forward(self, X, retain_derived=True):
        """modules.py:574-594."""
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, L.torch_dtype(self.mode))
        if not hasattr(self, "conv2"):
            self.in_ch = Xd.shape[3]
            self._init_conv2()
        if not retain_derived and self.fuse_inference:      # see SkipConnectionConvModule.forward
            bn1_out = self.conv1.forward(Xd, False, fused_bn=self.batchnorm1)
            bn2_out = self.conv2.forward(bn1_out, False, fused_bn=self.batchnorm2)
            Y = self.add3.forward([Xd, bn2_out], False)
            return L.to_numpy(Y) if as_np else Y
        self._tail = None
        if retain_derived and self._fusable():
            q1, q2 = (torch.empty(2 * c.out_ch, dtype=torch.float64, device=Xd.device) for c in (self.conv1, self.conv2))
            conv1_out = self.conv1.forward(Xd, True, stats_out=q1)
            bn1_out = self.batchnorm1.forward(conv1_out, True, sums=q1)
            conv2_out = self.conv2.forward(bn1_out, True, stats_out=q2)
            bn2, act = self.batchnorm2, self.act_fn
            s2 = bn2.forward_stats(conv2_out, sums=q2)
            want_sum = act.code not in (0, 2)
            S, Y = _pair_add_act(bn2, conv2_out, s2, None, Xd, None, act, want_sum)
            self._tail = (S if want_sum else Y, conv2_out, s2)
            self.add3.X.append([None, None])
            self.add3.derived_variables["sum"].append(S if want_sum else None)
            self._dv.update({"conv1_out": conv1_out, "conv2_out": conv2_out, "batchnorm1_out": bn1_out,
                             "batchnorm2_out": _Lazy(lambda: bn2.normalise_with(conv2_out, s2))})
            return L.to_numpy(Y) if as_np else Y
        conv1_out = self.conv1.forward(Xd, retain_derived)
        bn1_out = self.batchnorm1.forward(conv1_out, retain_derived)
        conv2_out = self.conv2.forward(bn1_out, retain_derived)
        bn2_out = self.batchnorm2.forward(conv2_out, retain_derived)
        Y = self.add3.forward([Xd, bn2_out], retain_derived)
        if retain_derived:
            self._dv.update({"conv1_out": conv1_out, "conv2_out": conv2_out, "batchnorm1_out": bn1_out,
                             "batchnorm2_out": bn2_out})
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdY, retain_grads=True):
        """modules.py:596-612 (dLdAdd3_X is the returned dX: the reference accumulates into it in place)."""
        as_np = isinstance(dLdY, np.ndarray)
        dY = L.to_device(dLdY, L.torch_dtype(self.mode))
        if getattr(self, "_tail", None) is not None:
            S, conv2_out, s2 = self._tail
            act = self.act_fn
            dConv2_out, dX0 = _pair_bwd(dY, S, act, self.batchnorm2, conv2_out, s2, None, None, None, retain_grads)
            dBn1_out = self.conv2.backward(dConv2_out, retain_grads)
            a1 = self.conv1.act_fn
            z1 = None if a1.code in (0, 2) else self.conv1.derived_variables["Z"][-1]
            dZ1 = self.batchnorm1.backward_act(dBn1_out, a1, z1, retain_grads)
            dX = _device_add(dX0, self.conv1.backward_from_dz(dZ1, retain_grads))
            bn1 = self.batchnorm1
            self._dv.update({"dLdAdd3_X": dX, "dLdBn2": dX0, "dLdBn1": dBn1_out, "dLdConv2": dConv2_out,
                             "dLdConv1": _Lazy(lambda: bn1._bwd(dBn1_out, bn1.X[-1], bn1._saved[-1], False))})
            return L.to_numpy(dX) if as_np else dX
        dX0, dBn2_out = self.add3.backward(dY, retain_grads)
        dConv2_out = self.batchnorm2.backward(dBn2_out, retain_grads)
        dBn1_out = self.conv2.backward(dConv2_out, retain_grads)
        dConv1_out = self.batchnorm1.backward(dBn1_out, retain_grads)
        dX = _device_add(dX0, self.conv1.backward(dConv1_out, retain_grads))
        self._dv.update({"dLdAdd3_X": dX, "dLdBn2": dBn2_out, "dLdBn1": dBn1_out, "dLdConv2": dConv2_out,
                         "dLdConv1": dConv1_out})
        return L.to_numpy(dX) if as_np else dX


class WavenetResidualModule(ModuleBase):
    """modules/modules.py:119-357: causal dilated Conv1D (width 2) -> tanh * sigmoid gate -> 1x1 Conv1D, added to the
    residual path and to the skip path.  Like the reference, `kernel_width` is stored but the dilated convolution
    always has width 2 (modules.py:161-169).  Unlike it, `X_skip=None` on the first block means zeros (the reference
    passes the None on to Add and fails, modules.py:297-299)."""

    def __init__(self, ch_residual, ch_dilation, dilation, kernel_width, optimizer=None, init="glorot_uniform",
                 mode="fp32"):
        super().__init__()
        self.init = init
        self.dilation = dilation
        self.optimizer = optimizer
        self.ch_residual = ch_residual
        self.ch_dilation = ch_dilation
        self.kernel_width = kernel_width
        self.mode = mode
        self._init_params()

    def _init_params(self):
        self._dv = {}
        ident = Affine(slope=1, intercept=0)
        self.conv_dilation = Conv1D(stride=1, pad="causal", init=self.init, kernel_width=2, dilation=self.dilation,
                                    out_ch=self.ch_dilation, optimizer=self.optimizer, act_fn=ident, mode=self.mode)
        self.tanh = Tanh()
        self.sigm = Sigmoid()
        self.multiply_gate = Multiply(act_fn=ident)
        self.conv_1x1 = Conv1D(stride=1, pad="same", dilation=0, init=self.init, kernel_width=1,
                               out_ch=self.ch_residual, optimizer=self.optimizer, act_fn=ident, mode=self.mode)
        self.add_residual = Add(act_fn=ident)
        self.add_skip = Add(act_fn=ident)

    _names = ["conv_1x1", "add_skip", "add_residual", "conv_dilation", "multiply_gate"]

    def _by_component(self, attr):
        return {n: getattr(getattr(self, n), attr) for n in self._names}

    @property
    def parameters(self):
        return {"components": self._by_component("parameters")}

    @property
    def gradients(self):
        return {"components": self._by_component("gradients")}

    @property
    def hyperparameters(self):
        return {"layer": "WavenetResidualModule", "init": self.init, "dilation": self.dilation,
                "optimizer": self.optimizer, "ch_residual": self.ch_residual, "ch_dilation": self.ch_dilation,
                "kernel_width": self.kernel_width, "component_ids": list(self._names),
                "components": self._by_component("hyperparameters")}

    @property
    def derived_variables(self):
        dv = {"conv_1x1_out": None, "conv_dilation_out": None, "multiply_gate_out": None,
              "components": self._by_component("derived_variables")}
        dv.update(self._dv)
        return dv

    def forward(self, X_main, X_skip=None):
        """modules.py:283-318.  Returns (Y_main, Y_skip)."""
        as_np = isinstance(X_main, np.ndarray)
        dt = L.torch_dtype(self.mode)
        Xm = L.to_device(X_main, dt)
        conv_dilation_out = self.conv_dilation.forward(Xm)
        tanh_gate = self.tanh.fn(conv_dilation_out)
        sigm_gate = self.sigm.fn(conv_dilation_out)
        multiply_gate_out = self.multiply_gate.forward([tanh_gate, sigm_gate])
        conv_1x1_out = self.conv_1x1.forward(multiply_gate_out)
        Xs = torch.zeros_like(conv_1x1_out) if X_skip is None else L.to_device(X_skip, dt)
        self.X_main, self.X_skip = Xm, Xs
        Y_skip = self.add_skip.forward([Xs, conv_1x1_out])
        Y_main = self.add_residual.forward([Xm, conv_1x1_out])
        self._dv.update({"tanh_out": tanh_gate, "sigm_out": sigm_gate, "conv_dilation_out": conv_dilation_out,
                         "multiply_gate_out": multiply_gate_out, "conv_1x1_out": conv_1x1_out})
        return (L.to_numpy(Y_main), L.to_numpy(Y_skip)) if as_np else (Y_main, Y_skip)

    def backward(self, dY_skip, dY_main=None):
        """modules.py:320-357.  Returns (dX_main, dX_skip)."""
        as_np = isinstance(dY_skip, np.ndarray)
        dt = L.torch_dtype(self.mode)
        dX_skip, dConv_1x1_out = self.add_skip.backward(L.to_device(dY_skip, dt))
        dX_main = None
        if dY_main is not None:
            dX_main, dConv_1x1_main = self.add_residual.backward(L.to_device(dY_main, dt))
            dConv_1x1_out = _device_add(dConv_1x1_out, dConv_1x1_main)
        dMultiply_out = self.conv_1x1.backward(dConv_1x1_out)
        dTanh_out, dSigm_out = self.multiply_gate.backward(dMultiply_out)
        z = self._dv["conv_dilation_out"]
        dTanh_in = _dz_prep(dTanh_out, z, self.tanh, z.dtype)
        dSigm_in = _dz_prep(dSigm_out, z, self.sigm, z.dtype)
        dDilation_out = _device_add(dTanh_in, dSigm_in)
        conv_back = self.conv_dilation.backward(dDilation_out)
        dX_main = conv_back if dX_main is None else _device_add(dX_main, conv_back)
        self._dv.update({"dLdTanh": dTanh_out, "dLdSigmoid": dSigm_out, "dLdConv_1x1": dConv_1x1_out,
                         "dLdMultiply": dMultiply_out, "dLdConv_dilation": dDilation_out})
        return (L.to_numpy(dX_main), L.to_numpy(dX_skip)) if as_np else (dX_main, dX_skip)
