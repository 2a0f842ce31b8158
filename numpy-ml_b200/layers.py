"""Conv2D / Deconv2D / Conv1D / Pool2D / Flatten / BatchNorm2D / Add / Multiply / FullyConnected with the reference's layer API
(numpy_ml/neural_nets/layers/layers.py): same constructor arguments, lazy parameter init on the
first forward, `parameters` / `gradients` / `hyperparameters` / `derived_variables` dicts, list-valued
`backward`, `+=` gradient accumulation, `flush_gradients` / `update` / `freeze` / `summary`.

Differences a user can see (documented in INTEGRATION.md):
  * tensors live on the GPU.  NumPy inputs are accepted and answered with float64 NumPy arrays;
    device tensors are answered with device tensors.  `parameters`, `gradients` and the retained
    `X` / `Z` are device tensors (fp32 parameters and gradients; activations in the mode's dtype).
  * one extra constructor keyword, `mode` in {"fp32", "tf32", "bf16"} (default "fp32", the
    exactness mode), selects the arithmetic of the convolution kernels.
"""
import ctypes as C
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib as L
from . import utils as U
from .activations import ActivationInitializer
from .optimizers import OptimizerInitializer


def _as_param(v):
    """Parameters may be assigned as NumPy arrays (like the reference); keep fp32 device tensors."""
    if isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 and v.is_contiguous():
        return v
    return L.to_device(v, torch.float32)


class LayerBase(ABC):
    """layers/layers.py:27-136."""

    def __init__(self, optimizer=None):
        self.X = []
        self.act_fn = None
        self.trainable = True
        self.optimizer = OptimizerInitializer(optimizer)()
        self.gradients = {}
        self.parameters = {}
        self.derived_variables = {}
        super().__init__()

    @abstractmethod
    def _init_params(self, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def forward(self, z, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def backward(self, out, **kwargs):
        raise NotImplementedError

    def freeze(self):
        self.trainable = False

    def unfreeze(self):
        self.trainable = True

    def flush_gradients(self, zero=True):
        """Zero the gradients and drop every retained input / derived variable (layers.py:66-74).
        Gradients that live in a dist.GradBucket are one contiguous slice: one memset instead of one per tensor
        (`zero=False`: the caller has already zeroed the whole bucket)."""
        assert self.trainable, "Layer is frozen"
        self.X = []
        for k in self.derived_variables:
            self.derived_variables[k] = []
        if not zero:
            return
        bucket = getattr(self, "_bucket", None)
        if bucket is not None:
            bucket.zero_layer(self)
        else:
            for k, v in self.gradients.items():
                v.zero_()

    def update(self, cur_loss=None):
        """Optimizer step on the accrued gradients, then flush (layers.py:76-86)."""
        assert self.trainable, "Layer is frozen"
        self.optimizer.step()
        for k, v in self.gradients.items():
            if k in self.parameters:
                self.parameters[k] = self.optimizer(_as_param(self.parameters[k]), v, k, cur_loss)
        self._wgen = self.__dict__.get("_wgen", 0) + 1      # the kernels' packed copies of W are stale now
        self.flush_gradients()

    def set_params(self, summary_dict):
        """layers.py:88-128."""
        layer, sd = self, dict(summary_dict)
        for k in ["parameters", "hyperparameters"]:
            if k in sd:
                entry = sd.pop(k)
                sd.update(entry)
        for k, v in sd.items():
            if k in self.parameters:
                layer.parameters[k] = _as_param(v)
            if k in self.hyperparameters:
                if k == "act_fn":
                    layer.act_fn = ActivationInitializer(v)()
                elif k == "optimizer":
                    layer.optimizer = OptimizerInitializer(sd[k])()
                elif k not in ["wrappers", "optimizer"]:
                    setattr(layer, k, v)
        return layer

    def summary(self):
        return {"layer": self.hyperparameters["layer"], "parameters": self.parameters,
                "hyperparameters": self.hyperparameters}

    # ---- helpers shared by the layers ---------------------------------------------------------
    def _opt_hparams(self):
        return {"cache": self.optimizer.cache, "hyperparameters": self.optimizer.hyperparameters}


class WeightInitializer(object):
    """initializers/initializers.py:242-308 (consumes np.random exactly like the reference)."""

    def __init__(self, act_fn_str, mode="glorot_uniform"):
        if mode not in ["he_normal", "he_uniform", "glorot_normal", "glorot_uniform", "std_normal", "trunc_normal"]:
            raise ValueError("Unrecognize initialization mode: {}".format(mode))
        self.mode, self.act_fn = mode, act_fn_str

    def __call__(self, weight_shape):
        if self.mode == "glorot_uniform":
            return U.glorot_uniform(weight_shape, self._calc_glorot_gain())
        if self.mode == "glorot_normal":
            return U.glorot_normal(weight_shape, self._calc_glorot_gain())
        if self.mode == "he_uniform":
            return U.he_uniform(weight_shape)
        if self.mode == "he_normal":
            return U.he_normal(weight_shape)
        if self.mode == "std_normal":
            return np.random.randn(*weight_shape)
        return U.truncated_normal(0, 1, weight_shape)

    def _calc_glorot_gain(self):
        import re
        gain, s = 1.0, self.act_fn.lower()
        if s == "tanh":
            gain = 5.0 / 3.0
        elif s == "relu":
            gain = np.sqrt(2)
        elif "leaky relu" in s:
            alpha = re.match(r"leaky relu\(alpha=(.*)\)", s).groups()[0]
            gain = np.sqrt(2 / 1 + float(alpha) ** 2)
        return gain


def _dz_prep(dY, Z, act, out_dtype, db=None, accumulate=True):
    """dZ = dLdy * act'(Z) (+ db += column sums), one fused pass (layers.py:3103, 3109)."""
    lib = L.load()
    rows, ch = dY.numel() // dY.shape[-1], dY.shape[-1]
    identity = act.code == 0
    if identity and dY.dtype == out_dtype and db is None:
        return dY
    # identity activation in the GEMMs' own dtype: dZ IS dLdy, the pass only has to produce the column sums
    alias = identity and dY.dtype == out_dtype
    dZ = dY if alias else torch.empty(dY.shape, dtype=out_dtype, device=dY.device)
    ws = None
    if db is not None:
        d = U.make_desc(1, 1, 1, 1, ch, 1, 1, 1, 1, (0, 0, 0, 0), 1, 1)
        d.N, d.OH, d.OW = 1, 1, int(rows)
        ws = L.workspace(lib.npml_workspace_bytes(d, L.OP_DZ_PREP))
    with L.profiled("npml_dz_prep"):
        L.check(lib.npml_dz_prep(rows, ch, act.code, act.p0, act.p1, L.ptr(dY), L.code_of(dY),
                                 None if identity else L.ptr(Z), 0 if identity else L.code_of(Z),
                                 None if alias else L.ptr(dZ), L.code_of(dZ), L.ptr(db), 1 if accumulate else 0, L.ptr(ws),
                                 0 if ws is None else ws.numel(), L.stream()), "npml_dz_prep")
    return dZ


def _same_shapes(Xs, who):
    """The element-wise kernels index every input with the first one's element count: unequal shapes would read out
    of bounds (the reference's `out += X[i]` raises or broadcasts)."""
    for i, x in enumerate(Xs[1:], 1):
        if tuple(x.shape) != tuple(Xs[0].shape):
            raise ValueError("{}: input {} has shape {}, expected {}".format(who, i, tuple(x.shape), tuple(Xs[0].shape)))


class _ConvBase(LayerBase):
    """State handling shared by Conv2D and Deconv2D (they differ only in geometry and entry points)."""
    _layer_name = None
    _ops = None  # (fprop, dgrad, wgrad) op ids

    def _init_params(self):
        init_weights = WeightInitializer(str(self.act_fn), mode=self.init)
        fr, fc = self.kernel_shape
        W = init_weights((fr, fc, self.in_ch, self.out_ch))
        self.parameters = {"W": L.to_device(W, torch.float32),
                           "b": torch.zeros((1, 1, 1, self.out_ch), dtype=torch.float32, device=L.device())}
        self.gradients = {"W": torch.zeros_like(self.parameters["W"]), "b": torch.zeros_like(self.parameters["b"])}
        self.derived_variables = {"Z": [], "out_rows": [], "out_cols": []}
        self.is_initialized = True

    def _geometry(self, X_shape):
        raise NotImplementedError

    def _desc(self, X_shape, accumulate=0):
        n, h, w, c = X_shape
        fr, fc = self.kernel_shape
        p4, oh, ow, dil = self._geometry(X_shape)
        return U.make_desc(n, h, w, c, self.out_ch, fr, fc, self.stride, dil, p4, oh, ow, self.mode, self.act_fn,
                           accumulate)

    def _call(self, name, d, op, *args):
        lib = L.load()
        ws = L.workspace(lib.npml_workspace_bytes(d, op))
        with L.profiled(name):
            L.check(getattr(lib, name)(d, *args, L.ptr(ws), ws.numel(), L.stream()), name)

    _DESC_KEY = ("N", "H", "W", "C", "K", "fr", "fc", "stride", "dil", "pr1", "pr2", "pc1", "pc2", "OH", "OW", "mode")

    def _call_w(self, name, d, op, W, *args):
        """fprop / dgrad: the operand-type copy of the weights that the kernels read (Wt[tap][co][ci], bf16 / tf32) lives
        in a workspace this layer keeps per op, and is rebuilt only when W changed -- a new tensor, an in-place torch
        op (`_version`), or an optimizer step (`update()` bumps `_wgen`) -- instead of in front of every call
        (NPML_FLAG_WEIGHTS_PACKED).  While a CUDA graph is being captured the packed copy is reused only if the caller
        declared the weights constant across replays (`_lib.GRAPH_REUSE_PACKED`): a captured step that contains its own
        optimizer update must repack inside the graph."""
        lib = L.load()
        cache = self.__dict__.setdefault("_wpack", {})
        key = tuple(getattr(d, f) for f in self._DESC_KEY)
        gen = self.__dict__.get("_wgen", 0)
        ent = cache.get(op)
        hit = (ent is not None and ent[1] is W and ent[2] == W._version and ent[3] == gen and ent[4] == key)
        if hit and not L.GRAPH_REUSE_PACKED and torch.cuda.is_current_stream_capturing():
            hit = False
        if hit:
            ws = ent[0]
            d.flags |= L.FLAG_WEIGHTS_PACKED
        else:
            nbytes = int(lib.npml_workspace_bytes(d, op))
            ws = ent[0] if (ent is not None and ent[0].numel() >= nbytes) else \
                torch.empty(max(nbytes, 256), dtype=torch.uint8, device=W.device)
            cache[op] = (ws, W, W._version, gen, key)
        with L.profiled(name):
            L.check(getattr(lib, name)(d, *args, L.ptr(ws), ws.numel(), L.stream()), name)

    def _uses_tc(self, X_shape=None):
        """True if the forward of this layer runs on the tcgen05 kernels for the last (or given) input shape."""
        if X_shape is None:
            X_shape = tuple(self.X[-1].shape) if self.X else self._last_shape
        return bool(L.load().npml_uses_tensor_cores(self._desc(tuple(X_shape)), self._ops[0]))

    def forward(self, X, retain_derived=True, fused_bn=None, stats_out=None):
        """Y = act_fn(conv(X, W) + b); bias and activation are the kernel's epilogue
        (layers.py:3006-3046 / 3438-3478).

        `stats_out` (an extension, Conv2D only): a float64 device vector of 2 * out_ch entries that receives the
        per-channel [sum Y, sum Y^2] over (n_ex, rows, cols) -- the reductions of a training-mode BatchNorm2D that
        directly follows (layers.py:1146-1147) -- from the convolution's own epilogue (npml_conv2d_fprop_stats), so
        that BatchNorm2D does not have to read Y again for them (`BatchNorm2D.forward(Y, sums=stats_out)`).

        `fused_bn` (an extension): a BatchNorm2D that directly follows this layer and is in inference mode for this
        call (frozen, or `retain_derived=False`: layers.py:1141-1146 then normalises with the running statistics).
        Its per-channel affine becomes part of the same epilogue and the return value is BatchNorm2D(Y) -- exactly
        what `fused_bn.forward(self.forward(X))` would give, without the intermediate tensor."""
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, L.torch_dtype(self.mode))
        if not self.is_initialized:
            self.in_ch = Xd.shape[3]
            self._init_params()
        W = self.parameters["W"] = _as_param(self.parameters["W"])
        b = self.parameters["b"] = _as_param(self.parameters["b"])
        d = self._desc(tuple(Xd.shape))
        self._last_shape = tuple(Xd.shape)
        out_shape = (d.N, d.OH, d.OW, d.K)
        pre = "npml_conv2d_" if self._layer_name == "Conv2D" else "npml_deconv2d_"
        if fused_bn is not None:
            if fused_bn.trainable and retain_derived:
                raise ValueError("fused_bn needs a BatchNorm2D in inference mode (frozen or retain_derived=False)")
            scale, shift = fused_bn.folded_affine(d.K)
            Z = torch.empty(out_shape, dtype=Xd.dtype, device=Xd.device) if retain_derived else None
            Y = torch.empty(out_shape, dtype=Xd.dtype, device=Xd.device)
            self._call_w(pre + "fprop_bn", d, self._ops[0], W, L.ptr(Xd), L.ptr(W), L.ptr(b), L.ptr(scale), L.ptr(shift),
                         L.ptr(Z), L.ptr(Y))
        else:
            Z = torch.empty(out_shape, dtype=Xd.dtype, device=Xd.device)
            Y = Z if self.act_fn.code == 0 else torch.empty_like(Z)
            zy = (L.ptr(Z) if (retain_derived or Y is Z) else None, None if Y is Z else L.ptr(Y))
            if stats_out is not None:
                if self._layer_name != "Conv2D" or d.N == 0:
                    raise ValueError("stats_out is implemented for non-empty Conv2D forwards")
                assert stats_out.dtype == torch.float64 and stats_out.numel() == 2 * d.K
                self._call_w(pre + "fprop_stats", d, self._ops[0], W, L.ptr(Xd), L.ptr(W), L.ptr(b), zy[0], zy[1],
                             L.ptr(stats_out))
            else:
                self._call_w(pre + "fprop", d, self._ops[0], W, L.ptr(Xd), L.ptr(W), L.ptr(b), zy[0], zy[1])
        if retain_derived:
            self.X.append(Xd)
            self.derived_variables["Z"].append(Z)
            self.derived_variables["out_rows"].append(d.OH)
            self.derived_variables["out_cols"].append(d.OW)
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdy, retain_grads=True, index=None):
        """dX for every retained forward; dW / db accumulate into `gradients`
        (layers.py:3048-3116 / 3480-3568).  `index` (an extension) runs the backward of ONE retained forward,
        so that a pipelined caller can interleave the chunks of a batch with their host copies."""
        assert self.trainable, "Layer is frozen"
        if index is not None:
            as_np = isinstance(dLdy, np.ndarray)
            dx = self._bwd(L.to_device(dLdy, L.torch_dtype(self.mode)), self.X[index],
                           self.derived_variables["Z"][index], retain_grads)
            return L.to_numpy(dx) if as_np else dx
        if not isinstance(dLdy, list):
            dLdy = [dLdy]
        dX = []
        for dy, x, z in zip(dLdy, self.X, self.derived_variables["Z"]):
            as_np = isinstance(dy, np.ndarray)
            dx = self._bwd(L.to_device(dy, L.torch_dtype(self.mode)), x, z, retain_grads)
            dX.append(L.to_numpy(dx) if as_np else dx)
        return dX[0] if len(self.X) == 1 else dX

    def backward_from_dz(self, dZ, retain_grads=True, index=-1):
        """Backward of ONE retained forward from dZ = dLdy * act'(Z) itself: for callers that have already applied this
        layer's activation derivative inside another pass (BatchNorm2D.backward_act folds it into its dX pass)."""
        assert self.trainable, "Layer is frozen"
        from .activations import Identity
        return self._bwd(L.to_device(dZ, L.torch_dtype(self.mode)), self.X[index], self.derived_variables["Z"][index],
                         retain_grads, act=Identity())

    def _bwd(self, dLdy, X, Z, retain_grads=True, act=None):
        W = self.parameters["W"] = _as_param(self.parameters["W"])
        pre = "npml_conv2d_" if self._layer_name == "Conv2D" else "npml_deconv2d_"
        conv = self._layer_name == "Conv2D"
        # Conv2D: db is one more accumulator row of the wgrad pass; Deconv2D: db comes with the dZ prologue
        dZ = _dz_prep(dLdy, Z, self.act_fn if act is None else act, X.dtype,
                      self.gradients["b"] if (retain_grads and not conv) else None)
        d = self._desc(tuple(X.shape), accumulate=1)
        # parameter gradients first (the reference's order, layers.py:3106-3114): a data-parallel bucket can then
        # all-reduce dW / db on its own stream while dX is still being computed (dist.GradBucket, overlap=True)
        if retain_grads and conv and L.load().npml_conv2d_bwd_is_fused(d):
            # thin shapes (cfg1): dW, db and dX from one pass over dZ, one launch (csrc/thin_bwd.cu)
            dX = torch.empty_like(X)
            self._call("npml_conv2d_bwd", d, L.OP_CONV_BWD, L.ptr(X), L.ptr(dZ), L.ptr(W), L.ptr(self.gradients["W"]),
                       L.ptr(self.gradients["b"]), L.ptr(dX))
            hook = getattr(self, "_grads_ready_hook", None)
            if hook is not None:
                hook(self)
            return dX
        if retain_grads and conv:
            self._call("npml_conv2d_wgrad_db", d, self._ops[2], L.ptr(X), L.ptr(dZ), L.ptr(self.gradients["W"]),
                       L.ptr(self.gradients["b"]))
        elif retain_grads:
            self._call(pre + "wgrad", d, self._ops[2], L.ptr(X), L.ptr(dZ), L.ptr(self.gradients["W"]))
        hook = getattr(self, "_grads_ready_hook", None)
        if retain_grads and hook is not None:
            hook(self)
        dX = torch.empty_like(X)
        d.accumulate = 0
        self._call_w(pre + "dgrad", d, self._ops[1], W, L.ptr(dZ), L.ptr(W), L.ptr(dX))
        return dX


class Conv2D(_ConvBase):
    """layers/layers.py:2895-3178."""
    _layer_name = "Conv2D"
    _ops = (L.OP_CONV_FPROP, L.OP_CONV_DGRAD, L.OP_CONV_WGRAD)

    def __init__(self, out_ch, kernel_shape, pad=0, stride=1, dilation=0, act_fn=None, optimizer=None,
                 init="glorot_uniform", mode="fp32"):
        super().__init__(optimizer)
        self.pad = pad
        self.init = init
        self.in_ch = None
        self.out_ch = out_ch
        self.stride = stride
        self.dilation = dilation
        self.kernel_shape = kernel_shape
        self.act_fn = ActivationInitializer(act_fn)()
        self.parameters = {"W": None, "b": None}
        self.is_initialized = False
        self.mode = mode
        assert mode in L.MODES

    @property
    def hyperparameters(self):
        return {"layer": "Conv2D", "pad": self.pad, "init": self.init, "in_ch": self.in_ch, "out_ch": self.out_ch,
                "stride": self.stride, "dilation": self.dilation, "act_fn": str(self.act_fn),
                "kernel_shape": self.kernel_shape, "optimizer": self._opt_hparams()}

    def _geometry(self, X_shape):
        fr, fc = self.kernel_shape
        p4 = U.resolve_pad_2D(X_shape, self.pad, (fr, fc), self.stride, self.dilation)
        oh, ow = U.conv_out_hw(X_shape[1], X_shape[2], fr, fc, p4, self.stride, self.dilation)
        return p4, oh, ow, self.dilation + 1

    def _backward_naive(self, dLdy, retain_grads=True):
        """The reference keeps a second, loop-form definition of the backward pass next to the vectorised one
        (layers.py:3118-3178) as an in-repo cross-check.  Its counterpart here: the same gradients from the fp32
        CUDA-core gather kernels (direct_fp32.cu -- one thread per output element, plain loops over taps and
        channels, no tensor cores, no TMA), whatever `mode` the layer runs in.  Same contract as `backward`."""
        assert self.trainable, "Layer is frozen"
        if not isinstance(dLdy, list):
            dLdy = [dLdy]
        mode, self.mode = self.mode, "fp32"
        try:
            dXs = []
            for dy, x, z in zip(dLdy, self.X, self.derived_variables["Z"]):
                as_np = isinstance(dy, np.ndarray)
                dx = self._bwd(L.to_device(dy, torch.float32), x.float(), z.float(), retain_grads)
                dXs.append(L.to_numpy(dx) if as_np else dx.to(x.dtype))
        finally:
            self.mode = mode
        return dXs[0] if len(self.X) == 1 else dXs


class Deconv2D(_ConvBase):
    """layers/layers.py:3353-3568 (no dilation argument)."""
    _layer_name = "Deconv2D"
    _ops = (L.OP_DECONV_FPROP, L.OP_DECONV_DGRAD, L.OP_DECONV_WGRAD)

    def __init__(self, out_ch, kernel_shape, pad=0, stride=1, act_fn=None, optimizer=None, init="glorot_uniform",
                 mode="fp32"):
        super().__init__(optimizer)
        self.pad = pad
        self.init = init
        self.in_ch = None
        self.stride = stride
        self.out_ch = out_ch
        self.kernel_shape = kernel_shape
        self.act_fn = ActivationInitializer(act_fn)()
        self.parameters = {"W": None, "b": None}
        self.is_initialized = False
        self.mode = mode
        assert mode in L.MODES

    @property
    def hyperparameters(self):
        return {"layer": "Deconv2D", "pad": self.pad, "init": self.init, "in_ch": self.in_ch, "out_ch": self.out_ch,
                "stride": self.stride, "act_fn": str(self.act_fn), "kernel_shape": self.kernel_shape,
                "optimizer": self._opt_hparams()}

    def _geometry(self, X_shape):
        fr, fc = self.kernel_shape
        p4, oh, ow = U.deconv_geometry(X_shape[1], X_shape[2], fr, fc, self.pad, self.stride)
        return p4, oh, ow, 1


class Conv1D(_ConvBase):
    """layers/layers.py:2603-2892.  The reference adds a unit row dimension and runs the 2-D path
    (utils.py:668-717, layers.py:2804-2838); so does this layer, on the same kernels: X (n_ex, l_in, in_ch),
    W (kernel_width, in_ch, out_ch), b (1, 1, out_ch), pad int | (left, right) | 'same' | 'causal'."""
    _layer_name = "Conv2D"          # entry-point family
    _ops = (L.OP_CONV_FPROP, L.OP_CONV_DGRAD, L.OP_CONV_WGRAD)

    def __init__(self, out_ch, kernel_width, pad=0, stride=1, dilation=0, act_fn=None, init="glorot_uniform",
                 optimizer=None, mode="fp32"):
        super().__init__(optimizer)
        self.pad = pad
        self.init = init
        self.in_ch = None
        self.out_ch = out_ch
        self.stride = stride
        self.dilation = dilation
        self.kernel_width = kernel_width
        self.kernel_shape = (1, kernel_width)
        self.act_fn = ActivationInitializer(act_fn)()
        self.parameters = {"W": None, "b": None}
        self.is_initialized = False
        self.mode = mode
        assert mode in L.MODES

    def _init_params(self):
        init_weights = WeightInitializer(str(self.act_fn), mode=self.init)
        W = init_weights((self.kernel_width, self.in_ch, self.out_ch))
        self.parameters = {"W": L.to_device(W, torch.float32),
                           "b": torch.zeros((1, 1, self.out_ch), dtype=torch.float32, device=L.device())}
        self.gradients = {"W": torch.zeros_like(self.parameters["W"]), "b": torch.zeros_like(self.parameters["b"])}
        self.derived_variables = {"Z": [], "out_rows": [], "out_cols": []}
        self.is_initialized = True

    @property
    def hyperparameters(self):
        return {"layer": "Conv1D", "pad": self.pad, "init": self.init, "in_ch": self.in_ch, "out_ch": self.out_ch,
                "stride": self.stride, "dilation": self.dilation, "act_fn": str(self.act_fn),
                "kernel_width": self.kernel_width, "optimizer": self._opt_hparams()}

    def _geometry(self, X_shape):
        n, _, l_in, c = X_shape
        p = U.resolve_pad_1D((n, l_in, c), self.pad, self.kernel_width, self.stride, self.dilation)
        p4 = (0, 0, p[0], p[1])
        oh, ow = U.conv_out_hw(1, l_in, 1, self.kernel_width, p4, self.stride, self.dilation)
        return p4, oh, ow, self.dilation + 1

    def forward(self, X, retain_derived=True):
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, L.torch_dtype(self.mode))
        if Xd.dim() != 3:
            raise ValueError("Conv1D expects (n_ex, l_in, in_ch), got {}".format(tuple(Xd.shape)))
        Y = super().forward(Xd.unsqueeze(1), retain_derived).squeeze(1)
        if retain_derived:                     # keep what the reference keeps: 3-D X / Z, Z.shape[1:] (layers.py:2746-2750)
            dv = self.derived_variables
            self.X[-1] = self.X[-1].squeeze(1)
            dv["Z"][-1] = dv["Z"][-1].squeeze(1)
            dv["out_rows"][-1], dv["out_cols"][-1] = dv["Z"][-1].shape[1], dv["Z"][-1].shape[2]
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdy, retain_grads=True):
        assert self.trainable, "Layer is frozen"
        if not isinstance(dLdy, list):
            dLdy = [dLdy]
        dX = []
        for dy, x, z in zip(dLdy, self.X, self.derived_variables["Z"]):
            as_np = isinstance(dy, np.ndarray)
            dx = self._bwd(L.to_device(dy, L.torch_dtype(self.mode)).unsqueeze(1), x.unsqueeze(1), z.unsqueeze(1),
                           retain_grads).squeeze(1)
            dX.append(L.to_numpy(dx) if as_np else dx)
        return dX[0] if len(self.X) == 1 else dX


class Pool2D(LayerBase):
    """layers/layers.py:3181-3350: max / average pooling over (fr, fc) windows of the zero-padded NHWC input.
    The reference's four nested Python loops become one gather kernel per direction (npml_pool2d_fwd / _bwd)."""

    def __init__(self, kernel_shape, stride=1, pad=0, mode="max", optimizer=None):
        super().__init__(optimizer)
        self.pad = pad
        self.mode = mode
        self.in_ch = None
        self.out_ch = None
        self.stride = stride
        self.kernel_shape = kernel_shape
        self.is_initialized = False

    def _init_params(self):
        self.derived_variables = {"out_rows": [], "out_cols": []}
        self.is_initialized = True

    @property
    def hyperparameters(self):
        return {"layer": "Pool2D", "act_fn": None, "pad": self.pad, "mode": self.mode, "in_ch": self.in_ch,
                "out_ch": self.out_ch, "stride": self.stride, "kernel_shape": self.kernel_shape,
                "optimizer": self._opt_hparams()}

    def _desc(self, X_shape):
        if self.mode not in ("max", "average"):
            raise ValueError("Unrecognized pooling mode: {}".format(self.mode))
        n, h, w, c = X_shape
        fr, fc = self.kernel_shape
        p4 = U.resolve_pad_2D(X_shape, self.pad, (fr, fc), self.stride)
        oh, ow = U.conv_out_hw(h, w, fr, fc, p4, self.stride, 0)
        return U.make_desc(n, h, w, c, c, fr, fc, self.stride, 1, p4, oh, ow), 0 if self.mode == "max" else 1

    def forward(self, X, retain_derived=True):
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, torch.float32) if as_np else L.to_device(X, X.dtype if X.dtype != torch.float64
                                                                   else torch.float32)
        if not self.is_initialized:
            self.in_ch = self.out_ch = Xd.shape[3]
            self._init_params()
        d, pm = self._desc(tuple(Xd.shape))
        Y = torch.empty((d.N, d.OH, d.OW, d.C), dtype=Xd.dtype, device=Xd.device)
        with L.profiled("npml_pool2d_fwd"):
            L.check(L.load().npml_pool2d_fwd(d, pm, L.ptr(Xd), L.ptr(Y), L.code_of(Xd), L.stream()), "npml_pool2d_fwd")
        if retain_derived:
            self.X.append(Xd)
            self.derived_variables["out_rows"].append(d.OH)
            self.derived_variables["out_cols"].append(d.OW)
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdY, retain_grads=True):
        assert self.trainable, "Layer is frozen"
        if not isinstance(dLdY, list):
            dLdY = [dLdY]
        dXs = []
        for X, dy in zip(self.X, dLdY):
            as_np = isinstance(dy, np.ndarray)
            dyd = L.to_device(dy, X.dtype)
            d, pm = self._desc(tuple(X.shape))
            dX = torch.empty_like(X)
            with L.profiled("npml_pool2d_bwd"):
                L.check(L.load().npml_pool2d_bwd(d, pm, L.ptr(dyd), L.ptr(X), L.ptr(dX), L.code_of(X), L.stream()),
                        "npml_pool2d_bwd")
            dXs.append(L.to_numpy(dX) if as_np else dX)
        return dXs[0] if len(self.X) == 1 else dXs


class Flatten(LayerBase):
    """layers/layers.py:859-966: a reshape (no kernel; a view of the same device buffer)."""

    def __init__(self, keep_dim="first", optimizer=None):
        super().__init__(optimizer)
        self.keep_dim = keep_dim
        self._init_params()

    def _init_params(self):
        self.gradients = {}
        self.parameters = {}
        self.derived_variables = {"in_dims": []}

    @property
    def hyperparameters(self):
        return {"layer": "Flatten", "keep_dim": self.keep_dim, "optimizer": self._opt_hparams()}

    def forward(self, X, retain_derived=True):
        if retain_derived:
            self.derived_variables["in_dims"].append(tuple(X.shape))
        if self.keep_dim == -1:
            return X.reshape(1, -1)
        rs = (X.shape[0], -1) if self.keep_dim == "first" else (-1, X.shape[-1])
        return X.reshape(*rs)

    def backward(self, dLdy, retain_grads=True):
        if not isinstance(dLdy, list):
            dLdy = [dLdy]
        out = [dy.reshape(*dims) for dy, dims in zip(dLdy, self.derived_variables["in_dims"])]
        return out[0] if len(dLdy) == 1 else out


class BatchNorm2D(LayerBase):
    """layers/layers.py:969-1215.  Statistics are per channel over (n_ex, rows, cols); the variance is the
    biased one for both the normalisation and the running estimate (layers.py:1147)."""

    def __init__(self, momentum=0.9, epsilon=1e-5, optimizer=None, sync=False):
        super().__init__(optimizer)
        self.in_ch = None
        self.out_ch = None
        self.epsilon = epsilon
        self.momentum = momentum
        # sync (extension): under batch sharding the per-channel sums are all-reduced over ranks, so
        # every rank normalises with the statistics of the WHOLE batch exactly like the single-process reference
        self.sync = bool(sync)
        self.parameters = {"scaler": None, "intercept": None, "running_var": None, "running_mean": None}
        self.is_initialized = False
        self._saved = []      # (mean, rstd) of every retained forward

    def _init_params(self):
        dev = L.device()
        scaler = np.random.rand(self.in_ch)
        self.parameters = {"scaler": L.to_device(scaler, torch.float32),
                           "intercept": torch.zeros(self.in_ch, dtype=torch.float32, device=dev),
                           "running_var": torch.ones(self.in_ch, dtype=torch.float32, device=dev),
                           "running_mean": torch.zeros(self.in_ch, dtype=torch.float32, device=dev)}
        self.gradients = {"scaler": torch.zeros(self.in_ch, dtype=torch.float32, device=dev),
                          "intercept": torch.zeros(self.in_ch, dtype=torch.float32, device=dev)}
        self.is_initialized = True

    @property
    def hyperparameters(self):
        return {"layer": "BatchNorm2D", "act_fn": None, "in_ch": self.in_ch, "out_ch": self.out_ch,
                "epsilon": self.epsilon, "momentum": self.momentum, "optimizer": self._opt_hparams()}

    def folded_affine(self, in_ch=None):
        """(scale, shift) device vectors with BatchNorm2D_inference(x) = scale * x + shift (layers.py:1141-1156 on the
        running statistics), for the convolution's fused epilogue (npml_conv2d_fprop_bn)."""
        if not self.is_initialized:
            self.in_ch = self.out_ch = int(in_ch)
            self._init_params()
        P = self.parameters
        for k in P:
            P[k] = _as_param(P[k])
        ch = self.in_ch
        out = torch.empty(2, ch, dtype=torch.float32, device=L.device())
        L.check(L.load().npml_bn2d_fold(ch, L.ptr(P["scaler"]), L.ptr(P["intercept"]), L.ptr(P["running_mean"]),
                                        L.ptr(P["running_var"]), float(self.epsilon), L.ptr(out[0]), L.ptr(out[1]),
                                        L.stream()), "npml_bn2d_fold")
        return out[0], out[1]

    def reset_running_stats(self):
        assert self.trainable, "Layer is frozen"
        self.parameters["running_mean"] = torch.zeros(self.in_ch, dtype=torch.float32, device=L.device())
        self.parameters["running_var"] = torch.ones(self.in_ch, dtype=torch.float32, device=L.device())

    def flush_gradients(self, zero=True):
        super().flush_gradients(zero)
        self._saved = []

    def _ws(self, rows, ch):
        d = U.make_desc(1, 1, 1, 1, ch, 1, 1, 1, 1, (0, 0, 0, 0), 1, int(rows))
        return L.workspace(L.load().npml_workspace_bytes(d, L.OP_BN))

    def forward(self, X, retain_derived=True, sums=None):
        """layers.py:1095-1158.  `sums` (an extension): the per-channel [sum x, sum x^2] of X if the producer of X has
        already reduced them (Conv2D.forward(..., stats_out=)); the statistics pass over X is then skipped."""
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, torch.float32) if as_np else L.to_device(X, X.dtype if X.dtype != torch.float64
                                                                   else torch.float32)
        if not self.is_initialized:
            self.in_ch = self.out_ch = Xd.shape[3]
            self._init_params()
        P = self.parameters
        for k in P:
            P[k] = _as_param(P[k])
        ch = Xd.shape[3]
        rows = Xd.numel() // ch
        y = torch.empty_like(Xd)
        mean = torch.empty(ch, dtype=torch.float32, device=Xd.device)
        rstd = torch.empty(ch, dtype=torch.float32, device=Xd.device)
        use_batch = 1 if (self.trainable and retain_derived) else 0
        ws = self._ws(rows, ch)
        if use_batch and (self.sync or sums is not None):
            from . import dist as D
            lib = L.load()
            total = rows
            with L.profiled("npml_bn2d_fwd"):
                if sums is None:
                    sums = torch.empty(2 * ch, dtype=torch.float64, device=Xd.device)
                    L.check(lib.npml_bn2d_stats(rows, ch, L.ptr(Xd), L.code_of(Xd), L.ptr(sums), L.ptr(ws), ws.numel(),
                                                L.stream()), "npml_bn2d_stats")
                if self.sync:
                    D.require_equal_shards(rows)       # total = rows * world below; raises on ragged / empty shards
                    D.allreduce_sum_f64(sums)
                    total = rows * D.world_size()
                L.check(lib.npml_bn2d_apply(rows, total, ch, L.ptr(Xd), L.ptr(y), L.code_of(Xd), L.ptr(P["scaler"]),
                                            L.ptr(P["intercept"]), L.ptr(P["running_mean"]), L.ptr(P["running_var"]),
                                            float(self.momentum), float(self.epsilon), L.ptr(sums), L.ptr(mean),
                                            L.ptr(rstd), L.ptr(ws), ws.numel(), L.stream()), "npml_bn2d_apply")
            if retain_derived:
                self.X.append(Xd)
                self._saved.append((mean, rstd))
            return L.to_numpy(y) if as_np else y
        with L.profiled("npml_bn2d_fwd"):
            L.check(L.load().npml_bn2d_fwd(rows, ch, L.ptr(Xd), L.ptr(y), L.code_of(Xd), L.ptr(P["scaler"]),
                                           L.ptr(P["intercept"]), L.ptr(P["running_mean"]), L.ptr(P["running_var"]),
                                           float(self.momentum), float(self.epsilon), use_batch, L.ptr(mean),
                                           L.ptr(rstd), L.ptr(ws), ws.numel(), L.stream()), "npml_bn2d_fwd")
        if retain_derived:
            self.X.append(Xd)
            self._saved.append((mean, rstd) if use_batch else None)
        return L.to_numpy(y) if as_np else y

    # ---- pieces of forward / backward for callers that fuse the normalisation into their own pass (modules.py) ----
    def forward_stats(self, X, sums=None):
        """Training-mode statistics of X without normalising it (`sums`: already reduced by the producer of X): per-channel (mean, rstd) of the batch (all ranks
        under `sync`), the running-statistics update (layers.py:1146-1150), and X retained for the backward pass.  The
        caller normalises inside a fused pass (npml_bn2d_pair_add_act_fwd)."""
        assert self.trainable, "Layer is frozen"
        Xd = L.to_device(X, X.dtype if X.dtype != torch.float64 else torch.float32)
        if not self.is_initialized:
            self.in_ch = self.out_ch = Xd.shape[3]
            self._init_params()
        P = self.parameters
        for k in P:
            P[k] = _as_param(P[k])
        ch = Xd.shape[3]
        rows = Xd.numel() // ch
        lib = L.load()
        ws = self._ws(rows, ch)
        mean = torch.empty(ch, dtype=torch.float32, device=Xd.device)
        rstd = torch.empty(ch, dtype=torch.float32, device=Xd.device)
        total = rows
        with L.profiled("npml_bn2d_stats"):
            if sums is None:
                sums = torch.empty(2 * ch, dtype=torch.float64, device=Xd.device)
                L.check(lib.npml_bn2d_stats(rows, ch, L.ptr(Xd), L.code_of(Xd), L.ptr(sums), L.ptr(ws), ws.numel(),
                                            L.stream()), "npml_bn2d_stats")
            if self.sync:
                from . import dist as D
                D.require_equal_shards(rows)
                D.allreduce_sum_f64(sums)
                total = rows * D.world_size()
            L.check(lib.npml_bn2d_finalize(total, ch, L.ptr(sums), L.ptr(P["running_mean"]), L.ptr(P["running_var"]),
                                           float(self.momentum), float(self.epsilon), L.ptr(mean), L.ptr(rstd), L.stream()),
                    "npml_bn2d_finalize")
        self.X.append(Xd)
        self._saved.append((mean, rstd))
        return mean, rstd

    def normalise_with(self, X, saved):
        """scaler * (X - mean) * rstd + intercept for given (mean, rstd): what `forward` would have returned for a
        retained input whose normalisation was fused into another pass (used to materialise derived variables)."""
        P, (mean, rstd) = self.parameters, saved
        ch = X.shape[3]
        y = torch.empty_like(X)
        zero = torch.zeros_like(X)
        L.check(L.load().npml_bn2d_pair_add_act_fwd(X.numel() // ch, ch, L.ptr(X), L.ptr(zero), L.code_of(X),
                                                    L.ptr(P["scaler"]), L.ptr(P["intercept"]), L.ptr(mean), L.ptr(rstd),
                                                    None, None, None, None, 0, 0.0, 0.0, None, L.ptr(y), L.stream()),
                "npml_bn2d_pair_add_act_fwd")
        return y

    def backward_act(self, dLdy, act, zm=None, retain_grads=True, index=-1):
        """Backward of ONE retained forward whose dX pass also applies act'(z) of the layer in front of this one
        (x = act(z)): returns dZ = BN'(dLdy) * act'(z) (npml_bn2d_bwd_act).  zm: z, or None for ReLU / identity."""
        assert self.trainable, "Layer is frozen"
        X, saved = self.X[index], self._saved[index]
        dLdy = L.to_device(dLdy, X.dtype)
        P, G = self.parameters, self.gradients
        ch = X.shape[3]
        rows = X.numel() // ch
        mean, rstd = saved if saved is not None else self._batch_stats(X)
        dZ = torch.empty_like(X)
        ws = self._ws(rows, ch)
        lib = L.load()
        local = glob = None
        total = rows
        with L.profiled("npml_bn2d_bwd"):
            if self.sync and saved is not None:
                from . import dist as D
                local = torch.empty(2 * ch, dtype=torch.float64, device=X.device)
                L.check(lib.npml_bn2d_bwd_stats(rows, ch, L.ptr(dLdy), L.ptr(X), L.code_of(X), L.ptr(mean), L.ptr(rstd),
                                                L.ptr(local), L.ptr(ws), ws.numel(), L.stream()), "npml_bn2d_bwd_stats")
                glob = D.allreduce_sum_f64(local.clone()) if D.world_size() > 1 else local
                total = rows * D.world_size()
            L.check(lib.npml_bn2d_bwd_act(rows, total, ch, L.ptr(dLdy), L.ptr(X), L.ptr(dZ), L.code_of(X),
                                          L.ptr(_as_param(P["scaler"])), L.ptr(mean), L.ptr(rstd), L.ptr(local), L.ptr(glob),
                                          L.ptr(G["scaler"]) if retain_grads else None,
                                          L.ptr(G["intercept"]) if retain_grads else None, 1, act.code, act.p0, act.p1,
                                          L.ptr(zm), L.ptr(ws), ws.numel(), L.stream()), "npml_bn2d_bwd_act")
        return dZ

    def backward(self, dLdy, retain_grads=True):
        assert self.trainable, "Layer is frozen"
        if not isinstance(dLdy, list):
            dLdy = [dLdy]
        dX = []
        for i, (dy, x) in enumerate(zip(dLdy, self.X)):
            as_np = isinstance(dy, np.ndarray)
            dx = self._bwd(L.to_device(dy, x.dtype), x, self._saved[i] if i < len(self._saved) else None,
                           retain_grads)
            dX.append(L.to_numpy(dx) if as_np else dx)
        return dX[0] if len(self.X) == 1 else dX

    def _batch_stats(self, X):
        """(mean, rstd) of X; the reference's _bwd always recomputes them from X (layers.py:1207)."""
        ch = X.shape[3]
        rows = X.numel() // ch
        mean = torch.empty(ch, dtype=torch.float32, device=X.device)
        rstd = torch.empty(ch, dtype=torch.float32, device=X.device)
        scratch_rm, scratch_rv = torch.zeros_like(mean), torch.ones_like(mean)
        ws = self._ws(rows, ch)
        tmp = torch.empty_like(X)
        L.check(L.load().npml_bn2d_fwd(rows, ch, L.ptr(X), L.ptr(tmp), L.code_of(X), L.ptr(self.parameters["scaler"]),
                                       L.ptr(self.parameters["intercept"]), L.ptr(scratch_rm), L.ptr(scratch_rv),
                                       float(self.momentum), float(self.epsilon), 1, L.ptr(mean), L.ptr(rstd),
                                       L.ptr(ws), ws.numel(), L.stream()), "npml_bn2d_fwd")
        return mean, rstd

    def _bwd(self, dLdy, X, saved, retain_grads=True):
        P = self.parameters
        ch = X.shape[3]
        rows = X.numel() // ch
        mean, rstd = saved if saved is not None else self._batch_stats(X)
        dX = torch.empty_like(X)
        ws = self._ws(rows, ch)
        G = self.gradients
        if self.sync and saved is not None:
            from . import dist as D
            lib = L.load()
            local = torch.empty(2 * ch, dtype=torch.float64, device=X.device)
            with L.profiled("npml_bn2d_bwd"):
                L.check(lib.npml_bn2d_bwd_stats(rows, ch, L.ptr(dLdy), L.ptr(X), L.code_of(X), L.ptr(mean), L.ptr(rstd),
                                                L.ptr(local), L.ptr(ws), ws.numel(), L.stream()), "npml_bn2d_bwd_stats")
                glob = D.allreduce_sum_f64(local.clone()) if D.world_size() > 1 else local
                L.check(lib.npml_bn2d_bwd_apply(rows, rows * D.world_size(), ch, L.ptr(dLdy), L.ptr(X), L.ptr(dX),
                                                L.code_of(X), L.ptr(_as_param(P["scaler"])), L.ptr(mean), L.ptr(rstd),
                                                L.ptr(local), L.ptr(glob),
                                                L.ptr(G["scaler"]) if retain_grads else None,
                                                L.ptr(G["intercept"]) if retain_grads else None, 1, L.ptr(ws),
                                                ws.numel(), L.stream()), "npml_bn2d_bwd_apply")
            return dX
        with L.profiled("npml_bn2d_bwd"):
            L.check(L.load().npml_bn2d_bwd(rows, ch, L.ptr(dLdy), L.ptr(X), L.ptr(dX), L.code_of(X),
                                           L.ptr(_as_param(P["scaler"])), L.ptr(mean), L.ptr(rstd),
                                           L.ptr(G["scaler"]) if retain_grads else None,
                                           L.ptr(G["intercept"]) if retain_grads else None, 1, L.ptr(ws), ws.numel(),
                                           L.stream()), "npml_bn2d_bwd")
        return dX


class Add(LayerBase):
    """Element-wise sum of several inputs followed by an activation (layers/layers.py:633-742)."""

    def __init__(self, act_fn=None, optimizer=None):
        super().__init__(optimizer)
        self.act_fn = ActivationInitializer(act_fn)()
        self._init_params()

    def _init_params(self):
        self.gradients = {}
        self.parameters = {}
        self.derived_variables = {"sum": []}

    @property
    def hyperparameters(self):
        return {"layer": "Sum", "act_fn": str(self.act_fn), "optimizer": self._opt_hparams()}

    def forward(self, X, retain_derived=True):
        as_np = isinstance(X[0], np.ndarray)
        first = X[0]
        dt = torch.float32 if (as_np or first.dtype == torch.float64) else first.dtype
        Xd = [L.to_device(x, dt) for x in X]
        _same_shapes(Xd, "Add")
        n = Xd[0].numel()
        total = torch.empty_like(Xd[0])
        Y = total if self.act_fn.code == 0 else torch.empty_like(total)
        arr = (C.c_void_p * len(Xd))(*[x.data_ptr() for x in Xd])
        a = self.act_fn
        with L.profiled("npml_add_act_fwd"):
            L.check(L.load().npml_add_act_fwd(len(Xd), arr, L.code_of(total), n, a.code, a.p0, a.p1, L.ptr(total),
                                              None if Y is total else L.ptr(Y), L.stream()), "npml_add_act_fwd")
        if retain_derived:
            self.X.append(Xd)
            self.derived_variables["sum"].append(total)
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdY, retain_grads=True):
        if not isinstance(dLdY, list):
            dLdY = [dLdY]
        grads = []
        for dy, x, ss in zip(dLdY, self.X, self.derived_variables["sum"]):
            as_np = isinstance(dy, np.ndarray)
            g = self._bwd(L.to_device(dy, ss.dtype), x, ss)
            grads.append([L.to_numpy(v) for v in g] if as_np else g)
        return grads[0] if len(self.X) == 1 else grads

    def _bwd(self, dLdY, X, _sum):
        """Every input receives dLdY * act'(sum) (layers.py:739-742).  The reference returns one NEW array per input;
        so does this layer (a caller's in-place `dX +=` must not reach the sibling gradients or its own dLdY), unless
        the owner promises to only read them (`_share_grads`, set by the skip-connection modules)."""
        g = _dz_prep(dLdY, _sum, self.act_fn, _sum.dtype)
        if getattr(self, "_share_grads", False):
            return [g for _ in X]
        first = g.clone() if g is dLdY else g
        return [first] + [g.clone() for _ in X[1:]]


class Multiply(LayerBase):
    """Element-wise product of several inputs followed by an activation (layers/layers.py:745-856)."""

    def __init__(self, act_fn=None, optimizer=None):
        super().__init__(optimizer)
        self.act_fn = ActivationInitializer(act_fn)()
        self._init_params()

    def _init_params(self):
        self.gradients = {}
        self.parameters = {}
        self.derived_variables = {"product": []}

    @property
    def hyperparameters(self):
        return {"layer": "Multiply", "act_fn": str(self.act_fn), "optimizer": self._opt_hparams()}

    def forward(self, X, retain_derived=True):
        as_np = isinstance(X[0], np.ndarray)
        first = X[0]
        dt = torch.float32 if (as_np or first.dtype == torch.float64) else first.dtype
        Xd = [L.to_device(x, dt) for x in X]
        _same_shapes(Xd, "Multiply")
        prod = torch.empty_like(Xd[0])
        Y = prod if self.act_fn.code == 0 else torch.empty_like(prod)
        arr = (C.c_void_p * len(Xd))(*[x.data_ptr() for x in Xd])
        a = self.act_fn
        with L.profiled("npml_mul_act_fwd"):
            L.check(L.load().npml_mul_act_fwd(len(Xd), arr, L.code_of(prod), prod.numel(), a.code, a.p0, a.p1,
                                              L.ptr(prod), None if Y is prod else L.ptr(Y), L.stream()),
                    "npml_mul_act_fwd")
        if retain_derived:
            self.X.append(Xd)
            self.derived_variables["product"].append(prod)
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdY, retain_grads=True):
        if not isinstance(dLdY, list):
            dLdY = [dLdY]
        grads = []
        for dy, x, pr in zip(dLdY, self.X, self.derived_variables["product"]):
            as_np = isinstance(dy, np.ndarray)
            g = self._bwd(L.to_device(dy, pr.dtype), x, pr)
            grads.append([L.to_numpy(v) for v in g] if as_np else g)
        return grads[0] if len(self.X) == 1 else grads

    def _bwd(self, dLdY, X, prod):
        out = [torch.empty_like(prod) for _ in X]
        xs = (C.c_void_p * len(X))(*[x.data_ptr() for x in X])
        gs = (C.c_void_p * len(X))(*[g.data_ptr() for g in out])
        a = self.act_fn
        with L.profiled("npml_mul_act_bwd"):
            L.check(L.load().npml_mul_act_bwd(len(X), xs, gs, L.code_of(prod), prod.numel(), a.code, a.p0, a.p1,
                                              L.ptr(dLdY), L.ptr(prod), L.stream()), "npml_mul_act_bwd")
        return out


class FullyConnected(_ConvBase):
    """layers/layers.py:2010-2170: Y = act_fn(X @ W + b) with X (n_ex, n_in), W (n_in, n_out), b (1, n_out).
    A dense layer is a 1x1 convolution over a 1x1 image, so it runs on the Conv2D entry points unchanged
    (dX = dZ @ W.T -> dgrad, dW = X.T @ dZ and db -> wgrad_db)."""
    _layer_name = "Conv2D"
    _ops = (L.OP_CONV_FPROP, L.OP_CONV_DGRAD, L.OP_CONV_WGRAD)

    def __init__(self, n_out, act_fn=None, init="glorot_uniform", optimizer=None, mode="fp32"):
        super().__init__(optimizer)
        self.init = init
        self.n_in = None
        self.n_out = n_out
        self.out_ch = n_out
        self.pad, self.stride, self.dilation, self.kernel_shape = 0, 1, 0, (1, 1)
        self.act_fn = ActivationInitializer(act_fn)()
        self.parameters = {"W": None, "b": None}
        self.is_initialized = False
        self.mode = mode
        assert mode in L.MODES

    def _init_params(self):
        init_weights = WeightInitializer(str(self.act_fn), mode=self.init)
        W = init_weights((self.n_in, self.n_out))
        self.parameters = {"W": L.to_device(W, torch.float32),
                           "b": torch.zeros((1, self.n_out), dtype=torch.float32, device=L.device())}
        self.gradients = {"W": torch.zeros_like(self.parameters["W"]), "b": torch.zeros_like(self.parameters["b"])}
        self.derived_variables = {"Z": []}
        self.is_initialized = True

    @property
    def hyperparameters(self):
        return {"layer": "FullyConnected", "init": self.init, "n_in": self.n_in, "n_out": self.n_out,
                "act_fn": str(self.act_fn), "optimizer": self._opt_hparams()}

    def _geometry(self, X_shape):
        return (0, 0, 0, 0), 1, 1, 1

    def forward(self, X, retain_derived=True):
        as_np = isinstance(X, np.ndarray)
        Xd = L.to_device(X, L.torch_dtype(self.mode))
        if Xd.dim() != 2:
            raise ValueError("FullyConnected expects (n_ex, n_in), got {}".format(tuple(Xd.shape)))
        if not self.is_initialized:
            self.n_in = self.in_ch = Xd.shape[1]
            self._init_params()
        n, d = Xd.shape
        if retain_derived:
            self.derived_variables.setdefault("out_rows", [])
            self.derived_variables.setdefault("out_cols", [])
        Y = super().forward(Xd.view(n, 1, 1, d), retain_derived).view(n, self.n_out)
        if retain_derived:
            dv = self.derived_variables
            self.X[-1] = self.X[-1].view(n, d)
            dv["Z"][-1] = dv["Z"][-1].view(n, self.n_out)
            dv.pop("out_rows"), dv.pop("out_cols")          # the reference keeps only Z (layers.py:2056)
        return L.to_numpy(Y) if as_np else Y

    def backward(self, dLdy, retain_grads=True):
        assert self.trainable, "Layer is frozen"
        if not isinstance(dLdy, list):
            dLdy = [dLdy]
        dX = []
        for dy, x, z in zip(dLdy, self.X, self.derived_variables["Z"]):
            as_np = isinstance(dy, np.ndarray)
            n, d = x.shape
            dx = self._bwd(L.to_device(dy, L.torch_dtype(self.mode)).view(n, 1, 1, self.n_out), x.view(n, 1, 1, d),
                           z.view(n, 1, 1, self.n_out), retain_grads).view(n, d)
            dX.append(L.to_numpy(dx) if as_np else dx)
        return dX[0] if len(self.X) == 1 else dX
