"""SGD with momentum and clip-norm on device parameters (numpy_ml/neural_nets/optimizers/optimizers.py:58-148).

The step runs on the flat fp32 parameter / gradient tensors after the gradient all-reduce
(SURVEY.md 8f rank 1).  Only SGD with a constant learning rate is provided in this round; the other
reference optimizers (AdaGrad / RMSProp / Adam) raise NotImplementedError."""
import math
import re

import torch

from . import _lib as L


class OptimizerBase(object):
    def __init__(self, lr):
        self.cache = {}
        self.cur_step = 0
        self.hyperparameters = {}
        self.lr = lr

    def __call__(self, param, param_grad, param_name, cur_loss=None):
        return self.update(param, param_grad, param_name, cur_loss)

    def step(self):
        self.cur_step += 1

    def reset_step(self):
        self.cur_step = 0


class SGD(OptimizerBase):
    def __init__(self, lr=0.01, momentum=0.0, clip_norm=None, lr_scheduler=None, **kwargs):
        super().__init__(lr)
        if lr_scheduler not in (None, "constant") and not str(lr_scheduler).lower().startswith("constant"):
            raise NotImplementedError("only the constant learning-rate schedule is provided")
        self.hyperparameters = {"id": "SGD", "lr": lr, "momentum": momentum, "clip_norm": clip_norm,
                                "lr_scheduler": "ConstantScheduler(lr={})".format(lr)}

    def __str__(self):
        H = self.hyperparameters
        return "SGD(lr={}, momentum={}, clip_norm={}, lr_scheduler={})".format(H["lr"], H["momentum"], H["clip_norm"],
                                                                             H["lr_scheduler"])

    def update(self, param, param_grad, param_name, cur_loss=None):
        """u = momentum*cache + lr*g (g rescaled to `clip_norm` if longer); cache = u; param -= u
        (optimizers.py:110-148).  Updates `param` in place and returns it."""
        H = self.hyperparameters
        lib = L.load()
        if param_name not in self.cache:
            self.cache[param_name] = torch.zeros_like(param_grad)
        scale = 1.0
        if H["clip_norm"] is not None:
            out = torch.empty(1, dtype=torch.float64, device=param_grad.device)
            ws = L.workspace(1 << 20)
            L.check(lib.npml_sumsq_f32(L.ptr(param_grad), param_grad.numel(), L.ptr(out), L.ptr(ws), ws.numel(),
                                       L.stream()), "npml_sumsq_f32")
            nrm = math.sqrt(float(out.item()))
            if nrm > H["clip_norm"]:
                scale = H["clip_norm"] / nrm
        L.check(lib.npml_sgd_update(L.ptr(param), L.ptr(param_grad), L.ptr(self.cache[param_name]), param.numel(),
                                    float(H["lr"]), float(H["momentum"]), float(scale), L.stream()), "npml_sgd_update")
        return param


def _eval(v):
    try:
        return float(v) if v not in ("none", "None") else None
    except ValueError:
        return v


class OptimizerInitializer(object):
    """None | OptimizerBase | "SGD(lr=0.01, ...)" | {"hyperparameters": ..., "cache": ...}
    (initializers/initializers.py:176-239)."""

    def __init__(self, param=None):
        self.param = param

    def __call__(self):
        p = self.param
        if p is None:
            return SGD()
        if isinstance(p, OptimizerBase):
            return p
        if isinstance(p, str):
            s = p.lower()
            kwargs = {i: _eval(j) for i, j in re.findall(r"([a-zA-Z_]*)=([^,)]*)", s)}
            if "sgd" in s:
                kwargs.pop("lr_scheduler", None)
                return SGD(**kwargs)
            raise NotImplementedError("{}".format(s))
        if isinstance(p, dict):
            op = p.get("hyperparameters")
            if op is None:
                raise ValueError("`param` dictionary has no `hyperparemeters` key")
            if op["id"] != "SGD":
                raise NotImplementedError("{}".format(op["id"]))
            opt = SGD(lr=op.get("lr", 0.01), momentum=op.get("momentum", 0.0), clip_norm=op.get("clip_norm"))
            if p.get("cache"):
                opt.cache = dict(p["cache"])
            return opt
        raise ValueError("Unknown optimizer: {}".format(p))
