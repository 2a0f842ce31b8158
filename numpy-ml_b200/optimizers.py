"""SGD, AdaGrad, RMSProp and Adam on device parameters, with the reference's constructor arguments, string forms,
`hyperparameters` / `cache` dicts and update rules (numpy_ml/neural_nets/optimizers/optimizers.py:58-482).

The step runs in place on the fp32 parameter / gradient tensors after the gradient all-reduce:
one elementwise kernel (`npml_optimizer_update`) per parameter; clip-norm needs one deterministic sum-of-squares
reduction first (the reference clips each parameter tensor by its own L2 norm, after the all-reduce)."""
import math
import re
from copy import deepcopy

import torch

from . import _lib as L
from .schedulers import SchedulerInitializer, _eval

KIND = {"SGD": 0, "AdaGrad": 1, "RMSProp": 2, "Adam": 3}


class OptimizerBase(object):
    """optimizers.py:8-56."""
    _id = None

    def __init__(self, lr, scheduler=None):
        self.cache = {}
        self.cur_step = 0
        self.hyperparameters = {}
        self.lr_scheduler = SchedulerInitializer(scheduler, lr=lr)()

    def __call__(self, param, param_grad, param_name, cur_loss=None):
        return self.update(param, param_grad, param_name, cur_loss)

    def step(self):
        self.cur_step += 1

    def reset_step(self):
        self.cur_step = 0

    def copy(self):
        return deepcopy(self)

    def set_params(self, hparam_dict=None, cache_dict=None):
        if hparam_dict is not None:
            for k, v in hparam_dict.items():
                if k in self.hyperparameters:
                    self.hyperparameters[k] = v
                    if k == "lr_scheduler":
                        self.lr_scheduler = SchedulerInitializer(v, lr=None)()
        if cache_dict is not None:
            for k, v in cache_dict.items():
                self.cache[k] = v

    # ---- shared by the concrete optimizers -------------------------------------------------------------
    def _launch(self, param, grad, s0, s1, lr, a=0.0, b=0.0, eps=0.0, c1=1.0, c2=1.0):
        """One elementwise kernel.  With clip_norm set, ||grad||^2 is reduced on the device and the factor
        t / ||g|| (if ||g|| > t, optimizers.py:136-139) is formed inside the update kernel: no host read-back."""
        if param.dtype != torch.float32 or grad.dtype != torch.float32 or not param.is_cuda:
            raise TypeError("optimizers update fp32 device tensors in place")
        lib, t = L.load(), self.hyperparameters["clip_norm"]
        if t is None:
            L.check(lib.npml_optimizer_update(KIND[self._id], L.ptr(param), L.ptr(grad), L.ptr(s0), L.ptr(s1),
                                              param.numel(), float(lr), float(a), float(b), float(eps), float(c1),
                                              float(c2), 1.0, L.stream()), "npml_optimizer_update")
            return param
        ssq = torch.empty(1, dtype=torch.float64, device=grad.device)
        ws = L.workspace(1 << 20)
        L.check(lib.npml_sumsq_f32(L.ptr(grad), grad.numel(), L.ptr(ssq), L.ptr(ws), ws.numel(), L.stream()),
                "npml_sumsq_f32")
        L.check(lib.npml_optimizer_update_clipped(KIND[self._id], L.ptr(param), L.ptr(grad), L.ptr(s0), L.ptr(s1),
                                                  param.numel(), float(lr), float(a), float(b), float(eps), float(c1),
                                                  float(c2), L.ptr(ssq), float(t), L.stream()),
                "npml_optimizer_update_clipped")
        return param


class SGD(OptimizerBase):
    """update = momentum * cache + lr * grad; cache = update; param -= update (optimizers.py:58-148)."""
    _id = "SGD"

    def __init__(self, lr=0.01, momentum=0.0, clip_norm=None, lr_scheduler=None, **kwargs):
        super().__init__(lr, lr_scheduler)
        self.hyperparameters = {"id": "SGD", "lr": lr, "momentum": momentum, "clip_norm": clip_norm,
                                "lr_scheduler": str(self.lr_scheduler)}

    def __str__(self):
        H = self.hyperparameters
        return "SGD(lr={}, momentum={}, clip_norm={}, lr_scheduler={})".format(H["lr"], H["momentum"], H["clip_norm"],
                                                                             H["lr_scheduler"])

    def update(self, param, param_grad, param_name, cur_loss=None):
        H = self.hyperparameters
        lr = self.lr_scheduler(self.cur_step, cur_loss)
        if param_name not in self.cache:
            self.cache[param_name] = torch.zeros_like(param_grad)
        return self._launch(param, param_grad, self.cache[param_name], None, lr, a=H["momentum"])


class AdaGrad(OptimizerBase):
    """cache += grad**2; param -= lr * grad / (sqrt(cache) + eps) (optimizers.py:156-259)."""
    _id = "AdaGrad"

    def __init__(self, lr=0.01, eps=1e-7, clip_norm=None, lr_scheduler=None, **kwargs):
        super().__init__(lr, lr_scheduler)
        self.hyperparameters = {"id": "AdaGrad", "lr": lr, "eps": eps, "clip_norm": clip_norm,
                                "lr_scheduler": str(self.lr_scheduler)}

    def __str__(self):
        H = self.hyperparameters
        return "AdaGrad(lr={}, eps={}, clip_norm={}, lr_scheduler={})".format(H["lr"], H["eps"], H["clip_norm"],
                                                                             H["lr_scheduler"])

    def update(self, param, param_grad, param_name, cur_loss=None):
        H = self.hyperparameters
        lr = self.lr_scheduler(self.cur_step, cur_loss)
        if param_name not in self.cache:
            self.cache[param_name] = torch.zeros_like(param_grad)
        return self._launch(param, param_grad, self.cache[param_name], None, lr, eps=H["eps"])


class RMSProp(OptimizerBase):
    """cache = decay*cache + (1-decay)*grad**2; param -= lr * grad / (sqrt(cache) + eps) (optimizers.py:262-361)."""
    _id = "RMSProp"

    def __init__(self, lr=0.001, decay=0.9, eps=1e-7, clip_norm=None, lr_scheduler=None, **kwargs):
        super().__init__(lr, lr_scheduler)
        self.hyperparameters = {"id": "RMSProp", "lr": lr, "eps": eps, "decay": decay, "clip_norm": clip_norm,
                                "lr_scheduler": str(self.lr_scheduler)}

    def __str__(self):
        H = self.hyperparameters
        return "RMSProp(lr={}, eps={}, decay={}, clip_norm={}, lr_scheduler={})".format(
            H["lr"], H["eps"], H["decay"], H["clip_norm"], H["lr_scheduler"])

    def update(self, param, param_grad, param_name, cur_loss=None):
        H = self.hyperparameters
        lr = self.lr_scheduler(self.cur_step, cur_loss)
        if param_name not in self.cache:
            self.cache[param_name] = torch.zeros_like(param_grad)
        return self._launch(param, param_grad, self.cache[param_name], None, lr, a=H["decay"], eps=H["eps"])


class Adam(OptimizerBase):
    """Exponential moving averages of grad and grad**2 with bias correction (optimizers.py:364-482); the per-parameter
    cache is {"t", "mean", "var"} like the reference's."""
    _id = "Adam"

    def __init__(self, lr=0.001, decay1=0.9, decay2=0.999, eps=1e-7, clip_norm=None, lr_scheduler=None, **kwargs):
        super().__init__(lr, lr_scheduler)
        self.hyperparameters = {"id": "Adam", "lr": lr, "eps": eps, "decay1": decay1, "decay2": decay2,
                                "clip_norm": clip_norm, "lr_scheduler": str(self.lr_scheduler)}

    def __str__(self):
        H = self.hyperparameters
        return "Adam(lr={}, decay1={}, decay2={}, eps={}, clip_norm={}, lr_scheduler={})".format(
            H["lr"], H["decay1"], H["decay2"], H["eps"], H["clip_norm"], H["lr_scheduler"])

    def update(self, param, param_grad, param_name, cur_loss=None):
        H = self.hyperparameters
        lr = self.lr_scheduler(self.cur_step, cur_loss)
        if param_name not in self.cache:
            self.cache[param_name] = {"t": 0, "mean": torch.zeros_like(param_grad), "var": torch.zeros_like(param_grad)}
        C = self.cache[param_name]
        C["t"] += 1
        d1, d2 = H["decay1"], H["decay2"]
        return self._launch(param, param_grad, C["mean"], C["var"], lr, a=d1, b=d2, eps=H["eps"],
                            c1=1.0 - d1 ** C["t"], c2=1.0 - d2 ** C["t"])


_BY_NAME = {"sgd": SGD, "adagrad": AdaGrad, "rmsprop": RMSProp, "adam": Adam}


class OptimizerInitializer(object):
    """None | OptimizerBase | "Adam(lr=0.001, ...)" | {"hyperparameters": ..., "cache": ...}
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
            # the scheduler's own "(...)" holds key=value pairs too: split it off first
            sched = None
            m = re.search(r"lr_scheduler=([a-z]+scheduler\([^)]*\)|[^,)]*)", s)
            if m:
                sched = m.group(1)
                s_wo = s[:m.start()] + s[m.end():]
            else:
                s_wo = s
            kwargs = {k: _eval(v) for k, v in re.findall(r"([a-zA-Z_0-9]*)=([^,)]*)", s_wo)}
            if sched not in (None, "none"):
                kwargs["lr_scheduler"] = sched
            for name in ("adagrad", "rmsprop", "adam", "sgd"):
                if name in s:
                    return _BY_NAME[name](**kwargs)
            raise NotImplementedError("{}".format(s))
        if isinstance(p, dict):
            op = p.get("hyperparameters")
            if op is None:
                raise ValueError("`param` dictionary has no `hyperparemeters` key")
            if op["id"].lower() not in _BY_NAME:
                raise NotImplementedError("{}".format(op["id"]))
            opt = _BY_NAME[op["id"].lower()]()
            opt.set_params(op, p.get("cache"))
            return opt
        raise ValueError("Unknown optimizer: {}".format(p))
