"""Activation objects with the reference's names, string forms and kink conventions
(numpy_ml/neural_nets/activations/activations.py).  On the hot path they are not called
element by element: a layer passes `code/p0/p1` to the CUDA epilogue (`fn`) or to the
dZ = dLdy * act'(Z) prologue (`grad`).  `fn`/`grad` below run the same device functors stand-alone."""
import re

import numpy as np
import torch

from . import _lib as L


class ActivationBase(object):
    code, p0, p1 = 0, 0.0, 0.0

    def __call__(self, z):
        return self.fn(z)

    def fn(self, z):
        """act(z) on the device; NumPy in -> float64 NumPy out, tensor in -> tensor out."""
        as_np = isinstance(z, np.ndarray)
        zd = L.to_device(z, torch.float32 if as_np else (z.dtype if z.dtype != torch.float64 else torch.float32))
        y = torch.empty_like(zd)
        L.check(L.load().npml_act_fwd(L.ptr(zd), L.ptr(y), L.code_of(zd), zd.numel(), self.code, self.p0, self.p1,
                                      L.stream()), "npml_act_fwd")
        return L.to_numpy(y) if as_np else y

    def grad(self, x):
        """act'(x), with the reference's conventions at the kinks."""
        as_np = isinstance(x, np.ndarray)
        xd = L.to_device(x, torch.float32 if as_np else (x.dtype if x.dtype != torch.float64 else torch.float32))
        ones = torch.ones_like(xd)
        out = torch.empty_like(xd)
        c = L.code_of(xd)
        L.check(L.load().npml_dz_prep(xd.numel(), 1, self.code, self.p0, self.p1, L.ptr(ones), c, L.ptr(xd), c,
                                      L.ptr(out), c, None, 0, None, 0, L.stream()), "npml_dz_prep")
        return L.to_numpy(out) if as_np else out


class Identity(ActivationBase):                     # activations.py:384-409 (Affine(1, 0))
    code = 0

    def __str__(self):
        return "Identity"


class Affine(ActivationBase):                       # activations.py:342-381
    code = 1

    def __init__(self, slope=1, intercept=0):
        self.slope, self.intercept = slope, intercept
        self.p0, self.p1 = float(slope), float(intercept)
        if self.p0 == 1.0 and self.p1 == 0.0:
            self.code = 0     # Affine(1, 0) is the identity: the kernels skip the epilogue / write Y once

    def __str__(self):
        return "Affine(slope={}, intercept={})".format(self.slope, self.intercept)


class ReLU(ActivationBase):                         # activations.py:73-137
    code = 2

    def __str__(self):
        return "ReLU"


class LeakyReLU(ActivationBase):                    # activations.py:140-207
    code = 3

    def __init__(self, alpha=0.3):
        self.alpha = alpha
        self.p0 = float(alpha)

    def __str__(self):
        return "Leaky ReLU(alpha={})".format(self.alpha)


class Tanh(ActivationBase):                         # activations.py:304-339
    code = 4

    def __str__(self):
        return "Tanh"


class Sigmoid(ActivationBase):                      # activations.py:30-70
    code = 5

    def __str__(self):
        return "Sigmoid"


class ELU(ActivationBase):                          # activations.py:412-488
    code = 6

    def __init__(self, alpha=1.0):
        self.alpha = alpha
        self.p0 = float(alpha)

    def __str__(self):
        return "ELU(alpha={})".format(self.alpha)


class SELU(ActivationBase):                         # activations.py:532-612
    code = 7

    def __str__(self):
        return "SELU"


class HardSigmoid(ActivationBase):                  # activations.py:615-666
    code = 8

    def __str__(self):
        return "Hard Sigmoid"


class SoftPlus(ActivationBase):                     # activations.py:669-721
    code = 9

    def __str__(self):
        return "SoftPlus"


class GELU(ActivationBase):                         # activations.py:210-301 (always the tanh form, :232)
    code = 10

    def __init__(self, approximate=True):
        self.approximate = True

    def __str__(self):
        return "GELU(approximate={})".format(self.approximate)


class Exponential(ActivationBase):                  # activations.py:491-529
    code = 11

    def __str__(self):
        return "Exponential"


class ActivationInitializer(object):
    """None | ActivationBase | str -> activation object (initializers/initializers.py:41-103).
    Unlike the reference, "gelu(...)" and "elu(alpha=...)" strings parse (the reference passes a
    wrong kwarg / an undefined name there, initializers.py:96-100)."""

    def __init__(self, param=None):
        self.param = param

    def __call__(self):
        p = self.param
        if p is None:
            return Identity()
        if isinstance(p, ActivationBase):
            return p
        if isinstance(p, str):
            return self.init_from_str(p)
        raise ValueError("Unknown activation: {}".format(p))

    def init_from_str(self, act_str):
        s = act_str.lower()
        simple = {"relu": ReLU, "tanh": Tanh, "selu": SELU, "sigmoid": Sigmoid, "identity": Identity,
                  "hardsigmoid": HardSigmoid, "hard sigmoid": HardSigmoid, "softplus": SoftPlus,
                  "exponential": Exponential}
        if s in simple:
            return simple[s]()
        if "affine" in s:
            slope, intercept = re.match(r"affine\(slope=(.*), intercept=(.*)\)", s).groups()
            return Affine(float(slope), float(intercept))
        if "leaky relu" in s:
            return LeakyReLU(float(re.match(r"leaky relu\(alpha=(.*)\)", s).groups()[0]))
        if "gelu" in s:
            return GELU()
        if "elu" in s:
            return ELU(alpha=float(re.match(r"elu\(alpha=(.*)\)", s).groups()[0]))
        raise ValueError("Unknown activation: {}".format(s))
