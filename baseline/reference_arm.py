"""The reference arm of bench.py: the UNMODIFIED ddbourgin/numpy-ml (staged by baseline/stage_ref.sh under the
git-ignored baseline/_ref/) driven through its own public layer API -- Conv2D / Deconv2D / SkipConnectionConvModule
`forward` + `backward` in float64 -- on the objects and synthetic inputs of SURVEY.md 8d.  Nothing of this repository's
package, kernels or oracle is on that path.  Only bench.py imports this module."""
import collections
import collections.abc
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "numpy_ml", "neural_nets"))


def load():
    """Import numpy_ml from baseline/_ref.  Two shims, neither touches the reference's files: `collections.Hashable`
    (removed in Python 3.10; numpy_ml/utils/data_structures.py:3 still imports it) and nothing else."""
    collections.Hashable = collections.abc.Hashable
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import numpy_ml.neural_nets.layers as layers
    import numpy_ml.neural_nets.modules as modules
    import numpy_ml.neural_nets.activations as activations
    return types.SimpleNamespace(layers=layers, modules=modules, activations=activations)


def build(ref, cfg, seed):
    """The reference object of a BASELINE config, weights from the reference's own initialiser."""
    np.random.seed(seed)
    if cfg["kind"] == "conv":
        return ref.layers.Conv2D(out_ch=cfg["K"], kernel_shape=cfg["k"], pad=cfg["pad"], stride=cfg["stride"],
                                 dilation=cfg["dilation"])
    if cfg["kind"] == "deconv":
        return ref.layers.Deconv2D(out_ch=cfg["K"], kernel_shape=cfg["k"], pad=cfg["pad"], stride=cfg["stride"])
    return ref.modules.SkipConnectionConvModule(out_ch1=cfg["K"], out_ch2=cfg["K"], kernel_shape1=(3, 3),
                                                kernel_shape2=(3, 3), kernel_shape_skip=(1, 1), pad1=1, pad2=1,
                                                stride1=1, stride2=1, stride_skip=1, act_fn=ref.activations.ReLU())


def flush(obj):
    """Drop the retained inputs and gradients between repetitions without an optimizer step (per
    component for the module -- never module.update())."""
    comps = getattr(obj, "components", None)
    if comps is None:
        obj.flush_gradients()
        return
    for c in comps:
        if hasattr(c, "flush_gradients"):
            c.flush_gradients()
    if hasattr(obj, "_dv"):
        obj._dv = {}
    obj.X = []


def step(obj, X, dY):
    """One forward + backward (dX, dW, db) in float64 through the reference's public API."""
    Y = obj.forward(X)
    dX = obj.backward(dY)
    flush(obj)
    return Y, dX
