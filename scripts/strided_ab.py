"""A/B of the slab kernel's stride > 1 paths against the per-tile gather kernel (same process cannot switch: the plan
reads the environment on every call, so the variable is flipped between timed loops).

    python scripts/strided_ab.py
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("numpy-ml_b200")


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def case(name, layer, xs):
    X = torch.randn(xs, device="cuda").to(torch.bfloat16)
    Y = layer.forward(X)
    dY = torch.randn(Y.shape, device="cuda").to(torch.bfloat16)
    out = []
    for var in (None, "NPML_DISABLE_SLAB_PLANES", "NPML_DISABLE_SLAB_CLASSES"):
        for k in ("NPML_DISABLE_SLAB_PLANES", "NPML_DISABLE_SLAB_CLASSES"):
            os.environ.pop(k, None)
        if var:
            os.environ[var] = "1"
        layer.flush_gradients()
        f = timeit(lambda: layer.forward(X, retain_derived=False))
        layer.forward(X)

        def bwd():
            layer.backward(dY, retain_grads=False)
        b = timeit(bwd)
        out.append("{}: fwd {:.1f} us, dX {:.1f} us".format(var or "slab", f, b))
    print(name, "|", " | ".join(out), flush=True)


case("Conv2D 3x3 s2 p1 256x56x56x64->128", nb.Conv2D(128, (3, 3), pad=1, stride=2, mode="bf16"), (256, 56, 56, 64))
case("Conv2D 3x3 s2 p1 128x28x28x128->256", nb.Conv2D(256, (3, 3), pad=1, stride=2, mode="bf16"), (128, 28, 28, 128))
case("Deconv2D 4x4 s2 p1 128x16x16x256->128 (cfg5)", nb.Deconv2D(128, (4, 4), pad=1, stride=2, mode="bf16"), (128, 16, 16, 256))
