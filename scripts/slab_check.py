#!/usr/bin/env python
"""Debug aid (GPU box): run Conv2D forward/backward through the slab kernel for a few shapes and print the
normwise error against a float64 torch convolution (quick look; the parity tests proper are tests/)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nb = importlib.import_module("numpy-ml_b200")

SHAPES = [
    # N, H, W, C, K, k, pad, dilation
    (2, 16, 16, 64, 64, 3, "same", 0),
    (4, 56, 56, 64, 64, 3, "same", 0),
    (3, 32, 32, 64, 128, 3, 1, 0),
    (3, 32, 32, 128, 128, 3, 1, 0),
    (3, 32, 32, 64, 128, 1, 0, 0),
    (2, 20, 17, 24, 40, 3, "same", 1),
    (2, 13, 19, 16, 8, 5, 2, 0),
    (1, 9, 9, 256, 72, 2, 0, 0),
]


def ref(X, W, b, dY, pad4, dil):
    x = torch.tensor(X, dtype=torch.float64, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    w = torch.tensor(W, dtype=torch.float64, device="cuda").permute(3, 2, 0, 1)
    xp = F.pad(x, (pad4[2], pad4[3], pad4[0], pad4[1]))
    z = F.conv2d(xp, w, torch.tensor(b.ravel(), dtype=torch.float64, device="cuda"), dilation=dil + 1)
    z.backward(torch.tensor(dY, dtype=torch.float64, device="cuda").permute(0, 3, 1, 2))
    return z.permute(0, 2, 3, 1).detach().cpu().numpy(), x.grad.permute(0, 2, 3, 1).cpu().numpy()


def err(a, r):
    return float(np.linalg.norm((a - r).ravel()) / max(np.linalg.norm(r.ravel()), 1e-30))


def main():
    rng = np.random.default_rng(0)
    bad = 0
    for dm in ("0",):
        for (N, H, Wd, C, K, k, pad, dil) in SHAPES:
            X = rng.standard_normal((N, H, Wd, C), dtype=np.float32).astype(np.float64)
            W = (rng.standard_normal((k, k, C, K), dtype=np.float32) / np.sqrt(k * k * C)).astype(np.float64)
            b = (0.1 * rng.standard_normal((1, 1, 1, K), dtype=np.float32)).astype(np.float64)
            for mode, tol in (("bf16", 2e-2), ("tf32", 5e-3)):
                L = nb.Conv2D(K, (k, k), pad=pad, dilation=dil, mode=mode)
                L.forward(X, retain_derived=False)
                L.parameters["W"], L.parameters["b"] = W, b
                Y = L.forward(X)
                dY = rng.standard_normal(Y.shape, dtype=np.float32).astype(np.float64)
                dX = L.backward(dY)
                p4 = nb.utils.resolve_pad_2D(X.shape, pad, (k, k), 1, dil)
                Yr, dXr = ref(X, W, b, dY, p4, dil)
                eY, eX = err(Y, Yr), err(dX, dXr)
                ok = eY <= tol and eX <= tol
                bad += (not ok)
                print("{}{} N{} {}x{} C{}->K{} k{} pad={} dil={}: tc={} errY={:.2e} errdX={:.2e} {}".format(
                    "", mode, N, H, Wd, C, K, k, pad, dil, L._uses_tc(), eY, eX, "ok" if ok else "FAIL"), flush=True)
    torch.cuda.synchronize()
    print("slab_check:", "OK" if bad == 0 else "{} failures".format(bad))


if __name__ == "__main__":
    main()
