"""GPU diagnostic for csrc/tc_ws_conv.cu: the same Conv2D forward / backward with NPML_WS=1 (weight-stationary kernel)
(or the variant given on the command line, `ws_debug.py 2`) and NPML_WS=0 (slab kernel) in one process, per-tap so that a wrong tap / half / shift shows up by name."""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("numpy-ml_b200")
VARIANT = sys.argv[1] if len(sys.argv) > 1 else "1"      # "2": the experimental row-tile / TMA-store variant


def run(layer, X, dY, ws):
    os.environ["NPML_WS"] = VARIANT if ws else "0"
    layer.flush_gradients()
    Y = layer.forward(X)
    dX = layer.backward(dY)
    torch.cuda.synchronize()
    return pkg.to_numpy(Y), pkg.to_numpy(dX)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def case(xs, oc, ks, pad, dil, act=None, per_tap=True):
    rng = np.random.default_rng(1)
    L = pkg.Conv2D(out_ch=oc, kernel_shape=ks, pad=pad, stride=1, dilation=dil, act_fn=act, mode="bf16")
    X = rng.standard_normal(xs, dtype=np.float32)
    L.forward(X, retain_derived=False)
    W = (rng.standard_normal(ks + (xs[3], oc)) / np.sqrt(ks[0] * ks[1] * xs[3])).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, 1, 1, oc))).astype(np.float32)
    L.parameters["W"], L.parameters["b"] = W, b
    dY = rng.standard_normal(pkg.to_numpy(L.forward(X, retain_derived=False)).shape).astype(np.float32)
    Yr, dXr = run(L, X, dY, False)
    Yw, dXw = run(L, X, dY, True)
    print("case", xs, oc, ks, pad, dil, act, "| Y rel diff %.3e | dX rel diff %.3e" % (rel(Yw, Yr), rel(dXw, dXr)), flush=True)
    bad = rel(Yw, Yr) > 1e-2 or rel(dXw, dXr) > 1e-2
    if bad and per_tap:
        for r in range(ks[0]):
            for t in range(ks[1]):
                W1 = np.zeros_like(W)
                W1[r, t] = W[r, t]
                L.parameters["W"] = W1
                Yr1, dXr1 = run(L, X, dY, False)
                Yw1, dXw1 = run(L, X, dY, True)
                print("   tap (%d,%d): Y %.3e dX %.3e" % (r, t, rel(Yw1, Yr1), rel(dXw1, dXr1)), flush=True)
        e = np.abs(Yw - Yr)
        print("   Y err by channel block of 16:", np.round(e.reshape(-1, e.shape[-1]).max(0).reshape(-1, 16).max(1), 3))
        print("   Y err by column xp:", np.round(e.max(axis=(0, 1, 3)), 3))
        print("   Y err by row yp:", np.round(e.max(axis=(0, 2, 3)), 3))
        print("   Y err by image:", np.round(e.max(axis=(1, 2, 3)), 3))
    return bad


if __name__ == "__main__":
    t0 = time.time()
    bad = False
    bad |= case((4, 16, 16, 64), 64, (3, 3), "same", 0)             # pair mode, cfg2 miniature
    bad |= case((5, 10, 10, 64), 128, (3, 3), 1, 0, "ReLU")         # forward single mode, dX pair mode with 2 K blocks
    bad |= case((2, 9, 11, 72), 40, (3, 2), (1, 2), 0, "Sigmoid")   # channel tails
    bad |= case((3, 12, 12, 16), 16, (4, 4), 1, 0)                  # all taps paired
    bad |= case((2, 9, 9, 32), 32, (5, 5), 2, 0, "Tanh")            # 15 groups
    bad |= case((9, 5, 5, 64), 64, (3, 3), "same", 3)               # dilation 3 -> delta 4
    bad |= case((64, 56, 56, 64), 64, (3, 3), "same", 0, per_tap=False)   # many units per CTA
    print("RESULT", "MISMATCH" if bad else "all cases agree", "%.1fs" % (time.time() - t0))
