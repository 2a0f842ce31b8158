"""What the second output of the convolution epilogue costs: cfg4's conv1 (256 x 32 x 32 x 64 -> 128, 3x3) forward with
an identity activation (one output tensor) and with ReLU (Z and Y), and conv2 (128 -> 128) for scale; CUDA graph of
10 back-to-back calls, median of 10 replays."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("numpy-ml_b200")
rng = np.random.default_rng(0)

def t(fn):
    REP = 10
    for i in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            fn()
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / REP)
    return float(np.median(ts[2:]))

for cin, cout in ((64, 128), (128, 128), (64, 64)):
    X = torch.from_numpy(rng.standard_normal((256, 32, 32, cin), dtype=np.float32)).cuda().to(torch.bfloat16)
    for act in (None, "relu", "tanh"):
        L = pkg.Conv2D(out_ch=cout, kernel_shape=(3, 3), pad=1, act_fn=act, mode="bf16")
        L.forward(X, retain_derived=False)
        print(cin, cout, act, "retain (Z and Y kept): %.1f us" % t(lambda: (L.flush_gradients(zero=False), L.forward(X, retain_derived=True))),
              "| inference (Y only): %.1f us" % t(lambda: L.forward(X, retain_derived=False)), flush=True)
