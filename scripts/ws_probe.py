"""Timing experiments on the weight-stationary kernel (cfg2 forward alone, CUDA events, graph of 10 back-to-back calls, median of 10 replays): which of its
shared-memory clients costs the tensor pipe its idle cycles.  NPML_WS_DBG: 1 = epilogue without shared / global traffic,
2 = no TMA refills after the first two slabs, 3 = both (results are wrong by construction; timing only)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("numpy-ml_b200")
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.standard_normal((256, 56, 56, 64), dtype=np.float32)).cuda().to(torch.bfloat16)
L = pkg.Conv2D(out_ch=64, kernel_shape=(3, 3), pad="same", mode="bf16")
os.environ["NPML_WS"] = "0"
L.forward(X, retain_derived=False)

def t(env):
    for k in ("NPML_WS", "NPML_WS_NT", "NPML_WS_DBG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    REP = 10
    for i in range(3):
        L.forward(X, retain_derived=False)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            L.forward(X, retain_derived=False)
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / REP)
    return float(np.median(ts[2:]))

for env in ({"NPML_WS": "0"},
            {"NPML_WS": "1", "NPML_WS_NT": "128"}, {"NPML_WS": "1", "NPML_WS_NT": "128", "NPML_WS_DBG": "1"},
            {"NPML_WS": "1", "NPML_WS_NT": "128", "NPML_WS_DBG": "2"}, {"NPML_WS": "1", "NPML_WS_NT": "128", "NPML_WS_DBG": "3"},
            {"NPML_WS": "1", "NPML_WS_NT": "96"}, {"NPML_WS": "1", "NPML_WS_NT": "96", "NPML_WS_DBG": "1"},
            {"NPML_WS": "1", "NPML_WS_NT": "96", "NPML_WS_DBG": "2"}, {"NPML_WS": "1", "NPML_WS_NT": "96", "NPML_WS_DBG": "3"},
            {"NPML_WS": "1", "NPML_WS_NT": "64"}, {"NPML_WS": "1", "NPML_WS_NT": "64", "NPML_WS_DBG": "3"}):
    print(env, "%.1f us" % t(env), flush=True)
