"""Time the HBM-bound passes of the cfg4 step (SkipConnectionConvModule, N=256, 32x32x128, bf16) one entry point at a
time and print achieved GB/s against the bytes each pass must move.  Buffers rotate over several sets so that a pass
never finds its inputs in L2 from the previous iteration.

    python scripts/elementwise_bench.py [rows] [ch]
"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
nb = importlib.import_module("numpy-ml_b200")
L = nb._lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 32 * 32
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dt = torch.bfloat16
SETS = 4
T = rows * ch * 2          # bytes of one tensor


def bufs(k):
    return [[torch.randn(rows, ch, device="cuda", dtype=torch.float32).to(dt) for _ in range(k)] for _ in range(SETS)]


def timeit(name, fn, nbytes, iters=12):
    for i in range(3):
        fn(i % SETS)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % SETS)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print("{:34s} {:8.1f} us  {:7.0f} GB/s  ({:.0f} MB)".format(name, us, nbytes / us / 1e3, nbytes / 1e6))


lib = L.load()
st = L.stream
ws = L.workspace(64 << 20)

B = bufs(3)
timeit("torch copy (read+write)", lambda i: B[i][1].copy_(B[i][0]), 2 * T)

db = torch.zeros(ch, device="cuda")
relu = 2
timeit("dz_prep relu + db (r2 w1)", lambda i: L.check(lib.npml_dz_prep(rows, ch, relu, 0.0, 0.0, L.ptr(B[i][0]), 1, L.ptr(B[i][1]), 1,
                                                                      L.ptr(B[i][2]), 1, L.ptr(db), 1, L.ptr(ws), ws.numel(), st()), "dz"), 3 * T)
timeit("dz_prep relu (r2 w1)", lambda i: L.check(lib.npml_dz_prep(rows, ch, relu, 0.0, 0.0, L.ptr(B[i][0]), 1, L.ptr(B[i][1]), 1,
                                                                 L.ptr(B[i][2]), 1, None, 0, None, 0, st()), "dz"), 3 * T)
timeit("dz_prep identity db only (r1)", lambda i: L.check(lib.npml_dz_prep(rows, ch, 0, 0.0, 0.0, L.ptr(B[i][0]), 1, None, 0,
                                                                          None, 1, L.ptr(db), 1, L.ptr(ws), ws.numel(), st()), "dz"), T)
import ctypes as C


def add(i):
    arr = (C.c_void_p * 2)(B[i][0].data_ptr(), B[i][1].data_ptr())
    L.check(lib.npml_add_act_fwd(2, arr, 1, rows * ch, 0, 0.0, 0.0, L.ptr(B[i][2]), None, st()), "add")


timeit("add identity (r2 w1)", add, 3 * T)

g = torch.rand(ch, device="cuda") + 0.5
h = torch.zeros(ch, device="cuda")
rm, rv = torch.zeros(ch, device="cuda"), torch.ones(ch, device="cuda")
mean, rstd = torch.empty(ch, device="cuda"), torch.empty(ch, device="cuda")
timeit("bn fwd (r2 w1)", lambda i: L.check(lib.npml_bn2d_fwd(rows, ch, L.ptr(B[i][0]), L.ptr(B[i][1]), 1, L.ptr(g), L.ptr(h), L.ptr(rm),
                                                            L.ptr(rv), 0.9, 1e-5, 1, L.ptr(mean), L.ptr(rstd), L.ptr(ws), ws.numel(), st()), "bn"), 3 * T)
dg, dh = torch.zeros(ch, device="cuda"), torch.zeros(ch, device="cuda")
timeit("bn bwd (r4 w1)", lambda i: L.check(lib.npml_bn2d_bwd(rows, ch, L.ptr(B[i][0]), L.ptr(B[i][1]), L.ptr(B[i][2]), 1, L.ptr(g), L.ptr(mean),
                                                            L.ptr(rstd), L.ptr(dg), L.ptr(dh), 1, L.ptr(ws), ws.numel(), st()), "bnb"), 5 * T)
