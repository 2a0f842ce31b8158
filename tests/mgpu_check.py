"""Multi-GPU check of the peer-memory all-reduce and of the sharded step (run on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Not collected by pytest (the round-end GPU test box has one GPU); bench.py's `dp_parity` repeats the end-to-end part of
it at every N of the driver's scaling run."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    pkg = importlib.import_module("numpy-ml_b200")
    D, L = pkg.dist, pkg._lib
    st = D.init()
    world, rank = st["world"], st["rank"]
    dev = torch.device("cuda", torch.cuda.current_device())
    if rank == 0:
        print("collective:", D.collective_name(), flush=True)
    assert st["peer"], D.collective_name()
    lib = L.load()
    ok = True
    # 1. values: every size class (ragged tails, one block, many blocks), both dtypes, both channels, many calls in a row
    for it, (n, dt) in enumerate([(1, torch.float32), (5, torch.float32), (320, torch.float32), (36928, torch.float32),
                                  (524416, torch.float32), (1000003, torch.float32), (256, torch.float64),
                                  (383, torch.float64), (36928, torch.float32), (128, torch.float32)] * 3):
        g = torch.Generator(device=dev); g.manual_seed(100 * it + rank)
        x = torch.randn(n + 4, generator=g, device=dev, dtype=dt)[4:] if dt == torch.float32 and n % 4 == 0 else \
            torch.randn(n, generator=g, device=dev, dtype=dt)
        want = torch.zeros(n, dtype=torch.float64, device=dev)
        for r in range(world):                       # the sum in rank order, recomputed locally from every rank's seed
            gr = torch.Generator(device=dev); gr.manual_seed(100 * it + r)
            xr = torch.randn(n + 4, generator=gr, device=dev, dtype=dt)[4:] if dt == torch.float32 and n % 4 == 0 else \
                torch.randn(n, generator=gr, device=dev, dtype=dt)
            want = (want.to(dt) + xr).double()
        got = x.clone()
        D._allreduce(got, it % 2)
        torch.cuda.synchronize()
        via_nccl = n * got.element_size() > D.PEER_MAX_BYTES          # NCCL adds in its own order: compare with a tolerance
        good = torch.allclose(got.double(), want, rtol=0, atol=1e-5) if via_nccl else torch.equal(got.double(), want)
        if not good:
            ok = False
            print("rank {}: MISMATCH n={} dtype={} max err {}".format(rank, n, dt, (got.double() - want).abs().max().item()), flush=True)
    # 2. CUDA graph replay: the call counter lives in device memory
    buf = torch.ones(4096, device=dev) * (rank + 1)
    src = buf.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            buf.copy_(src)
            D._allreduce(buf, D.CH_SIDE)
        for _ in range(7):
            g.replay()
    torch.cuda.synchronize()
    if not torch.equal(buf, torch.full_like(buf, world * (world + 1) / 2)):
        ok = False
        print("rank {}: graph replay mismatch {}".format(rank, buf[:4].tolist()), flush=True)
    # 3. latency of the gradient-sized messages against NCCL (device time, max over ranks is taken by the reader)
    for n in (320, 36928, 295168, 524416):
        x = torch.randn(n, device=dev) * 1e-3
        res = {}
        for name, fn in (("peer", lambda: lib.npml_peer_allreduce_sum_f32(L.ptr(x), n, 0, L.stream())),
                         ("nccl", lambda: lib.npml_allreduce_sum_f32(L.ptr(x), n, L.stream()))):
            for _ in range(5):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50):
                fn()
            b.record()
            torch.cuda.synchronize()
            res[name] = D.all_reduce_max(a.elapsed_time(b) / 50 * 1e3)
        if rank == 0:
            print("allreduce {:>7} fp32 ({:>7.1f} KB): peer {:6.1f} us   nccl {:6.1f} us".format(n, n * 4 / 1024, res["peer"], res["nccl"]), flush=True)
    assert lib.npml_peer_status() == 0
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU CHECK", "PASSED" if int(flag[0]) else "FAILED", flush=True)
    D.destroy()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
