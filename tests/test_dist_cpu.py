"""World-size-2 `gloo` checks of the data-parallel host logic on CPU (no kernels): the flat gradient bucket, its SUM
all-reduce (the reference sums gradients over the batch, layers/layers.py:3087-3089 -- never an average), the
float64 all-reduce used for sync-BN statistics, the max-over-ranks timing helper and the batch sharding."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeLayer(object):
    """Stands in for a Conv2D / BatchNorm2D layer: GradBucket only needs the `gradients` dict."""

    def __init__(self, shapes, fill):
        self.gradients = {k: torch.full(s, float(fill)) for k, s in shapes.items()}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import importlib
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    D = importlib.import_module("numpy-ml_b200").dist
    st = D.init(backend="gloo")
    assert st["world"] == world and D.rank() == rank and D.world_size() == world
    conv = _FakeLayer({"W": (3, 3, 2, 4), "b": (1, 1, 1, 4)}, rank + 1)
    bn = _FakeLayer({"scaler": (4,), "intercept": (4,)}, 10 * (rank + 1))
    bucket = D.GradBucket([conv, bn], overlap=True)          # overlap needs CUDA: must degrade to the plain path
    assert not bucket.overlap
    assert bucket.flat.numel() == 72 + 4 + 4 + 4
    assert bucket.owns_only([conv, bn]) and not bucket.owns_only([conv])
    # the layers' gradients are views into the bucket
    conv.gradients["W"][0, 0, 0, 0] += 100.0 * (rank + 1)
    bucket.allreduce()
    total = sum(r + 1 for r in range(world))
    assert float(conv.gradients["W"][0, 0, 0, 0]) == total + 100.0 * total      # SUM, not mean
    assert torch.all(conv.gradients["b"] == total)
    assert torch.all(bn.gradients["scaler"] == 10 * total)
    bucket.zero_layer(conv)
    assert float(conv.gradients["W"].abs().sum()) == 0.0 and torch.all(bn.gradients["intercept"] == 10 * total)
    # sync-BN statistics: float64 SUM
    s = torch.tensor([1.5 * (rank + 1), 1e-9 * (rank + 1)], dtype=torch.float64)
    D.allreduce_sum_f64(s)
    assert s[0].item() == 1.5 * total and abs(s[1].item() - 1e-9 * total) < 1e-24
    assert D.all_reduce_max(float(rank)) == float(world - 1)
    lo, hi = D.shard_batch(10)
    np.save(os.path.join(out_dir, "span{}.npy".format(rank)), np.array([lo, hi]))
    D.barrier()
    D.destroy()


def test_gradient_bucket_and_reductions_world2(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spans = [np.load(os.path.join(str(tmp_path), "span{}.npy".format(r))) for r in range(world)]
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 10
