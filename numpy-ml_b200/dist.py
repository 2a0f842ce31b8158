"""Batch-sharded data parallelism for the convolution layers: one process per GPU, every rank runs
forward/backward on its N/g slice of the batch and the summed dW / db (and BatchNorm gradient) buffers
are exchanged with ONE NCCL all-reduce per step.  Y and dX are per-example, so they need
no communication.  The reference sums gradients over the batch (layers/layers.py:3087-3089), so the
reduction is SUM, never an average.

torch.distributed is used only for rendezvous (broadcasting the ncclUniqueId); the all-reduce itself
is npml_allreduce_sum_f32 in libnpmlconv.so on the caller's stream.  With backend "gloo" (CPU tests)
the same bucket logic runs through torch.distributed.all_reduce."""
import ctypes as C
import os

import torch

from . import _lib as L

_state = {"ready": False, "world": 1, "rank": 0, "backend": None, "peer": False, "peer_why": "single process"}
PEER_MAX_BYTES = 5 << 19           # csrc/peer.cu kMaxMsgBytes: larger buffers go to ncclAllReduce
CH_MAIN, CH_SIDE = 0, 1            # peer all-reduce channels: the caller's stream / the gradient side stream


def init(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun style)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    _state.update(world=world, rank=rank, backend=backend)
    if world == 1:
        _state["ready"] = True
        return _state
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", str(rank))))
    if not dist.is_initialized():
        # gloo for the control plane (id broadcast, barriers over CPU tensors); data plane = our NCCL comm
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    if backend == "nccl":
        lib = L.load()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            L.check(lib.npml_comm_unique_id(buf), "npml_comm_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        dist.broadcast(ident, src=0)
        raw = (C.c_ubyte * 128)(*ident.tolist())
        L.check(lib.npml_comm_init_rank(raw, world, rank), "npml_comm_init_rank")
        _init_peer(lib, world, rank)
    _state["ready"] = True
    return _state


def _init_peer(lib, world, rank):
    """Map every rank's peer region (csrc/peer.cu) into this process.  All ranks agree on the outcome: if any of them
    cannot (no peer access between two of the GPUs, IPC disabled, NPML_PEER=0) everybody stays on NCCL."""
    import torch.distributed as dist
    ok, why = 1, ""
    handle = (C.c_ubyte * 64)()
    if os.environ.get("NPML_PEER", "1") == "0" or world > 8:
        ok, why = 0, "disabled (NPML_PEER=0)" if world <= 8 else "more than 8 ranks"
    elif lib.npml_peer_alloc(handle) != 0:
        ok, why = 0, lib.npml_last_error().decode()
    mine = torch.tensor(list(handle) + [ok], dtype=torch.uint8)
    every = [torch.zeros(65, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(every, mine)
    if all(int(t[64]) for t in every):
        raw = (C.c_ubyte * (64 * world))(*torch.cat([t[:64] for t in every]).tolist())
        if lib.npml_peer_open(raw, world, rank) != 0:
            ok, why = 0, lib.npml_last_error().decode()
    else:
        ok, why = 0, why or "a peer rank could not allocate its region"
    agree = torch.tensor([ok], dtype=torch.int32)
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)          # also the barrier npml_peer_open asks for
    _state["peer"] = bool(int(agree[0]))
    _state["peer_why"] = "peer-memory push all-reduce kernel (csrc/peer.cu)" if _state["peer"] else \
        "ncclAllReduce ({})".format(why or "a peer rank could not map the regions")


def collective_name():
    return _state["peer_why"] if _state["world"] > 1 else "none (single process)"


def _allreduce(buf, channel):
    """In-place SUM over ranks of a contiguous fp32 / fp64 device tensor on the current stream."""
    lib, n = L.load(), buf.numel()
    f64 = buf.dtype == torch.float64
    # dW || db of a layer and the sync-BN statistics (<= 2.5 MB): the one-kernel peer-memory exchange; anything larger
    # is bandwidth-bound and goes to NCCL's ring / in-switch reduction
    if _state["peer"] and n * buf.element_size() <= PEER_MAX_BYTES:
        fn = lib.npml_peer_allreduce_sum_f64 if f64 else lib.npml_peer_allreduce_sum_f32
        L.check(fn(L.ptr(buf), n, channel, L.stream()), "npml_peer_allreduce_sum")
    else:
        fn = lib.npml_allreduce_sum_f64 if f64 else lib.npml_allreduce_sum_f32
        L.check(fn(L.ptr(buf), n, L.stream()), "npml_allreduce_sum")


def world_size():
    return _state["world"]


def rank():
    return _state["rank"]


def shard_batch(n_ex, world=None, r=None, even=False):
    """Contiguous [start, stop) slice of the batch axis owned by rank `r`.  `even=True` (what sync-BN needs: it
    forms the global row count as rows * world) raises unless every rank gets the same number of examples."""
    world = _state["world"] if world is None else world
    r = _state["rank"] if r is None else r
    if even and n_ex % world != 0:
        raise ValueError("batch of {} examples does not split evenly over {} ranks".format(n_ex, world))
    per = (n_ex + world - 1) // world
    return min(r * per, n_ex), min((r + 1) * per, n_ex)


_equal_rows_checked = set()


def require_equal_shards(rows):
    """Sync-BN normalises with `rows * world` as the global count, so every rank must hold the same number of rows
    (an empty or ragged shard would silently skew mean / variance / dX, and an empty one deadlocks the collective).
    Checked with one host-side exchange the first time a row count is seen, never again for that count (so a
    captured CUDA graph replays without it)."""
    if _state["world"] == 1 or rows in _equal_rows_checked:
        return
    import torch.distributed as dist
    t = torch.tensor([rows, -rows], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if int(t[0]) != -int(t[1]):
        raise ValueError("sync BatchNorm2D needs equal shards on every rank: this rank has {} rows, the job has "
                         "between {} and {}".format(rows, -int(t[1]), int(t[0])))
    _equal_rows_checked.add(rows)


class GradBucket(object):
    """One flat fp32 buffer holding every gradient of a set of layers; the layers' `gradients` entries
    become views into it, so kernels accumulate straight into the bucket.

    overlap=False: `allreduce()` sums the whole bucket over ranks with ONE call on the current stream.
    overlap=True : every Conv2D / Deconv2D layer of the bucket announces its dW / db as soon as its wgrad
        kernel is enqueued (layers._ConvBase._bwd computes parameter gradients before dX); that layer's slice is
        all-reduced on a side stream while dX is still being computed, and `allreduce()` only reduces what is
        left (BatchNorm gradients) and joins the side stream.  Valid when there is ONE backward per step
        (gradients that keep accumulating over several backward calls must use overlap=False)."""

    def __init__(self, layers, overlap=False):
        self.layers = list(layers)
        entries = []
        for l in self.layers:
            for k in sorted(l.gradients):
                entries.append((l, k, l.gradients[k]))
        total = sum(g.numel() for _, _, g in entries)
        dev = entries[0][2].device if entries else "cpu"
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.slices = {}           # id(layer) -> (offset, numel) of its contiguous run in the bucket
        off = 0
        for l, k, g in entries:
            view = self.flat[off:off + g.numel()].view(g.shape)
            view.copy_(g)
            l.gradients[k] = view
            o0, n0 = self.slices.get(id(l), (off, 0))
            self.slices[id(l)] = (o0, n0 + g.numel())
            off += g.numel()
        for l in self.layers:
            l._bucket = self
        self.overlap = bool(overlap) and _state["world"] > 1 and self.flat.is_cuda
        self._done, self._dirty = set(), set()
        self.suspended = False     # True: steps run without any reduction (a caller measuring the step alone)
        self._side = torch.cuda.Stream() if self.overlap else None
        if self.overlap:
            for l in self.layers:
                if hasattr(l, "_bwd") and hasattr(l, "kernel_shape"):
                    l._grads_ready_hook = self._on_ready

    def zero_layer(self, layer):
        """Zero the contiguous slice that holds every gradient of `layer`."""
        off, n = self.slices.get(id(layer), (0, 0))
        if n:
            self.flat[off:off + n].zero_()

    def owns_only(self, layers):
        """True if the bucket holds the gradients of exactly the layers of `layers` that have any."""
        mine = {id(l) for l in self.layers}
        theirs = {id(l) for l in layers if l.gradients}
        return mine == theirs

    def _reduce(self, off, n, channel=CH_MAIN):
        buf = self.flat[off:off + n]
        if _state["backend"] == "nccl":
            _allreduce(buf, channel)
        else:
            import torch.distributed as dist
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)

    def _on_ready(self, layer):
        """dW / db of `layer` are enqueued on the current stream: reduce that slice on the side stream."""
        if self.suspended:
            return
        if id(layer) in self._done or len(layer.X) > 1:
            # a second backward into the same slice (list-valued backward, index=): an early reduce would mix reduced
            # and local contributions -- leave this layer to the deferred reduce in allreduce()
            self._dirty.add(id(layer))
            return
        off, n = self.slices[id(layer)]
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self._reduce(off, n, CH_SIDE)
        self._done.add(id(layer))

    def allreduce(self):
        """SUM over ranks, in place; on return the current stream is ordered after every reduction."""
        if _state["world"] == 1:
            return self.flat
        if not self.overlap:
            self._reduce(0, self.flat.numel())
            return self.flat
        both = self._done & self._dirty
        if both:
            raise RuntimeError("GradBucket(overlap=True): a layer's gradients were reduced early and then accumulated "
                               "into again in the same step; use overlap=False when one step runs several backward "
                               "calls per layer")
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            for l in self.layers:
                if id(l) not in self._done:
                    self._reduce(*self.slices[id(l)], CH_SIDE)
        torch.cuda.current_stream().wait_stream(self._side)
        self._done, self._dirty = set(), set()
        return self.flat


def allreduce_sum_f64(t):
    """In-place SUM over ranks of a small float64 device tensor (sync-BN statistics) on the current stream."""
    if _state["world"] == 1:
        return t
    if _state["backend"] == "nccl":
        _allreduce(t, CH_MAIN)
    else:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def barrier():
    if _state["world"] > 1:
        import torch.distributed as dist
        dist.barrier()


def all_reduce_max(value):
    """max over ranks of a host float (for max-over-ranks timing)."""
    if _state["world"] == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def destroy():
    import torch.distributed as dist
    if _state["backend"] == "nccl" and _state["world"] > 1:
        if _state["peer"]:
            if L.load().npml_peer_status() != 0:
                raise RuntimeError("a peer all-reduce timed out waiting for another rank (results are invalid)")
            L.load().npml_peer_close()
        L.load().npml_comm_destroy()
    if dist.is_initialized():
        dist.destroy_process_group()
    _state.update(ready=False, world=1, rank=0, peer=False)
