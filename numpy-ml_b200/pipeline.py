"""Host-resident batches through the layer API: the end-to-end path a NumPy user of the reference sees.

The reference's `Conv2D.forward(X)` / `backward(dLdy)` take and return host arrays (layers/layers.py:3006-3116).
`HostPipeline` is that call pair for a batch that lives in PINNED host memory: the batch is cut into chunks along the
example axis, every chunk's host->device copy, forward / backward kernels and device->host copy run on three streams
(copy-in, compute, copy-out), so the two PCIe directions and the kernels overlap, and dW / db accumulate over the
chunks exactly like the reference's `+=` over several retained forwards (layers.py:3087-3091).  A module with
BatchNorm2D needs whole-batch statistics and is run as ONE chunk.

`bind_to_gpu_numa_node()` pins the calling process (and so the pinned buffers it allocates afterwards) to the CPU
socket the GPU hangs off; with eight ranks per box the host side of PCIe is otherwise the bottleneck.
"""
import os

import numpy as np
import torch

from . import _lib as L


def gpu_numa_cpus(index=None):
    """CPU ids of the NUMA node the CUDA device `index` is attached to (sysfs), or None if unknown."""
    try:
        index = torch.cuda.current_device() if index is None else index
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/{:04x}:{:02x}:{:02x}.0/".format(dom, bus, dev)
        node = int(open(path + "numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node{}/cpulist".format(node)).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return sorted(cpus)
    except Exception:
        return None


def bind_to_gpu_numa_node(index=None):
    """Restrict this process to the CPUs next to its GPU.  Returns the CPU list used (None: left unbound)."""
    cpus = gpu_numa_cpus(index)
    if not cpus:
        return None
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def pinned(shape, dtype=torch.float32):
    return torch.empty(tuple(shape), dtype=dtype).pin_memory()


class HostPipeline(object):
    """forward + backward of one layer / module over pinned host buffers.

        pipe = HostPipeline(layer, chunks=4)
        pipe.step(Xh, dYh, Yh, dXh, gradh)      # asynchronous; torch.cuda.current_stream() is ordered after it

    Xh, dYh: pinned host tensors (fp32, bf16 or fp16); Yh, dXh: pinned host tensors that receive the results (any of
    those dtypes); gradh (optional): pinned fp32 vector that receives the flat gradient bucket `flat` (or the
    concatenated dW, db).  `reduce` (optional) is called on the compute stream after the last backward chunk (the
    data-parallel all-reduce)."""

    def __init__(self, model, chunks=4, flat=None, reduce=None):
        self.model, self.flat, self.reduce = model, flat, reduce
        self.single = not hasattr(model, "_bwd") or not hasattr(model, "kernel_shape")     # modules: one chunk
        self.chunks = 1 if self.single else int(chunks)
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self._dev_in = {}
        self.act_dtype = L.torch_dtype(getattr(model, "mode", "fp32"))

    def _staging(self, key, shape, dtype):
        buf = self._dev_in.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.dtype != dtype:
            buf = torch.empty(tuple(shape), dtype=dtype, device=L.device())
            self._dev_in[key] = buf
        return buf

    @staticmethod
    def _to(t, dtype):
        if t.dtype == dtype:
            return t
        if t.dtype in (torch.float32, torch.bfloat16) and dtype in (torch.float32, torch.bfloat16):
            return L.cast(t, dtype)
        return t.to(dtype)                                  # fp16 host formats: torch's conversion kernel

    def step(self, Xh, dYh, Yh, dXh, gradh=None):
        model, n = self.model, Xh.shape[0]
        bounds = [(i * n // self.chunks, (i + 1) * n // self.chunks) for i in range(self.chunks)]
        cur = torch.cuda.current_stream()
        self.s_in.wait_stream(cur)
        self.s_out.wait_stream(cur)
        ev_x, ev_dy, xs, dys = [], [], [], []
        with torch.cuda.stream(self.s_in):
            for i, (lo, hi) in enumerate(bounds):
                xs.append(self._staging(("x", i), (hi - lo,) + tuple(Xh.shape[1:]), Xh.dtype))
                xs[-1].copy_(Xh[lo:hi], non_blocking=True)
                ev_x.append(torch.cuda.Event()); ev_x[-1].record()
            for i, (lo, hi) in enumerate(bounds):
                dys.append(self._staging(("dy", i), (hi - lo,) + tuple(dYh.shape[1:]), dYh.dtype))
                dys[-1].copy_(dYh[lo:hi], non_blocking=True)
                ev_dy.append(torch.cuda.Event()); ev_dy[-1].record()
        model.flush_gradients()
        keep = []
        for i, (lo, hi) in enumerate(bounds):
            cur.wait_event(ev_x[i])
            Y = model.forward(self._to(xs[i], self.act_dtype))
            Yo = self._to(Y, Yh.dtype)
            keep.append(Yo)
            ev = torch.cuda.Event(); ev.record()
            self.s_out.wait_event(ev)
            with torch.cuda.stream(self.s_out):
                Yh[lo:hi].copy_(Yo, non_blocking=True)
        for i, (lo, hi) in enumerate(bounds):
            cur.wait_event(ev_dy[i])
            dyd = self._to(dys[i], self.act_dtype)
            dX = model.backward(dyd) if self.chunks == 1 else model.backward(dyd, index=i)
            dXo = self._to(dX, dXh.dtype)
            keep.append(dXo)
            ev = torch.cuda.Event(); ev.record()
            self.s_out.wait_event(ev)
            with torch.cuda.stream(self.s_out):
                dXh[lo:hi].copy_(dXo, non_blocking=True)
        if self.reduce is not None:
            self.reduce()
        if gradh is not None:
            flat = self.flat
            if flat is None:
                flat = torch.cat([model.gradients[k].reshape(-1) for k in sorted(model.gradients)])
            keep.append(flat)
            ev = torch.cuda.Event(); ev.record()
            self.s_out.wait_event(ev)
            with torch.cuda.stream(self.s_out):
                gradh.copy_(flat, non_blocking=True)
        cur.wait_stream(self.s_out)
        for t in keep:                       # the copies on s_out read these after this function returns
            t.record_stream(self.s_out)
        return keep

    def bytes_per_step(self, Xh, dYh, Yh, dXh, gradh=None):
        h2d = Xh.numel() * Xh.element_size() + dYh.numel() * dYh.element_size()
        d2h = Yh.numel() * Yh.element_size() + dXh.numel() * dXh.element_size()
        if gradh is not None:
            d2h += gradh.numel() * gradh.element_size()
        return int(h2d), int(d2h)


def dropin_step(model, X, dLdy):
    """The plain drop-in call pair of the reference: float64 NumPy in, float64 NumPy out
    (layers/layers.py:3006-3116); used by bench.py to report what an unmodified NumPy caller gets."""
    assert isinstance(X, np.ndarray) and isinstance(dLdy, np.ndarray)
    model.flush_gradients()
    Y = model.forward(X)
    dX = model.backward(dLdy)
    return Y, dX
