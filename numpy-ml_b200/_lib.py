"""ctypes binding of libnpmlconv.so (include/npml_conv.h) and the torch device-buffer plumbing.

There is NO CPU fallback: if the shared library is missing or a call fails, this module raises.
PyTorch is used only to own device memory and streams; every arithmetic op on the hot path is a
kernel of libnpmlconv.so.
"""
import ctypes as C
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnpmlconv.so")

MODES = {"fp32": 0, "tf32": 1, "bf16": 2}
F32, BF16, F64 = 0, 1, 2
OP_CONV_FPROP, OP_CONV_DGRAD, OP_CONV_WGRAD = 0, 1, 2
OP_DECONV_FPROP, OP_DECONV_DGRAD, OP_DECONV_WGRAD = 3, 4, 5
OP_DZ_PREP, OP_BN, OP_CONV_BWD = 6, 7, 8
FLAG_WEIGHTS_PACKED = 1


class ConvDesc(C.Structure):
    """Mirror of `npml_conv_desc` (include/npml_conv.h)."""
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "C", "K", "fr", "fc", "stride", "dil", "pr1", "pr2", "pc1",
                                         "pc2", "OH", "OW", "mode", "act")] + \
               [("act_p0", C.c_float), ("act_p1", C.c_float), ("accumulate", C.c_int32), ("flags", C.c_int32)]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_DESC = C.POINTER(ConvDesc)

# name -> (restype, argtypes); kept in step with include/npml_conv.h (tests check every symbol)
SIGNATURES = {
    "npml_conv2d_fprop": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_fprop_stats": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_fprop_bn": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_dgrad": (C.c_int, [_DESC, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_wgrad": (C.c_int, [_DESC, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_wgrad_db": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_bwd": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_conv2d_bwd_is_fused": (C.c_int, [_DESC]),
    "npml_deconv2d_fprop": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_deconv2d_fprop_bn": (C.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_deconv2d_dgrad": (C.c_int, [_DESC, _P, _P, _P, _P, _SZ, _P]),
    "npml_deconv2d_wgrad": (C.c_int, [_DESC, _P, _P, _P, _P, _SZ, _P]),
    "npml_workspace_bytes": (_SZ, [_DESC, C.c_int]),
    "npml_dz_prep": (C.c_int, [_I64, _I32, _I32, _F, _F, _P, _I32, _P, _I32, _P, _I32, _P, _I32, _P, _SZ, _P]),
    "npml_cast": (C.c_int, [_P, _I32, _P, _I32, _I64, _P]),
    "npml_add_act_fwd": (C.c_int, [_I32, C.POINTER(_P), _I32, _I64, _I32, _F, _F, _P, _P, _P]),
    "npml_mul_act_fwd": (C.c_int, [_I32, C.POINTER(_P), _I32, _I64, _I32, _F, _F, _P, _P, _P]),
    "npml_mul_act_bwd": (C.c_int, [_I32, C.POINTER(_P), C.POINTER(_P), _I32, _I64, _I32, _F, _F, _P, _P, _P]),
    "npml_act_fwd": (C.c_int, [_P, _P, _I32, _I64, _I32, _F, _F, _P]),
    "npml_bn2d_fwd": (C.c_int, [_I64, _I32, _P, _P, _I32, _P, _P, _P, _P, _F, _F, _I32, _P, _P, _P, _SZ, _P]),
    "npml_bn2d_fold": (C.c_int, [_I32, _P, _P, _P, _P, _F, _P, _P, _P]),
    "npml_bn2d_bwd": (C.c_int, [_I64, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _I32, _P, _SZ, _P]),
    "npml_bn2d_stats": (C.c_int, [_I64, _I32, _P, _I32, _P, _P, _SZ, _P]),
    "npml_bn2d_apply": (C.c_int, [_I64, _I64, _I32, _P, _P, _I32, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _SZ, _P]),
    "npml_bn2d_bwd_stats": (C.c_int, [_I64, _I32, _P, _P, _I32, _P, _P, _P, _P, _SZ, _P]),
    "npml_bn2d_bwd_apply": (C.c_int, [_I64, _I64, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _SZ, _P]),
    "npml_bn2d_finalize": (C.c_int, [_I64, _I32, _P, _P, _P, _F, _F, _P, _P, _P]),
    "npml_bn2d_pair_add_act_fwd": (C.c_int, [_I64, _I32, _P, _P, _I32] + [_P] * 8 + [_I32, _F, _F, _P, _P, _P]),
    "npml_bn2d_pair_bwd_stats": (C.c_int, [_I64, _I32, _P, _P, _I32, _F, _F, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "npml_bn2d_pair_bwd_apply": (C.c_int, [_I64, _I64, _I32, _P, _P, _I32, _F, _F, _P, _P, _I32] + [_P] * 6 + [_P, _P, _P, _P]
                                 + [_P] * 4 + [_I32, _P, _SZ, _P]),
    "npml_bn2d_bwd_act": (C.c_int, [_I64, _I64, _I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _F, _F, _P,
                                    _P, _SZ, _P]),
    "npml_sgd_update": (C.c_int, [_P, _P, _P, _I64, _F, _F, _F, _P]),
    "npml_optimizer_update": (C.c_int, [_I32, _P, _P, _P, _P, _I64] + [C.c_double] * 7 + [_P]),
    "npml_optimizer_update_clipped": (C.c_int, [_I32, _P, _P, _P, _P, _I64] + [C.c_double] * 6 + [_P, C.c_double, _P]),
    "npml_sumsq_f32": (C.c_int, [_P, _I64, _P, _P, _SZ, _P]),
    "npml_pool2d_fwd": (C.c_int, [_DESC, _I32, _P, _P, _I32, _P]),
    "npml_pool2d_bwd": (C.c_int, [_DESC, _I32, _P, _P, _P, _I32, _P]),
    "npml_im2col": (C.c_int, [_DESC, _P, _P, _P]),
    "npml_col2im": (C.c_int, [_DESC, _P, _P, _P]),
    "npml_comm_unique_id": (C.c_int, [_P]),
    "npml_comm_init_rank": (C.c_int, [_P, _I32, _I32]),
    "npml_allreduce_sum_f32": (C.c_int, [_P, _I64, _P]),
    "npml_allreduce_sum_f64": (C.c_int, [_P, _I64, _P]),
    "npml_comm_destroy": (C.c_int, []),
    "npml_peer_region_bytes": (_SZ, []),
    "npml_peer_alloc": (C.c_int, [_P]),
    "npml_peer_open": (C.c_int, [_P, _I32, _I32]),
    "npml_peer_ready": (C.c_int, []),
    "npml_peer_allreduce_sum_f32": (C.c_int, [_P, _I64, _I32, _P]),
    "npml_peer_allreduce_sum_f64": (C.c_int, [_P, _I64, _I32, _P]),
    "npml_peer_status": (C.c_int, []),
    "npml_peer_close": (C.c_int, []),
    "npml_last_error": (C.c_char_p, []),
    "npml_uses_tensor_cores": (C.c_int, [_DESC, C.c_int]),
    "npml_debug_ws_plan": (C.c_int, [_DESC, C.c_int, C.c_int, _P, _I32]),
    "npml_debug_rt_classes": (C.c_int, [_DESC, C.c_int, _P, _I32]),
    "npml_launch_count": (_I64, [_I32]),
    "npml_version": (C.c_int, []),
}

_lib = None
# True: a packed-weights workspace may be reused by calls recorded into a CUDA graph (layers._ConvBase._call_w); only
# for captures whose weights stay constant across replays (bench.py's forward + backward step has no optimizer update)
GRAPH_REUSE_PACKED = False
# when a list, the layers append (entry point, start event, end event) around every C-ABI call
PROFILE = None
# True while a step is being captured into a CUDA graph: the events must be "external" so that they become
# event-record nodes of the graph (and can be timed after a replay)
PROFILE_EXTERNAL = False


class profiled(object):
    """Context manager: CUDA events on the launching stream around one C-ABI call when PROFILE is a list."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE is not None:
            kw = {"external": True} if PROFILE_EXTERNAL else {}
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True, **kw), torch.cuda.Event(enable_timing=True, **kw)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and exc[0] is None:
            self.e1.record()
            PROFILE.append((self.name, self.e0, self.e1))
        return False


def load():
    """dlopen libnpmlconv.so and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libnpmlconv.so is not built (run `python numpy-ml_b200/build.py`); "
                           "there is no CPU fallback for the convolution hot path")
    try:
        import nvidia.nccl as _n  # torch's bundled NCCL, for npml_comm_*
        cand = os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ.setdefault("NPML_NCCL_LIB", cand)
    except Exception:
        pass
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        raise RuntimeError("{} failed ({}): {}".format(what, status, load().npml_last_error().decode()))


# ---------------------------------------------------------------- device buffers -----------------
def device():
    if not torch.cuda.is_available():
        raise RuntimeError("numpy-ml_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def torch_dtype(mode):
    return torch.bfloat16 if mode == "bf16" else torch.float32


_F64_ON_DEVICE = 1 << 16     # elements from which float64 <-> device dtype conversions run on the device


def code_of(t):
    return {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64}[t.dtype]


def cast(t, dtype):
    """Device-side dtype conversion through npml_cast (no torch arithmetic)."""
    if t.dtype == dtype:
        return t
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    check(load().npml_cast(ptr(t), code_of(t), ptr(out), code_of(out), t.numel(), stream()), "npml_cast")
    return out


def to_device(x, dtype):
    """NumPy array (any float dtype) or torch tensor -> contiguous device tensor of `dtype`."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64 and x.size >= _F64_ON_DEVICE:
            # the reference's own dtype: ship the float64 bytes and round on the device (npml_cast, round-to-nearest like
            # ndarray.astype) -- the host-side astype of a 51 M element batch costs more than the extra PCIe bytes
            t = torch.from_numpy(np.ascontiguousarray(x)).to(device(), non_blocking=False)
        else:
            t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device(), non_blocking=False)
    elif isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.to(device())
        if not t.is_contiguous():
            t = t.contiguous()
        if t.dtype not in (torch.float32, torch.bfloat16, torch.float64):
            raise TypeError("unsupported tensor dtype {}".format(t.dtype))
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32)).to(device())
    return cast(t, dtype)


def to_numpy(t):
    """Device tensor -> float64 NumPy array (what the reference hands back)."""
    if isinstance(t, np.ndarray):
        return t
    if isinstance(t, (list, tuple)):
        return [to_numpy(v) for v in t]
    t = t.detach()
    if t.numel() >= _F64_ON_DEVICE:
        return cast(t, torch.float64).cpu().numpy()          # widen on the device, not with a host-side astype
    return cast(t, torch.float32).cpu().numpy().astype(np.float64)


_ws = {}


def workspace(nbytes):
    """Grow-only per-device scratch buffer handed to the library (the library never allocates)."""
    dev = torch.cuda.current_device()
    buf = _ws.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device())
        _ws[dev] = buf
    return buf


def launch_count(reset=False):
    return int(load().npml_launch_count(1 if reset else 0))
