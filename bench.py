#!/usr/bin/env python
"""Headline benchmark of the numpy-ml convolution hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--mode bf16] [--impl reference]

One "step" = Conv2D.forward + Conv2D.backward (dX, dW, db) over one batch of synthetic input
(SURVEY.md 8d); under N > 1 every rank owns one batch shard (weak scaling: 256 images per GPU for cfg2)
and the step ends with ONE all-reduce (SUM) of the flat dW||db bucket.  Rank 0 prints one JSON line.

  value          images/s, whole job, inputs resident in HBM in the mode's storage dtype
  e2e            the same step through the public layer API with HOST (pinned) buffers: X and dLdy are
                 copied host->device and Y, dX, dW, db device->host inside the timed region
  roofline       the dominant kernel (largest share of the step), timed live with CUDA events on the
                 launching stream inside the timed region
  cpu_baseline   the CPU oracle (a restatement of the reference's im2col + GEMM + np.add.at algorithm,
                 oracle/conv_oracle.py) timed on this box's host cores on a bounded sample
  --impl reference  times that CPU path alone (the reference is pure Python/NumPy)
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the five BASELINE.json configs as concrete calls; per-GPU batch
CONFIGS = {
    "cfg1": dict(kind="conv", N=32, H=28, W=28, C=1, K=32, k=(3, 3), pad="same", stride=1, dilation=0,
                 name="Conv2D 3x3 s1 same, N=32, 28x28x1->32"),
    "cfg2": dict(kind="conv", N=256, H=56, W=56, C=64, K=64, k=(3, 3), pad="same", stride=1, dilation=0,
                 name="Conv2D 3x3 s1 same, N=256, 56x56x64->64"),
    "cfg3": dict(kind="conv", N=128, H=28, W=28, C=128, K=256, k=(3, 3), pad=0, stride=2, dilation=2,
                 name="Conv2D 3x3 s2 dilation2, N=128, 28x28x128->256"),
    "cfg4": dict(kind="skip", N=256, H=32, W=32, C=64, K=128, k=(3, 3), pad=1, stride=1, dilation=0,
                 name="SkipConnectionConvModule 3x3/3x3/1x1 + 3 BatchNorm2D + ReLU, N=256, 32x32x64->128"),
    "cfg5": dict(kind="deconv", N=128, H=16, W=16, C=256, K=128, k=(4, 4), pad=1, stride=2, dilation=0,
                 name="Deconv2D 4x4 s2 pad1, N=128, 16x16x256->128"),
}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p.update(json.load(open(f)))
            p["source"] = "measured"
        except Exception:
            pass
    return p


def out_hw(cfg):
    pkg_u = importlib.import_module("numpy-ml_b200").utils
    if cfg["kind"] == "skip":
        return cfg["H"], cfg["W"]
    if cfg["kind"] == "conv":
        p4 = pkg_u.resolve_pad_2D((cfg["N"], cfg["H"], cfg["W"], cfg["C"]), cfg["pad"], cfg["k"], cfg["stride"],
                                  cfg["dilation"])
        return pkg_u.conv_out_hw(cfg["H"], cfg["W"], cfg["k"][0], cfg["k"][1], p4, cfg["stride"], cfg["dilation"])
    _, oh, ow = pkg_u.deconv_geometry(cfg["H"], cfg["W"], cfg["k"][0], cfg["k"][1], cfg["pad"], cfg["stride"])
    return oh, ow


def flops_per_pass(cfg, n):
    """Nominal MACs*2 of ONE pass (fwd, dgrad or wgrad), padded taps counted."""
    oh, ow = out_hw(cfg)
    taps = cfg["k"][0] * cfg["k"][1]
    if cfg["kind"] == "skip":      # conv1 (C->K, 3x3) + conv2 (K->K, 3x3) + skip (C->K, 1x1); BN / Add are not counted
        return 2.0 * n * oh * ow * cfg["K"] * (cfg["C"] * taps + cfg["K"] * taps + cfg["C"])
    if cfg["kind"] == "conv":
        return 2.0 * n * oh * ow * cfg["K"] * cfg["C"] * taps
    return 2.0 * n * cfg["H"] * cfg["W"] * taps * cfg["C"] * cfg["K"]


def min_bytes_per_pass(cfg, n, esize):
    oh, ow = out_hw(cfg)
    x = n * cfg["H"] * cfg["W"] * cfg["C"] * esize
    y = n * oh * ow * cfg["K"] * esize
    w = cfg["k"][0] * cfg["k"][1] * cfg["C"] * cfg["K"] * 4
    return float(x + y + w)


# ---------------------------------------------------------------------------------------------------
#  CPU arm: the oracle port of the reference algorithm
# ---------------------------------------------------------------------------------------------------
def make_inputs(cfg, n, seed):
    rng = np.random.default_rng(seed)
    oh, ow = out_hw(cfg)
    X = rng.standard_normal((n, cfg["H"], cfg["W"], cfg["C"]), dtype=np.float32)
    dY = rng.standard_normal((n, oh, ow, cfg["K"]), dtype=np.float32)
    from oracle import conv_oracle as O
    if cfg["kind"] == "skip":
        return X, dY, skip_params(cfg, seed), None
    np.random.seed(seed)
    W = O.glorot_uniform(cfg["k"] + (cfg["C"], cfg["K"])).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, 1, 1, cfg["K"]))).astype(np.float32)
    return X, dY, W, b


def skip_params(cfg, seed):
    """Parameters of the cfg4 module (same on every rank): glorot conv weights, small biases, BN affine."""
    from oracle import conv_oracle as O
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    C, K = cfg["C"], cfg["K"]
    P = {}
    for t, shape in (("1", (3, 3, C, K)), ("2", (3, 3, K, K)), ("s", (1, 1, C, K))):
        P["W" + t] = O.glorot_uniform(shape).astype(np.float32).astype(np.float64)
        P["b" + t] = (0.1 * rng.standard_normal((1, 1, 1, K))).astype(np.float32).astype(np.float64)
        P["g" + t] = rng.uniform(0.5, 1.5, K).astype(np.float32).astype(np.float64)
        P["h" + t] = (0.1 * rng.standard_normal(K)).astype(np.float32).astype(np.float64)
        P["rm" + t], P["rv" + t] = np.zeros(K), np.ones(K)
    return P


def cpu_step(cfg, X, dY, W, b):
    """forward + backward of the reference algorithm in float64 (oracle port)."""
    from oracle import conv_oracle as O
    if cfg["kind"] == "skip":
        return O.skip_conv_module_fwd_bwd(X.astype(np.float64), dY.astype(np.float64), W, (3, 3), (3, 3), (1, 1), 1, 1,
                                          1, 1, 1, "relu")
    X, dY, W, b = (a.astype(np.float64) for a in (X, dY, W, b))
    if cfg["kind"] == "conv":
        Y, Z = O.conv2d_layer_fwd(X, W, b, cfg["stride"], cfg["pad"], cfg["dilation"])
        return O.conv2d_layer_bwd(dY, X, Z, W, cfg["stride"], cfg["pad"], cfg["dilation"])
    Y, Z = O.deconv2d_layer_fwd(X, W, b, cfg["stride"], cfg["pad"])
    return O.deconv2d_layer_bwd(dY, X, Z, W, cfg["stride"], cfg["pad"])


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def time_cpu(cfg, n_sample, reps, warmup, seed):
    X, dY, W, b = make_inputs(cfg, n_sample, seed)
    for _ in range(warmup):
        cpu_step(cfg, X, dY, W, b)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_step(cfg, X, dY, W, b)
        ts.append(time.perf_counter() - t0)
    return ts


def cpu_sample_size(cfg, budget_s, steps):
    # survey-time throughput of the reference on 8 vCPU, images/s
    rate = {"cfg1": 1250.0, "cfg2": 6.9, "cfg3": 110.0, "cfg4": 4.8, "cfg5": 2.75}[cfg["_id"]]
    n = int(max(2, min(cfg["N"], budget_s * rate / max(steps, 1))))
    return n


def run_reference(args, cfg):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = cpu_sample_size(cfg, 100.0, args.steps + args.warmup)
    ts = time_cpu(cfg, n, args.steps, args.warmup, 1000 + int(cfg["_id"][3:]))
    total = sum(ts)
    value = n * len(ts) / total
    sample = "{} of {} images per step, float64, oracle port of im2col+GEMM+np.add.at".format(n, cfg["N"])
    line = {"impl": "reference", "metric": "Conv2D fwd+bwd images/s", "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "mode": "reference-cpu", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": host_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
#  clocks sampling
# ---------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
#  GPU arm
# ---------------------------------------------------------------------------------------------------
ELEMENTWISE = ("npml_bn2d_fwd", "npml_bn2d_bwd", "npml_add_act_fwd", "npml_dz_prep")


def measured_traffic(cfg_id, mode, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get("{}/{}/{}".format(cfg_id, mode, kernel))
    except Exception:
        return None


def run_gpu(args, cfg):
    import torch
    pkg = importlib.import_module("numpy-ml_b200")
    L, D = pkg._lib, pkg.dist
    st = D.init()
    world, rank = st["world"], st["rank"]
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    mode = args.mode
    n = cfg["N"]
    kind = cfg["kind"]
    seed = 1000 + int(cfg["_id"][3:])
    X, dY, W, b = make_inputs(cfg, n, seed + 17 * rank)
    _, _, W, b = make_inputs(cfg, 2, seed)          # identical parameters on every rank
    act_dt = L.torch_dtype(mode)
    esize = 2 if mode == "bf16" else 4

    if kind == "conv":
        model = pkg.Conv2D(cfg["K"], cfg["k"], pad=cfg["pad"], stride=cfg["stride"], dilation=cfg["dilation"], mode=mode)
    elif kind == "deconv":
        model = pkg.Deconv2D(cfg["K"], cfg["k"], pad=cfg["pad"], stride=cfg["stride"], mode=mode)
    else:
        model = pkg.SkipConnectionConvModule(out_ch1=cfg["K"], out_ch2=cfg["K"], kernel_shape1=(3, 3), kernel_shape2=(3, 3),
                                             kernel_shape_skip=(1, 1), pad1=1, pad2=1, act_fn="relu", mode=mode,
                                             sync_bn=(world > 1))
    Xh = torch.from_numpy(X).pin_memory()
    dYh = torch.from_numpy(dY).pin_memory()
    Xd = L.to_device(Xh.to(dev), act_dt)
    dYd = L.to_device(dYh.to(dev), act_dt)
    model.forward(Xd, retain_derived=False)
    if kind == "skip":
        layers = [model.conv1, model.conv2, model.conv_skip, model.batchnorm1, model.batchnorm2, model.batchnorm_skip]
        for t, conv, bn in (("1", model.conv1, model.batchnorm1), ("2", model.conv2, model.batchnorm2),
                            ("s", model.conv_skip, model.batchnorm_skip)):
            conv.parameters["W"], conv.parameters["b"] = (L.to_device(W["W" + t], torch.float32),
                                                          L.to_device(W["b" + t], torch.float32))
            bn.parameters["scaler"], bn.parameters["intercept"] = (L.to_device(W["g" + t], torch.float32),
                                                                   L.to_device(W["h" + t], torch.float32))
        conv_layers = layers[:3]
    else:
        layers = conv_layers = [model]
        model.parameters["W"], model.parameters["b"] = L.to_device(W, torch.float32), L.to_device(b, torch.float32)
    # dW / db of a layer are all-reduced on a side stream while its dX is still being computed
    bucket = D.GradBucket(layers, overlap=True)

    def step():
        model.flush_gradients()
        Y = model.forward(Xd)
        dX = model.backward(dYd)
        if world > 1:
            bucket.allreduce()
        return Y, dX

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- CUDA graph of one step (launch-bound configs: the host needs ~0.18 ms to enqueue a step) -------------
    # forward + backward (+ all-reduce) are captured ONCE through the public layer API; the timed region replays the
    # graph.  The per-call CUDA events are captured as external event-record nodes, so the per-call durations are
    # still measured inside the timed region (those of the last replay).
    graph = None
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            L.PROFILE, L.PROFILE_EXTERNAL = [], True
            L.launch_count(reset=True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            graph, graph_prof, graph_launches = g, L.PROFILE, L.launch_count()
        except Exception as e:      # capture not possible (e.g. torch without external events): eager timed region
            if rank == 0:
                print("bench: CUDA graph capture failed ({}); timing eagerly".format(str(e).splitlines()[0]), file=sys.stderr)
            graph = None
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
        L.PROFILE, L.PROFILE_EXTERNAL = None, False
        if graph is not None:
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, per-call events inside ----------
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    # pre-roll: keep the GPU under the same load for ~0.4 s so that nvidia-smi (50 ms period) samples the clocks of
    # a loaded device; these steps are extra warm-up, not part of the timed region.  The number of steps is derived
    # from a max-over-ranks estimate so that every rank issues the same number of collectives.
    def one():
        if graph is not None:
            graph.replay()
        else:
            step()

    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(5):
        one()
    p1.record()
    torch.cuda.synchronize()
    est_ms = max(D.all_reduce_max(p0.elapsed_time(p1)) / 5.0, 1e-3)
    for _ in range(int(min(5000, max(10, 400.0 / est_ms)))):
        one()
    torch.cuda.synchronize()
    if graph is None:
        L.PROFILE = []
    L.launch_count(reset=True)
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            step()
    host_ms = 1e3 * (time.perf_counter() - t_host) / args.steps      # time the host needs to ENQUEUE one step
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    if graph is not None:
        launches = graph_launches * args.steps
        prof, prof_steps = graph_prof, 1
    else:
        launches = L.launch_count()
        prof, L.PROFILE = L.PROFILE, None
        prof_steps = args.steps
    clocks = sampler.stop()
    ms_total = D.all_reduce_max(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- per-call shares and the roofline of the dominant call ----------------------------------------
    per, count = {}, {}
    for name, a, z in prof:
        per[name] = per.get(name, 0.0) + a.elapsed_time(z)
        count[name] = count.get(name, 0) + 1
    step_ms = {k: v / prof_steps for k, v in per.items()}          # time per step spent in each entry point
    call_ms = {k: per[k] / count[k] for k in per}                   # mean duration of one call
    fl = flops_per_pass(cfg, n)
    pk = peaks()
    dominant = max(step_ms, key=step_ms.get)
    tc = all(bool(l._uses_tc()) for l in conv_layers)
    calls_per_step = count[dominant] / prof_steps
    if dominant in ELEMENTWISE:
        oh, ow = out_hw(cfg)
        elems = n * oh * ow * cfg["K"]
        passes = {"npml_bn2d_fwd": 2, "npml_bn2d_bwd": 3, "npml_add_act_fwd": 3, "npml_dz_prep": 3}[dominant]
        byt = float(elems * esize * passes)
        achieved = byt / (call_ms[dominant] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "peak_source": pk["source"]}
    elif tc:
        peak = pk["bf16_tflops_sustained"] if mode == "bf16" else pk["bf16_tflops_sustained"] / 2.0
        achieved = fl / calls_per_step / (call_ms[dominant] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak,
                "peak_source": pk["source"] + (" bf16 sustained" if mode == "bf16" else " bf16 sustained / 2 (tf32)")}
    else:
        byt = min_bytes_per_pass(cfg, n, esize)
        achieved = byt / (call_ms[dominant] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "peak_source": pk["source"]}
    roof["traffic"] = measured_traffic(cfg["_id"], mode, dominant)
    roof["kernel_ms"] = {k: round(v, 4) for k, v in call_ms.items()}
    roof["step_ms_by_call"] = {k: round(v, 4) for k, v in step_ms.items()}
    roof["share_of_step"] = round(step_ms[dominant] / ms_step, 3)
    tflops_step = 3.0 * fl * world / (ms_step * 1e-3) / 1e12

    # ---- end to end through the public API with HOST buffers ----------------------------------------------
    # X and dLdy start in pinned host memory (fp32); Y, dX, dW, db end in pinned host memory.  Host->device and
    # device->host copies run on their own streams so the two PCIe directions overlap; Conv2D / Deconv2D batches
    # are pipelined in chunks (forward / backward of chunk i while chunk i+1 is in flight).  The module of cfg4
    # needs whole-batch BatchNorm statistics, so it is not chunked.
    e2e_steps = max(1, min(args.steps, 5))
    oh, ow = out_hw(cfg)
    Yh = torch.empty((n, oh, ow, cfg["K"]), dtype=act_dt).pin_memory()
    dXh = torch.empty(X.shape, dtype=act_dt).pin_memory()
    gradh = torch.empty(bucket.flat.numel(), dtype=torch.float32).pin_memory()
    chunks = 1 if kind == "skip" else 4
    bounds = [(i * n // chunks, (i + 1) * n // chunks) for i in range(chunks)]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    xs32 = [torch.empty((hi - lo,) + X.shape[1:], dtype=torch.float32, device=dev) for lo, hi in bounds]
    dys32 = [torch.empty((hi - lo, oh, ow, cfg["K"]), dtype=torch.float32, device=dev) for lo, hi in bounds]

    def e2e_step():
        cur = torch.cuda.current_stream()
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        ev_x, ev_dy = [], []
        with torch.cuda.stream(s_in):
            for i, (lo, hi) in enumerate(bounds):
                xs32[i].copy_(Xh[lo:hi], non_blocking=True)
                ev_x.append(torch.cuda.Event()); ev_x[-1].record()
            for i, (lo, hi) in enumerate(bounds):
                dys32[i].copy_(dYh[lo:hi], non_blocking=True)
                ev_dy.append(torch.cuda.Event()); ev_dy[-1].record()
        model.flush_gradients()
        outs = []
        for i, (lo, hi) in enumerate(bounds):
            cur.wait_event(ev_x[i])
            Y = model.forward(L.cast(xs32[i], act_dt))
            outs.append(Y)
            ev = torch.cuda.Event(); ev.record()
            s_out.wait_event(ev)
            with torch.cuda.stream(s_out):
                Yh[lo:hi].copy_(Y, non_blocking=True)
        for i, (lo, hi) in enumerate(bounds):
            cur.wait_event(ev_dy[i])
            dyd = L.cast(dys32[i], act_dt)
            dX = model.backward(dyd) if chunks == 1 else model.backward(dyd, index=i)
            outs.append(dX)
            ev = torch.cuda.Event(); ev.record()
            s_out.wait_event(ev)
            with torch.cuda.stream(s_out):
                dXh[lo:hi].copy_(dX, non_blocking=True)
        if world > 1:
            bucket.allreduce()
        ev = torch.cuda.Event(); ev.record()
        s_out.wait_event(ev)
        with torch.cuda.stream(s_out):
            gradh.copy_(bucket.flat, non_blocking=True)
        cur.wait_stream(s_out)
        return outs

    e2e_step()
    torch.cuda.synchronize()
    D.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    torch.cuda.synchronize()
    D.barrier()
    e2e_ms = D.all_reduce_max(f0.elapsed_time(f1))
    e2e = {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
           "h2d_bytes_per_step": int(Xh.numel() * 4 + dYh.numel() * 4),
           "d2h_bytes_per_step": int(Yh.numel() * esize + dXh.numel() * esize + gradh.numel() * 4),
           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "chunks": chunks,
           "host_dtype_in": "f32", "host_dtype_out": "bf16" if mode == "bf16" else "f32"}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ns = cpu_sample_size(cfg, 20.0, 1)
        ts = time_cpu(cfg, ns, 1, 0, seed)
        cpu = {"value": ns / ts[0], "unit": "images/s", "cores": host_threads(), "kind": "port",
               "sample": "{} of {} images, one fwd+bwd, float64, oracle port of im2col+GEMM+np.add.at".format(ns, n)}

    if rank == 0:
        line = {"metric": "Conv2D fwd+bwd images/s", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[mode],
                "data": "synthetic",
                "config": {"workload": cfg["name"], "mode": mode, "per_gpu_batch": n, "global_batch": n * world,
                           "parallelism": "dp{}".format(world) + ("+syncbn" if (kind == "skip" and world > 1) else ""),
                           "tensor_cores": tc,
                           "l2": "inputs X + dLdy are {:.0f} MB per step (L2 is 126 MB); no explicit flush".format(
                               (Xd.numel() + dYd.numel()) * esize / 1e6)},
                "host_enqueue_ms_per_step": host_ms, "cuda_graph": graph is not None, "tflops": tflops_step, "frac_of_bf16_peak": tflops_step / (world * pk["bf16_tflops_sustained"]),
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    # teardown must never hang the job: a captured graph that contains NCCL kernels has to be destroyed before the
    # communicator, and a watchdog forces the exit if the clean shutdown still blocks
    watchdog = threading.Timer(30.0, lambda: os._exit(0))
    watchdog.daemon = True
    watchdog.start()
    sys.stdout.flush()
    if graph is not None:
        graph_prof = None
        graph.reset()
        graph = None
    torch.cuda.synchronize()
    D.barrier()
    D.destroy()
    watchdog.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="bf16", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of CUDA-graph replays")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    cfg["_id"] = args.config
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
