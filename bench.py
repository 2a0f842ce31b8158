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
    np.random.seed(seed)
    W = O.glorot_uniform(cfg["k"] + (cfg["C"], cfg["K"])).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, 1, 1, cfg["K"]))).astype(np.float32)
    return X, dY, W, b


def cpu_step(cfg, X, dY, W, b):
    """forward + backward of the reference algorithm in float64 (oracle port)."""
    from oracle import conv_oracle as O
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
    rate = {"cfg1": 1250.0, "cfg2": 6.9, "cfg3": 110.0, "cfg5": 2.75}[cfg["_id"]]
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
    seed = 1000 + int(cfg["_id"][3:])
    X, dY, W, b = make_inputs(cfg, n, seed + 17 * rank)
    _, _, W, b = make_inputs(cfg, 2, seed)          # identical parameters on every rank
    act_dt = L.torch_dtype(mode)
    esize = 2 if mode == "bf16" else 4

    if cfg["kind"] == "conv":
        layer = pkg.Conv2D(cfg["K"], cfg["k"], pad=cfg["pad"], stride=cfg["stride"], dilation=cfg["dilation"], mode=mode)
    else:
        layer = pkg.Deconv2D(cfg["K"], cfg["k"], pad=cfg["pad"], stride=cfg["stride"], mode=mode)
    Xh = torch.from_numpy(X).pin_memory()
    dYh = torch.from_numpy(dY).pin_memory()
    Xd = L.to_device(Xh.to(dev), act_dt)
    dYd = L.to_device(dYh.to(dev), act_dt)
    layer.forward(Xd, retain_derived=False)
    layer.parameters["W"], layer.parameters["b"] = L.to_device(W, torch.float32), L.to_device(b, torch.float32)
    bucket = D.GradBucket([layer])

    def step():
        layer.flush_gradients()
        Y = layer.forward(Xd)
        dX = layer.backward(dYd)
        if world > 1:
            bucket.allreduce()
        return Y, dX

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, per-kernel events inside -------
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    L.PROFILE = []
    L.launch_count(reset=True)
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    launches = L.launch_count()
    prof, L.PROFILE = L.PROFILE, None
    clocks = sampler.stop()
    ms_total = D.all_reduce_max(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- per-kernel shares and the roofline of the dominant kernel -----------------------------------
    per = {}
    for name, a, z in prof:
        per.setdefault(name, []).append(a.elapsed_time(z))
    kern = {k: float(np.mean(v)) for k, v in per.items()}
    fl = flops_per_pass(cfg, n)
    pk = peaks()
    dominant = max(kern, key=kern.get)
    tc = bool(layer._uses_tc())
    if tc:
        peak = pk["bf16_tflops_sustained"] if mode == "bf16" else pk["bf16_tflops_sustained"] / 2.0
        achieved = fl / (kern[dominant] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": pk["source"] + (" bf16 sustained" if mode == "bf16" else " bf16 sustained / 2 (tf32)")}
    else:
        byt = min_bytes_per_pass(cfg, n, esize)
        achieved = byt / (kern[dominant] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"]}
    roof["kernel_ms"] = {k: round(v, 4) for k, v in kern.items()}
    roof["share_of_step"] = round(kern[dominant] / ms_step, 3)
    tflops_step = 3.0 * fl / (ms_step * 1e-3) / 1e12

    # ---- end to end through the public API with host buffers -----------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    Yh = torch.empty((n,) + tuple(out_hw(cfg)) + (cfg["K"],), dtype=act_dt).pin_memory()
    dXh = torch.empty(X.shape, dtype=act_dt).pin_memory()
    dWh = torch.empty(W.shape, dtype=torch.float32).pin_memory()
    dbh = torch.empty(b.shape, dtype=torch.float32).pin_memory()

    def e2e_step():
        xd = L.cast(Xh.to(dev, non_blocking=True), act_dt)
        dyd = L.cast(dYh.to(dev, non_blocking=True), act_dt)
        layer.flush_gradients()
        Y = layer.forward(xd)
        dX = layer.backward(dyd)
        if world > 1:
            bucket.allreduce()
        Yh.copy_(Y, non_blocking=True)
        dXh.copy_(dX, non_blocking=True)
        dWh.copy_(layer.gradients["W"], non_blocking=True)
        dbh.copy_(layer.gradients["b"], non_blocking=True)

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
           "d2h_bytes_per_step": int(Yh.numel() * esize + dXh.numel() * esize + dWh.numel() * 4 + dbh.numel() * 4),
           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps}

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
                           "parallelism": "dp{}".format(world), "tensor_cores": tc,
                           "l2": "inputs X + dLdy are {:.0f} MB per step (L2 is 126 MB); no explicit flush".format(
                               (Xd.numel() + dYd.numel()) * esize / 1e6)},
                "tflops": tflops_step, "frac_of_bf16_peak": tflops_step / pk["bf16_tflops_sustained"],
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    D.destroy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="bf16", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    cfg["_id"] = args.config
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_gpu(args, cfg)


if __name__ == "__main__":
    main()
