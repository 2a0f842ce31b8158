"""Ad-hoc GPU check of the wgrad kernels against a torch fp64 reference (debug aid, not a test)."""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("numpy-ml_b200")
L, U = pkg._lib, pkg.utils
lib = L.load()
torch.manual_seed(0)

def run(mode, N, H, W, C, K, k, pad, s, D):
    dt = L.torch_dtype(mode)
    X = torch.randn(N, H, W, C, device="cuda").to(dt)
    fr, fc = k
    p4 = (pad,) * 4
    oh, ow = U.conv_out_hw(H, W, fr, fc, p4, s, D - 1)
    dZ = torch.randn(N, oh, ow, K, device="cuda").to(dt)
    d = U.make_desc(N, H, W, C, K, fr, fc, s, D, p4, oh, ow, mode)
    dW = torch.zeros(fr, fc, C, K, device="cuda")
    ws = L.workspace(lib.npml_workspace_bytes(d, L.OP_CONV_WGRAD))
    ws.zero_()
    tc = lib.npml_uses_tensor_cores(d, L.OP_CONV_WGRAD)
    L.check(lib.npml_conv2d_wgrad(d, L.ptr(X), L.ptr(dZ), L.ptr(dW), L.ptr(ws), ws.numel(), L.stream()), "wgrad")
    torch.cuda.synchronize()
    # reference: dW[r,t,c,k] = sum X[n, oh*s + r*D - p, ow*s + t*D - p, c] dZ[n,oh,ow,k]
    Xp = torch.nn.functional.pad(X.double(), (0, 0, pad, pad, pad, pad))
    ref = torch.zeros(fr, fc, C, K, device="cuda", dtype=torch.float64)
    for r in range(fr):
        for t in range(fc):
            win = Xp[:, r * D: r * D + (oh - 1) * s + 1: s, t * D: t * D + (ow - 1) * s + 1: s, :]
            ref[r, t] = torch.einsum("nhwc,nhwk->ck", win, dZ.double())
    err = float((dW.double() - ref).norm() / ref.norm())
    nz = float((dW == 0).float().mean())
    print("mode=%s tc=%d shape N%d %dx%dx%d->%d k%s s%d D%d: rel err %.3e  zeros %.3f  |dW| %.3e |ref| %.3e" % (
        mode, tc, N, H, W, C, K, k, s, D, err, nz, float(dW.norm()), float(ref.norm())))
    if err > 1e-2:
        # where is it wrong?  per-tap error, and try the transposed hypothesis
        for r in range(fr):
            print("   tap row", r, ["%.2e" % float((dW[r, t].double() - ref[r, t]).norm() / ref[r, t].norm()) for t in range(fc)])
        part = ws[: 4 * fr * fc * C * K].view(torch.float32)[: fr * fc * C * K].view(fr, fc, C, K)
        print("   split-0 partial norm %.3e, nonzero frac %.3f, nan %d" % (float(part.norm()), float((part != 0).float().mean()), int(torch.isnan(part).sum())))

for mode in ("bf16", "tf32"):
    run(mode, 4, 16, 16, 64, 64, (3, 3), 1, 1, 1)
    run(mode, 2, 16, 16, 128, 128, (1, 1), 0, 1, 1)
    run(mode, 3, 28, 28, 128, 256, (3, 3), 0, 2, 3)
    run(mode, 2, 9, 11, 72, 40, (3, 3), 1, 1, 1)
