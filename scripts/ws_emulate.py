"""CPU emulation of the index logic of csrc/tc_ws_conv.cu (plan_ws + the kernel's tile / pair / shift bookkeeping) in
NumPy float64, checked against the oracle: the weight-stationary decomposition (taps paired along M, bottom half shifted
by delta slots, flat padded slot grid) reproduces Conv2D forward (form 0) and Conv2D dX (form 1) exactly.
Test infrastructure only (imports oracle/)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import conv_oracle as O


def plan(N, IH, IW, CI, OH, OW, CO, fr, fc, D, pr, pc, form):
    dy, dx = {}, {}
    for r in range(fr):
        for t in range(fc):
            i = r * fc + t
            dy[i] = r * D - pr if form == 0 else pr - r * D
            dx[i] = t * D - pc if form == 0 else pc - t * D
    lo_y, hi_y = min(dy.values()), max(dy.values())
    lo_x, hi_x = min(dx.values()), max(dx.values())
    Hp, Wp = OH + hi_y - lo_y, OW + hi_x - lo_x
    pair = CO <= 64
    groups = []
    dist = 1 if D % 2 == 0 else 2
    if pair:
        for r in range(fr):
            used = [False] * fc
            for k in range(fc):
                if used[k]:
                    continue
                ia = r * fc + (k if form == 0 else fc - 1 - k)
                ib = -1
                if k + dist < fc and not used[k + dist]:
                    ib = r * fc + (k + dist if form == 0 else fc - 1 - (k + dist))
                    used[k + dist] = True
                used[k] = True
                off = (dy[ia] - lo_y) * Wp + (dx[ia] - lo_x)
                groups.append((off, ia, ib))
    else:
        for i in range(fr * fc):
            groups.append(((dy[i] - lo_y) * Wp + (dx[i] - lo_x), i, -1))
    return dict(Hp=Hp, Wp=Wp, pady=-lo_y, padx=-lo_x, pair=pair, delta=dist * D, groups=groups)


def emulate(inp, Wt, OH, OW, fr, fc, D, pr, pc, form, NT=32):
    """inp (N, IH, IW, CI); Wt[tap][co][ci]; returns out (N, OH, OW, CO)."""
    N, IH, IW, CI = inp.shape
    CO = Wt.shape[1]
    P = plan(N, IH, IW, CI, OH, OW, CO, fr, fc, D, pr, pc, form)
    Hp, Wp = P["Hp"], P["Wp"]
    slots = N * Hp * Wp
    pad = 8 if P["pair"] else 0
    halo = max(g[0] for g in P["groups"])
    # the flat slab: slot (n, yp, xp) holds input pixel (yp - pady, xp - padx), zero outside (TMA out-of-bounds fill)
    slab = np.zeros((slots + NT + pad + halo + 8, CI))
    for n in range(N):
        for yp in range(Hp):
            iy = yp - P["pady"]
            if not 0 <= iy < IH:
                continue
            for xp in range(Wp):
                ix = xp - P["padx"]
                if 0 <= ix < IW:
                    slab[(n * Hp + yp) * Wp + xp] = inp[n, iy, ix]
    out = np.zeros((N, OH, OW, CO))
    for s0 in range(0, slots, NT):
        NC = NT + pad
        top = np.zeros((CO, NC))
        bot = np.zeros((CO, NC))
        for off, ia, ib in P["groups"]:
            win = slab[s0 + off:s0 + off + NC]                 # (NC, CI): the B operand of the group
            top += Wt[ia] @ win.T
            if ib >= 0:
                bot += Wt[ib] @ win.T
        for q in range(NT):
            s = s0 + q
            n, rem = divmod(s, Hp * Wp)
            yp, xp = divmod(rem, Wp)
            if n < N and yp < OH and xp < OW:
                out[n, yp, xp] = top[:, q] + (bot[:, q + P["delta"]] if P["pair"] else 0.0)
    return out


def check(xs, oc, ks, pad, d, rng, rows_per_tile=0):
    X = rng.standard_normal(xs)
    W = rng.standard_normal(ks + (xs[3], oc))
    Z = O.conv2D(X, W, 1, pad, d)
    p4 = O.resolve_pad(X.shape, pad, ks, 1, d)
    D = d + 1
    Wt = W.reshape(ks[0] * ks[1], xs[3], oc).transpose(0, 2, 1)           # [tap][co][ci], forward
    kw = {}
    if rows_per_tile:              # the NPML_WS=2 tiling: a tile is a whole number of padded rows (NT = R * Wp)
        kw["NT"] = rows_per_tile * plan(xs[0], xs[1], xs[2], xs[3], Z.shape[1], Z.shape[2], oc, ks[0], ks[1], D, p4[0],
                                        p4[2], 0)["Wp"]
    got = emulate(X, Wt, Z.shape[1], Z.shape[2], ks[0], ks[1], D, p4[0], p4[2], 0, **kw)
    e1 = np.abs(got - Z).max()
    # dX = transposed form: gathers dZ with W[r, t, c, k] read as [tap][co = c][ci = k]
    dZ = rng.standard_normal(Z.shape)
    dXr, _, _ = O.conv2d_layer_bwd(dZ, X, Z, W, 1, pad, d)
    Wt2 = W.reshape(ks[0] * ks[1], xs[3], oc)                             # [tap][c][k]
    if rows_per_tile:
        kw["NT"] = rows_per_tile * plan(xs[0], Z.shape[1], Z.shape[2], oc, xs[1], xs[2], xs[3], ks[0], ks[1], D, p4[0],
                                        p4[2], 1)["Wp"]
    got2 = emulate(dZ, Wt2, xs[1], xs[2], ks[0], ks[1], D, p4[0], p4[2], 1, **kw)
    e2 = np.abs(got2 - dXr).max()
    print(xs, oc, ks, pad, d, "fwd err %.2e  dX err %.2e" % (e1, e2))
    assert e1 < 1e-9 and e2 < 1e-9


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    check((2, 7, 6, 8), 16, (3, 3), "same", 0, rng)       # pair mode, cfg2 miniature
    check((2, 6, 6, 8), 72, (3, 3), 1, 0, rng)            # single mode (out_ch > 64)
    check((1, 9, 8, 8), 8, (4, 4), 1, 0, rng)             # even kernel: all taps paired
    check((2, 6, 7, 8), 8, (2, 3), (1, 2), 0, rng)        # non-square, asymmetric pads
    check((1, 9, 9, 8), 8, (3, 3), "same", 1, rng)        # dilation 2 (delta = 2)
    check((2, 5, 5, 8), 8, (1, 1), 0, 0, rng)             # 1x1
    check((1, 8, 8, 8), 8, (5, 5), 2, 0, rng)             # 25 taps: 15 groups
    print("ok")
