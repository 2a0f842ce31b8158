"""CPU check of the index logic behind csrc/tc_ws_conv.cu (the weight-stationary convolution): scripts/ws_emulate.py
mirrors plan_ws (tap offsets on the flat padded slot grid, pairing of taps at an even slot distance, the shifted b half)
and the kernel's tile loop in NumPy float64; the decomposition must reproduce the oracle's Conv2D forward (form 0) and
dX (form 1) exactly.  The GPU parity of the kernel itself is tests/test_gpu_parity.py::test_weight_stationary_kernel_optin."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import ws_emulate as E  # noqa: E402

CASES = [
    ((2, 7, 6, 8), 16, (3, 3), "same", 0),      # pair mode: (0, 2) + (1) per row
    ((2, 6, 6, 8), 72, (3, 3), 1, 0),           # single mode
    ((1, 9, 8, 8), 8, (4, 4), 1, 0),            # (0, 2) + (1, 3): every tap paired
    ((2, 6, 7, 8), 8, (2, 3), (1, 2), 0),       # non-square, asymmetric pads
    ((1, 9, 9, 8), 8, (3, 3), "same", 1),       # D = 2: neighbours pair, delta = 2
    ((1, 11, 11, 8), 8, (3, 3), "same", 2),     # D = 3: delta = 6
    ((2, 5, 5, 8), 8, (1, 1), 0, 0),
    ((1, 8, 8, 8), 8, (5, 5), 2, 0),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_ws_decomposition_matches_oracle(case):
    xs, oc, ks, pad, d = CASES[case]
    E.check(xs, oc, ks, pad, d, np.random.default_rng(case))


def test_pairs_are_at_even_slot_distance():
    """tcgen05.ld.16x256b needs an even column offset (scripts/ts_probe2.cu): delta = D or 2D, never odd, at most 8."""
    for D in (1, 2, 3, 4):
        for form in (0, 1):
            P = E.plan(1, 12, 12, 8, 12, 12, 8, 3, 3, D, D, D, form)
            assert P["delta"] % 2 == 0 and P["delta"] <= 8
            offs = {}
            for off, ia, ib in P["groups"]:
                offs[ia] = off
            for off, ia, ib in P["groups"]:
                if ib >= 0:
                    # the partner's own offset is delta slots further right
                    dyb, dxb = divmod(ib, 3)
                    dya, dxa = divmod(ia, 3)
                    assert dya == dyb and abs(dxa - dxb) * D == P["delta"]
            taps = sorted([g[1] for g in P["groups"]] + [g[2] for g in P["groups"] if g[2] >= 0])
            assert taps == list(range(9))          # every tap exactly once


# ---------------------------------------------------------------------------------------------------
#  the C++ planners themselves (npml_debug_ws_plan: host only, no device) against the NumPy mirror above
# ---------------------------------------------------------------------------------------------------
def _cpp_plan(pkg, monkeypatch, variant, N, H, W, C, K, ks, pad, dil, op):
    import ctypes as C_
    from oracle import conv_oracle as O
    lib = pkg._lib.load()
    monkeypatch.setenv("NPML_WS", str(variant))
    p4 = O.resolve_pad((N, H, W, C), pad, ks, 1, dil)
    _, oh, ow, _ = O.calc_conv_out_dims((N, H, W, C), ks + (C, K), 1, p4, dil)
    d = pkg.utils.make_desc(N, H, W, C, K, ks[0], ks[1], 1, dil + 1, p4, oh, ow, "bf16")
    out = (C_.c_int32 * 256)()
    n = lib.npml_debug_ws_plan(d, op, variant, C_.cast(out, C_.c_void_p), 256)
    assert n >= 16
    return list(out[:n]), p4, oh, ow


PLAN_SHAPES = [
    # N, H, W, C, K, kernel, pad, dilation
    (256, 56, 56, 64, 64, (3, 3), "same", 0),       # cfg2
    (256, 32, 32, 64, 128, (3, 3), 1, 0),           # cfg4 conv1
    (8, 20, 24, 32, 48, (4, 4), 1, 0),
    (4, 17, 19, 16, 16, (5, 5), 2, 0),
    (4, 30, 30, 64, 64, (3, 3), "same", 1),         # D = 2
    (4, 30, 30, 64, 64, (3, 3), "same", 2),         # D = 3 -> delta 6
    (2, 12, 12, 64, 128, (1, 1), 0, 0),
]


@pytest.mark.parametrize("shape", range(len(PLAN_SHAPES)))
@pytest.mark.parametrize("op", [0, 1])               # NPML_OP_CONV_FPROP, NPML_OP_CONV_DGRAD
def test_cpp_planner_matches_numpy_mirror(monkeypatch, shape, op):
    from conftest import npml
    pkg = npml()
    N, H, W, C, K, ks, pad, dil = PLAN_SHAPES[shape]
    got, p4, oh, ow = _cpp_plan(pkg, monkeypatch, 1, N, H, W, C, K, ks, pad, dil, op)
    D = dil + 1
    if op == 0:     # forward: gathers X (N, H, W, C) into (N, OH, OW, K)
        ref = E.plan(N, H, W, C, oh, ow, K, ks[0], ks[1], D, p4[0], p4[2], 0)
        co, ci = K, C
    else:           # dX: gathers dZ (N, OH, OW, K) into (N, H, W, C)
        ref = E.plan(N, oh, ow, K, H, W, C, ks[0], ks[1], D, p4[0], p4[2], 1)
        co, ci = C, K
    nkb = (ci + 63) // 64
    ngroups_ref = len(ref["groups"])
    if ngroups_ref * nkb * 32 + 2 * (64 + (8 if ref["pair"] else 0)) > 512:
        assert got[0] == 0                          # the weights do not fit in tensor memory next to two accumulators
        return
    assert got[0] == 1
    sup, pair, delta, NT, NC, n_acc, UT, SR, Hp, Wp, pady, padx, ngroups, wcols, tiles, smem = got[:16]
    assert (Hp, Wp, pady, padx) == (ref["Hp"], ref["Wp"], ref["pady"], ref["padx"])
    assert bool(pair) == ref["pair"] and (not pair or delta == ref["delta"])
    groups = [tuple(got[16 + 3 * i:19 + 3 * i]) for i in range(ngroups)]
    assert groups == [tuple(g) for g in ref["groups"]]
    # resources: TMEM columns, shared memory, the tile covers what the epilogue walks, every slot is in some tile
    assert wcols == ngroups * nkb * 32 and wcols + n_acc * NC <= 512 and n_acc >= 2
    assert NT % 32 == 0 and NC == NT + (8 if pair else 0) and 1 <= UT <= n_acc
    assert smem <= 227 * 1024
    assert tiles * NT >= N * Hp * Wp
    halo = max(g[0] for g in groups)
    assert SR * Wp >= (Wp - 1) + UT * NT + (NC - NT) + halo       # the slab holds every slot a unit's MMAs read
