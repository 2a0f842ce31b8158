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


@pytest.mark.parametrize("rows", [1, 2, 3])
def test_ws_row_aligned_tiles(rows):
    """The tiling of the experimental NPML_WS=2 variant (csrc/tc_ws2_conv.cu): NT = rows * Wp, not a multiple of 32."""
    E.check((2, 7, 6, 8), 16, (3, 3), "same", 0, np.random.default_rng(40 + rows), rows_per_tile=rows)
    E.check((2, 6, 6, 8), 72, (3, 3), 1, 0, np.random.default_rng(50 + rows), rows_per_tile=rows)


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
