"""CPU check of the lane layout of the warp-per-element C3D10 stiffness kernel (csrc/elem_solid.cu: c3d10_K_warp_kernel).

The kernel's decomposition is restated lane by lane in numpy -- phase A lane (q, h): point q, nodes 0..5 / 6..9, gradient rows of 36
doubles per point; phase B lane = tile (t, b): node pairs (2t, b), (2t+1, b), b >= 2t; tile writes: 3x3 blocks, the mirror image as
6-wide rows for b > 2t, diagonal blocks symmetrised from one triangle -- with the derivative tables taken from the library's own
host entry points (femb_default_points / femb_shape_tables, no GPU involved), and compared with the oracle (reference
element.py:1191-1239).  It pins what the GPU test cannot show directly: every one of the 900 tile entries is written, an entry
written twice gets the same value both times (no race between lanes), and the result is bitwise symmetric."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GQ = 36   # C10W_GQ


def _tables():
    from femb200 import _lib, ops
    L = _lib.lib
    dp = ctypes.POINTER(ctypes.c_double)
    pts = np.zeros(64 * 4)
    nq = L.femb_default_points(ops.C3D10, pts.ctypes.data_as(dp))
    assert nq == 11
    N, dN = np.zeros(nq * 10), np.zeros(nq * 30)
    assert L.femb_shape_tables(ops.C3D10, pts.ctypes.data_as(dp), nq, N.ctypes.data_as(dp), dN.ctypes.data_as(dp)) == 0
    return nq, dN.reshape(nq, 10, 3), pts.reshape(-1, 4)[:nq, 3].copy()


def _inv3(m):
    c00, c01, c02 = m[4] * m[8] - m[5] * m[7], m[5] * m[6] - m[3] * m[8], m[3] * m[7] - m[4] * m[6]
    det = m[0] * c00 + m[1] * c01 + m[2] * c02
    i = 1.0 / det
    return det, [c00 * i, (m[2] * m[7] - m[1] * m[8]) * i, (m[1] * m[5] - m[2] * m[4]) * i, c01 * i, (m[0] * m[8] - m[2] * m[6]) * i,
                 (m[2] * m[3] - m[0] * m[5]) * i, c02 * i, (m[1] * m[6] - m[0] * m[7]) * i, (m[0] * m[4] - m[1] * m[3]) * i]


def _tile_of(lane):
    if lane < 10:
        return 0, lane
    if lane < 18:
        return 1, lane - 8
    if lane < 24:
        return 2, lane - 14
    if lane < 28:
        return 3, lane - 18
    return 4, lane - 20


def _warp(x, lam, mu, nq, dN, wts):
    xs = x.reshape(-1)
    gs, wd = np.zeros(16 * GQ), np.zeros(16)
    for lane in range(32):                      # phase A
        q, h = lane >> 1, lane & 1
        if q >= nq:
            continue
        J = [sum(dN[q, a, i] * xs[3 * a + k] for a in range(10)) for i in range(3) for k in range(3)]
        det, Ji = _inv3(J)
        if h == 0:
            wd[q] = det * wts[q]
        for aa in range(6 if h == 0 else 4):
            d = dN[q, aa + 6] if h else dN[q, aa]
            for c in range(3):
                gs[q * GQ + h * 18 + 3 * aa + c] = Ji[3 * c] * d[0] + Ji[3 * c + 1] * d[1] + Ji[3 * c + 2] * d[2]
    kt, cnt = np.full(900, np.nan), np.zeros(900, int)

    def put(idx, v):
        assert cnt[idx] == 0 or kt[idx] == v, f"entry {idx} rewritten with a different value"
        kt[idx], cnt[idx] = v, cnt[idx] + 1

    seen = set()
    for lane in range(30):                      # phase B
        t, b = _tile_of(lane)
        assert 2 * t <= b <= 9 and (t, b) not in seen
        seen.add((t, b))
        a0, a1 = 2 * t, 2 * t + 1
        S = np.zeros((2, 3, 3))
        for p in range(nq):
            g = gs[p * GQ:p * GQ + 30]
            for blk in range(2):
                S[blk] += np.outer(wd[p] * g[6 * t + 3 * blk:6 * t + 3 * blk + 3], g[3 * b:3 * b + 3])
        k = [lam * S[blk] + mu * S[blk].T + mu * np.trace(S[blk]) * np.eye(3) for blk in range(2)]
        for blk, a in ((0, a0), (1, a1)):
            if b == a:                          # diagonal block: the upper triangle decides
                k[blk] = np.triu(k[blk]) + np.triu(k[blk], 1).T
        for i in range(3):
            for j in range(3):
                put((3 * a0 + i) * 30 + 3 * b + j, k[0][i, j])
                if b >= a1:
                    put((3 * a1 + i) * 30 + 3 * b + j, k[1][i, j])
        if b > a0:                              # mirror rows: (K_{a0,b})^T | (K_{a1,b})^T in columns 6t..6t+5
            for j in range(3):
                row = (3 * b + j) * 30 + 6 * t
                for c, v in enumerate(list(k[0][:, j]) + list(k[1][:, j])):
                    put(row + c, v)
    return kt.reshape(30, 30), cnt


def test_every_tile_entry_written_and_matches_the_oracle():
    from femb200 import meshgen
    from oracle import fem_oracle as O
    nq, dN, wts = _tables()
    assert sorted(_tile_of(lane) for lane in range(30)) == sorted((t, b) for t in range(5) for b in range(2 * t, 10))
    c, t4 = meshgen.kuhn_cube(2, jitter=0.1)
    c10, e10 = O.c3d4_to_c3d10(c.numpy(), t4.numpy())[:2]
    E, nu = 2.1, 0.3
    cc = E / ((1 + nu) * (1 - 2 * nu))
    lam, mu = cc * nu, cc * (1 - 2 * nu) / 2
    rng = np.random.default_rng(0)
    for e in range(4):
        x = c10[e10[e]] + (0.02 * rng.standard_normal((10, 3)) if e >= 2 else 0.0)     # straight-sided and curved
        ref = O.c3d10_K(x, np.arange(10)[None, :], E, nu)[0]
        kt, cnt = _warp(x, lam, mu, nq, dN, wts)
        assert not np.isnan(kt).any() and cnt.min() >= 1
        assert int((cnt == 2).sum()) == 45 and cnt.max() == 2          # the five (2t+1, 2t+1) diagonal blocks, same values twice
        assert np.array_equal(kt, kt.T)
        assert np.abs(kt - ref).max() <= 1e-13 * np.abs(ref).max()
