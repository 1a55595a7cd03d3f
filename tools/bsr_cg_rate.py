#!/usr/bin/env python
"""CG / Jacobi-PCG iteration rate of a P1 elasticity operator in 3x3 block-CSR on a Kuhn cube of size n (one GPU).
    python tools/bsr_cg_rate.py --n 100        (FEMB_BSR_MERGED=1 selects the merged-reduction loop for A/B runs)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import element as el  # noqa: E402
from femb200 import meshgen, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
dev = torch.device("cuda:0")
c, t = meshgen.kuhn_cube(a.n, device=dev)
N = c.shape[0]
plan = el.CsrPlan(t, N, dev)
brow, bcol = plan.pattern(1)
A = ops.Bsr3.from_csr_values(brow, bcol, plan.assemble_c3d4(c, "elasticity", 1.0, 0.3))
mask = torch.ones((N, 3), dtype=torch.uint8, device=dev)
mask[c[:, 2] == 0] = 0
mask = mask.reshape(-1).contiguous()
F = torch.zeros((N, 3), dtype=torch.float64, device=dev)
F[c[:, 2] == 1, 2] = 1.0 / float((c[:, 2] == 1).sum())
minv = A.jacobi(mask)
out = {}
for name, kw in (("cg", dict(mask=mask)), ("pcg", dict(minv=minv))):
    best = 1e9
    for _ in range(3):
        _, info = A.cg_solve(F, tol=0.0, max_iter=a.iters, check_every=50, **kw)
        best = min(best, info["loop_ms"] / a.iters)
    out[name] = round(best * 1e3, 1)
print(f"n={a.n} nodes={N} blocks={bcol.numel()} us_per_iter={out} env={ {k: v for k, v in os.environ.items() if k.startswith('FEMB_')} }", flush=True)
