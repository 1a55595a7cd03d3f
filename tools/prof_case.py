#!/usr/bin/env python
"""Tiny driver for ncu captures: builds a Kuhn-cube Poisson case and launches the hot kernels a few times.

    python tools/prof_case.py --n 160 --what asm,spmv,cg --reps 3
"""
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
ap.add_argument("--n", type=int, default=160)
ap.add_argument("--what", default="asm,spmv,cg")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
coords, tets = meshgen.kuhn_cube(a.n, device=dev)
N = coords.shape[0]
plan = el.CsrPlan(tets, N, dev)
crow, col = plan.pattern(1)
vals = plan.assemble_c3d4(coords, "poisson")
what = a.what.split(",")
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timed(name, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / a.reps:.3f} ms", flush=True)


if "asm" in what:
    timed("assemble_fused", lambda: plan.assemble_c3d4(coords, "poisson", out=vals, check_singular=False))
if "asm3" in what:
    v3 = torch.empty(plan.nnz_nodes * 9, device=dev, dtype=torch.float64)
    timed("assemble_fused_elasticity", lambda: plan.assemble_c3d4(coords, "elasticity", 1.0, 0.3, out=v3, check_singular=False))
if "spmv" in what:
    x = torch.randn(N, dtype=torch.float64, device=dev)
    timed("spmv", lambda: ops.spmv(crow, col, vals, x))
if "cg" in what:
    F = torch.full((N, 1), 1.0 / N, dtype=torch.float64, device=dev)
    mask = torch.ones(N, dtype=torch.uint8, device=dev)
    mask[coords[:, 2] == 0] = 0
    timed("cg_10_iters", lambda: ops.cg_solve(crow, col, vals, F, mask=mask, tol=0.0, max_iter=10, check_every=10))
