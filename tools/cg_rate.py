#!/usr/bin/env python
"""CG iteration rate of the P1 Poisson operator on a Kuhn cube of size n (one GPU): the per-rank problem of an N-GPU run can
be emulated with n = 220 / N^(1/3) (n=110: the 1.35 M rows one of 8 ranks owns on BASELINE config 4).
    python tools/cg_rate.py --n 110 --iters 400"""
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
ap.add_argument("--n", type=int, default=110)
ap.add_argument("--iters", type=int, default=400)
a = ap.parse_args()
dev = torch.device("cuda:0")
coords, tets = meshgen.kuhn_cube(a.n, device=dev)
N = coords.shape[0]
plan = el.CsrPlan(tets, N, dev)
crow, col = plan.pattern(1)
vals = plan.assemble_c3d4(coords, "poisson")
F = torch.full((N, 1), 1.0 / N, dtype=torch.float64, device=dev)
mask = torch.ones(N, dtype=torch.uint8, device=dev)
mask[coords[:, 2] == 0] = 0
best = 1e9
for _ in range(3):
    u, info = ops.cg_solve(crow, col, vals, F, mask=mask, tol=0.0, max_iter=a.iters, check_every=50)
    best = min(best, info["loop_ms"] / a.iters)
x = torch.randn(N, dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.spmv(crow, col, vals, x)
e0.record()
for _ in range(20):
    ops.spmv(crow, col, vals, x)
e1.record()
torch.cuda.synchronize()
print(f"n={a.n} rows={N} nnz={vals.numel()} cg_us_per_iter={best * 1e3:.1f} spmv_us={e0.elapsed_time(e1) / 20 * 1e3:.1f} "
      f"env={ {k: v for k, v in os.environ.items() if k.startswith('FEMB_')} }", flush=True)
