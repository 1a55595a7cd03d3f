#!/usr/bin/env python
"""ncu driver for the 3-dof operators of BASELINE config 2: builds the P2 elasticity operator and launches the scalar-CSR and
the 3x3 block-CSR SpMV a few times.   python tools/prof_c2.py [--n 69] [--reps 2]"""
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
ap.add_argument("--n", type=int, default=69)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
c1, t1 = meshgen.kuhn_cube(a.n, device=dev)
coords, e10 = meshgen.p1_to_p2_lattice(a.n, meshgen.swap01(t1), device=dev)
N = coords.shape[0]
K = el.compute_c3d10_K_matrix(coords, e10, 1.0, 0.3, device=dev, dtype=torch.float64)
plan = el.CsrPlan(e10, N, dev)
crow, col = plan.pattern(3)
vals = plan.assemble(K, 3)
del K
brow, bcol = plan.pattern(1)
A = ops.Bsr3.from_csr_values(brow, bcol, vals)
x = torch.randn(3 * N, dtype=torch.float64, device=dev)
for _ in range(a.reps):
    ops.spmv(crow, col, vals, x)
    A.spmv(x)
torch.cuda.synchronize()
print("ok")
