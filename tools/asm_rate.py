#!/usr/bin/env python
"""Fused P1 Poisson assembly rate (coords -> CSR values) on a Kuhn cube of size n (one GPU).
    python tools/asm_rate.py --n 220            (FEMB_ASM_BLOCKS=1 selects the block-owned kernel; FEMB_ASM_BLOCK_ROWS / _THREADS / FEMB_ASM_VERBOSE tune it)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import element as el  # noqa: E402
from femb200 import meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=220)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--jitter", type=float, default=0.0)
a = ap.parse_args()
dev = "cuda:0"
c, t = meshgen.kuhn_cube(a.n, device=dev, jitter=a.jitter)
plan = el.CsrPlan(t, c.shape[0], dev)
vals = plan.assemble_c3d4(c, "poisson")
v2 = plan.assemble_c3d4(c, "poisson")
assert torch.equal(vals, v2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    plan.assemble_c3d4(c, "poisson", out=vals, check_singular=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
M, N, nnz = t.shape[0], c.shape[0], vals.numel()
by = M * 32 + N * 24 + nnz * 8
print(f"n={a.n} M={M} asm_ms={ms:.3f} Gelem/s={M / ms / 1e6:.1f} frac_strict={by / ms / 1e6 / 6448.7:.3f} "
      f"env={ {k: v for k, v in os.environ.items() if k.startswith('FEMB_')} }", flush=True)
