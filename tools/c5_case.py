#!/usr/bin/env python
"""BASELINE config 5: hybrid subdivided-mesh solver, coarse Kuhn n=8 (3,072 tets) refined 3x -> 1,572,864 tets."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import element as el  # noqa: E402
import solver as sv  # noqa: E402
from femb200 import meshgen  # noqa: E402

dev = "cuda:0"
c0, t0 = meshgen.kuhn_cube(8)


def load_fn(c, t):
    F = torch.zeros(c.shape[0], 3, dtype=torch.float64, device=c.device)
    top = c[:, 2] > 1 - 1e-9
    F[top, 2] = -1.0 / float(top.sum())
    return F


def fixed_fn(c):
    return torch.nonzero(c[:, 2] < 1e-9).reshape(-1)


sv.hybrid_subdivided_solver(c0, t0, 1, load_fn, fixed_fn, E=1.0, nu=0.3, tol=1e-8, device=dev, verbose=False)   # warm-up
torch.cuda.synchronize()
t = time.perf_counter()
u, cf, tf, info = sv.hybrid_subdivided_solver(c0, t0, 3, load_fn, fixed_fn, E=1.0, nu=0.3, tol=1e-8, device=dev, verbose=False)
torch.cuda.synchronize()
t_h = time.perf_counter() - t
K = el.compute_c3d4_K_matrix(cf, tf, 1.0, 0.3, device=dev, dtype=torch.float64)
torch.cuda.synchronize()
t = time.perf_counter()
uc, ic = sv.stable_conjugate_gradient_solver(K, tf, load_fn(cf, tf), fixed_fn(cf), tol=1e-8, max_iter=20000, device=dev, return_info=True, verbose=False)
torch.cuda.synchronize()
t_c = time.perf_counter() - t
err = float((u - uc).abs().max() / uc.abs().max())
print(json.dumps({"workload": f"hybrid cascade, coarse {t0.shape[0]} tets -> {tf.shape[0]} tets, {cf.shape[0]} nodes", "levels": info["levels"],
                  "hybrid_wall_s": round(t_h, 3), "cold_cg_wall_s": round(t_c, 3), "cold_cg_iterations": ic["iterations"],
                  "rel_diff_vs_cold_cg": err}))
