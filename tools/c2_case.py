#!/usr/bin/env python
"""BASELINE config 2: P2 tet linear elasticity, ~2 M tets, fp64: element K, CSR assembly, Jacobi-PCG on one B200.

    python tools/c2_case.py [--n 69] [--iters 100]
Prints one JSON line with per-stage times and roofline fractions (HBM and fp64-FMA where relevant).
"""
import argparse
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
from femb200 import meshgen, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=69)
ap.add_argument("--iters", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    out = None
    for _ in range(reps):
        out = None   # release the previous result first: a second multi-GB output block would be cudaMalloc'ed inside the timed region
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


c1, t1 = meshgen.kuhn_cube(a.n, device=dev)
coords, e10 = meshgen.p1_to_p2_lattice(a.n, meshgen.swap01(t1), device=dev)
del c1, t1
M, N = e10.shape[0], coords.shape[0]
E, nu = 1.0, 0.3
out = {"workload": f"P2 tet elasticity, Kuhn n={a.n}: {M} C3D10 tets, {N} nodes, {3 * N} dofs", "dtype": "f64"}

ms_K, K = timed(lambda: el.compute_c3d10_K_matrix(coords, e10, E, nu, device=dev, dtype=torch.float64))
bytes_K = M * (10 * 8 + 900 * 8) + N * 24
out["element_K"] = {"ms": round(ms_K, 2), "elems_per_s": round(M / ms_K * 1e3), "hbm_frac": round(bytes_K / ms_K / 1e6 / HBM, 3),
                    "fp64_frac_of_37TF": round(M * 9000 * 2 / (ms_K * 1e-3) / 1e12 / 37.0, 3)}  # ~9 k DFMA per element
assert float(K.sum(dim=2).abs().max()) < 1e-9 * float(K.abs().max()) + 1e-9  # rigid translations: rows sum to ~0 per dof group? (sanity only)

t0 = time.perf_counter()
plan = el.CsrPlan(e10, N, dev)
torch.cuda.synchronize()
out["plan_s"] = round(time.perf_counter() - t0, 2)
crow, col = plan.pattern(3)
vals = torch.empty(plan.nnz_nodes * 9, device=dev, dtype=torch.float64)
ms_A, _ = timed(lambda: plan.assemble(K, 3, out=vals))
nnz = vals.numel()
bytes_A = M * 900 * (8 + 4) + nnz * 8
out["assemble_from_Ke"] = {"ms": round(ms_A, 2), "elems_per_s": round(M / ms_A * 1e3), "nnz": nnz, "hbm_frac_two_step_bytes": round(bytes_A / ms_A / 1e6 / HBM, 3)}
out["assembled_elems_per_s_total"] = round(M / (ms_K + ms_A) * 1e3)

x = torch.randn(3 * N, dtype=torch.float64, device=dev)
ms_S, _ = timed(lambda: ops.spmv(crow, col, vals, x), reps=5)
bytes_S = nnz * 12 + 3 * N * 20
out["spmv"] = {"ms": round(ms_S, 3), "GBps": round(bytes_S / ms_S / 1e6, 1), "hbm_frac": round(bytes_S / ms_S / 1e6 / HBM, 3)}

brow, bcol = plan.pattern(1)
t0 = time.perf_counter()
A = ops.Bsr3.from_csr_values(brow, bcol, vals)
torch.cuda.synchronize()
ms_B, _ = timed(lambda: A.spmv(x), reps=5)
bytes_B = nnz * 8 + (nnz // 9) * 4 + N * 4 + 3 * N * 16
out["spmv_bsr3"] = {"ms": round(ms_B, 3), "GBps_own_bytes": round(bytes_B / ms_B / 1e6, 1), "hbm_frac_own_bytes": round(bytes_B / ms_B / 1e6 / HBM, 3),
                    "hbm_frac_csr_bytes": round(bytes_S / ms_B / 1e6 / HBM, 3), "convert_s": round(time.perf_counter() - t0, 3)}

fixed = torch.nonzero(coords[:, 2] == 0).reshape(-1)
mask = torch.ones((N, 3), dtype=torch.uint8, device=dev)
mask[fixed] = 0
mask = mask.reshape(-1).contiguous()
F = torch.zeros((N, 3), dtype=torch.float64, device=dev)
F[coords[:, 2] == 1, 2] = 1.0 / float((coords[:, 2] == 1).sum())
minv = ops.jacobi(crow, col, vals, mask)
u, info = ops.cg_solve(crow, col, vals, F, minv=minv, tol=0.0, max_iter=a.iters, check_every=min(a.iters, 50))
bytes_it = bytes_S + 11 * 3 * N * 8
out["jacobi_pcg"] = {"iters": info["iterations"], "ms_per_iter": round(info["loop_ms"] / a.iters, 3), "iters_per_s": round(a.iters / info["loop_ms"] * 1e3, 1),
                     "hbm_frac": round(bytes_it / (info["loop_ms"] / a.iters) / 1e6 / HBM, 3)}
u3, info3 = A.cg_solve(F, minv=minv, tol=0.0, max_iter=a.iters, check_every=min(a.iters, 50))
bytes_itb = bytes_B + 11 * 3 * N * 8
out["jacobi_pcg_bsr3"] = {"ms_per_iter": round(info3["loop_ms"] / a.iters, 3), "iters_per_s": round(a.iters / info3["loop_ms"] * 1e3, 1),
                          "hbm_frac_own_bytes": round(bytes_itb / (info3["loop_ms"] / a.iters) / 1e6 / HBM, 3),
                          "hbm_frac_csr_bytes": round(bytes_it / (info3["loop_ms"] / a.iters) / 1e6 / HBM, 3),
                          "max_abs_diff_vs_csr": float((u3 - u).abs().max())}
u2, info2 = ops.cg_solve(crow, col, vals, F, mask=mask, tol=0.0, max_iter=a.iters, check_every=min(a.iters, 50))
out["cg"] = {"ms_per_iter": round(info2["loop_ms"] / a.iters, 3), "iters_per_s": round(a.iters / info2["loop_ms"] * 1e3, 1)}
u4, info4 = A.cg_solve(F, mask=mask, tol=0.0, max_iter=a.iters, check_every=min(a.iters, 50))
out["cg_bsr3"] = {"ms_per_iter": round(info4["loop_ms"] / a.iters, 3), "iters_per_s": round(a.iters / info4["loop_ms"] * 1e3, 1),
                  "loop": "merged (2 kernels)" if os.environ.get("FEMB_BSR_MERGED") and not os.environ.get("FEMB_CG_CLASSIC") else "classic (3 kernels)"}
print(json.dumps(out), flush=True)
