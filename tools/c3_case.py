#!/usr/bin/env python
"""BASELINE config 3: mixed quadratic hex / wedge / tet mesh -- stiffness + mass element matrices, CSR assembly, surface and
face-connectivity extraction on one B200, fp64.

    python tools/c3_case.py [--n 60]
The box is split along x into C3D20 | C3D15 | C3D10 slabs (n^3/3 hexes, 2 n^3/3 wedges, 2 n^3 tets; n=60 -> 648,000 elements,
n=66 -> 862,488) sharing one node numbering.  Prints one JSON line: per type and stage the time, elements/s and the fraction
of the HBM roofline on SURVEY 8(d)'s algorithmic bytes  M (nen ib + nd^2 8) + 24 N  (element matrices) and
M nd^2 (8+4) + 8 nnz (two-step assembly).
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
from femb200 import meshgen, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=69)
a = ap.parse_args()
dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
KW = dict(device=dev, dtype=torch.float64)


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


coords, parts = meshgen.mixed_box_quadratic(a.n, device=dev)
N = coords.shape[0]
out = {"workload": f"mixed quadratic box n={a.n}: " + ", ".join(f"{v.shape[0]} {k}" for k, v in parts.items()) + f", {N} nodes",
       "dtype": "f64", "hbm_peak_gbs": HBM}
E, nu, rho = 1.0, 0.3, 1.0
tot_ms, tot_M = 0.0, 0
x = torch.randn(3 * N, dtype=torch.float64, device=dev)
for kind, conn in parts.items():
    M, nen = conn.shape
    nd = 3 * nen
    r = {"elements": M}
    bytes_K = M * (nen * 8 + nd * nd * 8) + N * 24
    ms, K = timed(lambda: el.compute_K_matrix(coords, conn, kind, E, nu, **KW))
    r["K"] = {"ms": round(ms, 2), "elems_per_s": round(M / ms * 1e3), "hbm_frac": round(bytes_K / ms / 1e6 / HBM, 3)}
    tot_ms += ms
    ms, Mm = timed(lambda: el.compute_M_matrix(coords, conn, kind, rho, **KW))
    r["M"] = {"ms": round(ms, 2), "elems_per_s": round(M / ms * 1e3), "hbm_frac": round(bytes_K / ms / 1e6 / HBM, 3)}
    tot_ms += ms
    t0 = time.perf_counter()
    plan = el.CsrPlan(conn, N, dev)
    torch.cuda.synchronize()
    r["plan_s"] = round(time.perf_counter() - t0, 3)
    crow, col = plan.pattern(3)
    vals = torch.empty(plan.nnz_nodes * 9, device=dev, dtype=torch.float64)
    nnz = vals.numel()
    bytes_A = M * nd * nd * 12 + nnz * 8
    for name, Ke in (("assemble_K", K), ("assemble_M", Mm)):
        ms, _ = timed(lambda: plan.assemble(Ke, 3, out=vals))
        r[name] = {"ms": round(ms, 2), "elems_per_s": round(M / ms * 1e3), "nnz": nnz, "hbm_frac": round(bytes_A / ms / 1e6 / HBM, 3)}
        tot_ms += ms
    # sanity on the assembled mass: total = rho * slab volume (per direction)
    ones = torch.zeros(3 * N, dtype=torch.float64, device=dev)
    ones[0::3] = 1
    r["mass_total"] = round(float(ops.spmv(crow, col, vals, ones).sum()), 12)
    ms, _ = timed(lambda: ops.spmv(crow, col, vals, x), reps=5)
    r["spmv"] = {"ms": round(ms, 3), "hbm_frac": round((nnz * 12 + 3 * N * 20) / ms / 1e6 / HBM, 3), "layout": "scalar CSR"}
    # what the solvers actually use for 3-dof operators: 3x3 block-CSR (8.44 instead of 12 bytes per nonzero, a third of the gathers)
    brow, bcol = plan.pattern(1)
    A3 = ops.Bsr3.from_csr_values(brow, bcol, vals)
    ms3, _ = timed(lambda: A3.spmv(x), reps=5)
    own = nnz * 8 + (nnz // 9) * 4 + N * 4 + 3 * N * 16
    r["spmv_bsr3"] = {"ms": round(ms3, 3), "hbm_frac_scalar_csr_bytes": round((nnz * 12 + 3 * N * 20) / ms3 / 1e6 / HBM, 3),
                      "hbm_frac_bytes_moved": round(own / ms3 / 1e6 / HBM, 3)}
    del A3
    tot_M += M
    out[kind] = r
    del K, Mm, vals, plan
out["stiffness_plus_mass_assembled_elems_per_s"] = round(tot_M / tot_ms * 1e3)

topo = {}
ms, (f, _) = timed(lambda: el.compute_hexahedral_surface_faces_with_extra_node(parts["c3d20"], device=dev), reps=2)
topo["hex_surface"] = {"ms": round(ms, 2), "faces": f.shape[0], "elems_per_s": round(parts["c3d20"].shape[0] / ms * 1e3)}
ms, s = timed(lambda: el.identify_hexahedral_shared_faces(parts["c3d20"], device=dev), reps=2)
topo["hex_shared"] = {"ms": round(ms, 2), "pairs": s.shape[0]}
ms, ((q, t), _) = timed(lambda: el.compute_wedge_surface_faces_with_extra_node(parts["c3d15"], device=dev), reps=2)
topo["wedge_surface"] = {"ms": round(ms, 2), "quads": q.shape[0], "tris": t.shape[0], "elems_per_s": round(parts["c3d15"].shape[0] / ms * 1e3)}
ms, (f, _) = timed(lambda: el.compute_tetrahedral_surface_faces_with_fourth_node(parts["c3d10"], device=dev), reps=2)
topo["tet_surface"] = {"ms": round(ms, 2), "faces": f.shape[0], "elems_per_s": round(parts["c3d10"].shape[0] / ms * 1e3)}
ms, s = timed(lambda: el.identify_tetrahedral_shared_faces(parts["c3d10"], device=dev), reps=2)
topo["tet_shared"] = {"ms": round(ms, 2), "pairs": s.shape[0]}
out["topology"] = topo
print(json.dumps(out), flush=True)
