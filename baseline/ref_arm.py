"""Reference arm of bench.py: the UNMODIFIED reference (its three Python files, installed by `__graft_entry__.build()` into
baseline/_ref/solver/, git-ignored, shipped to the GPU box with the snapshot) timed on the box's host cores.

The reference has no scalar (Poisson) operator: every solver of it applies 3 dofs per node (`dofs = node*3+{0,1,2}`,
reference solver/element.py:447-449).  The Poisson workload therefore goes through its public API the only way a user of the
reference could run it: element matrices `K_e = V * G^T G` from the reference's own gradients (`compute_c3d4_B_matrix`,
element.py:835-881, rows 0..2 of B hold d/dx, d/dy, d/dz) and volumes (`compute_tetrahedral_volumes`, :514-541), embedded as
`K_e (x) I_3` into the [M,12,12] layout `stable_conjugate_gradient_solver` (solver.py:144-229) takes; the load sits in
column 0 of F, so the iterates of column 0 are exactly the scalar CG iterates (columns 1, 2 stay zero).

Nothing of this module is imported by the product; it never imports femb200 / libfemb200.so.
"""
from __future__ import annotations

import contextlib
import io
import os
import re
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "solver")
REF_FILES = ("element.py", "shell.py", "solver.py")


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in REF_FILES)


def load():
    """Import the reference's modules (stub shim for its unused plotting imports).  The product mirrors the same
    top-level module names (element / shell / solver), so any of those already imported or on sys.path are evicted first."""
    if not available():
        raise ImportError(f"{REF_DIR} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    for name in ("plotly", "plotly.graph_objects", "pyvista"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
    for name in ("element", "shell", "solver"):
        mod = sys.modules.get(name)
        if mod is not None and os.path.dirname(os.path.abspath(getattr(mod, "__file__", ""))) != REF_DIR:
            del sys.modules[name]
    sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("_b200/solver")]
    sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    import element
    import solver
    assert os.path.dirname(os.path.abspath(element.__file__)) == REF_DIR, element.__file__
    return element, solver


def kuhn_cube(n):
    """Same mesh as femb200.meshgen.kuhn_cube (SURVEY 8 recipe), restated here so this arm never loads the product."""
    import torch
    ax = torch.arange(n + 1, dtype=torch.float64) / n
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    coords = torch.stack([X, Y, Z], dim=-1).reshape(-1, 3).contiguous()
    I, J, K = torch.meshgrid(torch.arange(n), torch.arange(n), torch.arange(n), indexing="ij")
    off = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))
    hexes = torch.stack([((I + a) * (n + 1) + (J + b)) * (n + 1) + (K + c) for (a, b, c) in off], dim=-1).reshape(-1, 8)
    kuhn = torch.tensor(((0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6)))
    return coords, hexes[:, kuhn].reshape(-1, 4).to(torch.int64).contiguous()


def poisson_problem(element, n):
    """(K [M,12,12], tets, F [N,3], fixed) of the Poisson cube: z=0 Dirichlet, f=1 lumped (SURVEY 8d inputs)."""
    import torch
    coords, tets = kuhn_cube(n)
    kw = dict(device="cpu", dtype=torch.float64)
    B = element.compute_c3d4_B_matrix(coords, tets, **kw)                        # [M,6,12]
    G = torch.stack([B[:, 0, 0::3], B[:, 1, 1::3], B[:, 2, 2::3]], dim=1)         # [M,3,4] = grad N_a
    del B
    V = element.compute_tetrahedral_volumes(coords, tets, **kw)
    K4 = V.view(-1, 1, 1) * torch.bmm(G.transpose(1, 2), G)                       # [M,4,4]
    M = tets.shape[0]
    K12 = torch.zeros((M, 12, 12), dtype=torch.float64)
    for i in range(3):
        K12[:, i::3, i::3] = K4
    N = coords.shape[0]
    F = torch.zeros((N, 3), dtype=torch.float64)
    F[:, 0].index_add_(0, tets.reshape(-1), (V / 4).repeat_interleave(4))
    fixed = torch.nonzero(coords[:, 2] == 0).reshape(-1)
    return K12, tets, F, fixed


def _quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def time_cg(n, iters, warmup=1, threads=None):
    """Wall time of `stable_conjugate_gradient_solver(..., tol=0, max_iter=iters)` on the n-cube (after `warmup` iterations in a
    first call).  Returns dict(rate it/s, ms_per_iter, tets, nodes, setup_s, threads)."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    element, solver = load()
    t0 = time.perf_counter()
    K, tets, F, fixed = poisson_problem(element, n)
    setup = time.perf_counter() - t0
    kw = dict(device="cpu", dtype=torch.float64)
    if warmup > 0:
        _quiet(solver.stable_conjugate_gradient_solver, K, tets, F, fixed, tol=0.0, max_iter=warmup, **kw)
    t0 = time.perf_counter()
    _quiet(solver.stable_conjugate_gradient_solver, K, tets, F, fixed, tol=0.0, max_iter=iters, **kw)
    dt = time.perf_counter() - t0
    return {"n": n, "tets": int(tets.shape[0]), "nodes": int(F.shape[0]), "iters": iters, "rate": iters / dt, "ms_per_iter": dt / iters * 1e3,
            "setup_s": round(setup, 2), "threads": torch.get_num_threads()}


def solve_c1(tol=1e-8, n=20):
    """BASELINE config 1 exactly: 20^3 Kuhn cube Poisson, CG to 1e-8 on torch CPU.  Returns (iterations, wall_s, u_max)."""
    import torch
    element, solver = load()
    K, tets, F, fixed = poisson_problem(element, n)
    t0 = time.perf_counter()
    u, text = _quiet(solver.stable_conjugate_gradient_solver, K, tets, F, fixed, tol=tol, max_iter=5000, device="cpu", dtype=torch.float64)
    dt = time.perf_counter() - t0
    m = re.search(r"Converged after (\d+) iterations", text)
    return (int(m.group(1)) if m else -1), dt, float(u[:, 0].max())


def main():
    """`python baseline/ref_arm.py --n 64 --iters 10` -> one JSON line (used by bench.py's cpu_baseline leg in a subprocess, so
    the reference's top-level modules `element` / `solver` never meet the product's modules of the same names)."""
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--c1", action="store_true", help="also solve BASELINE config 1 (n=20, tol 1e-8) and report iterations / wall")
    a = ap.parse_args()
    out = time_cg(a.n, a.iters, a.warmup)
    if a.c1:
        its, wall, umax = solve_c1()
        out["c1"] = {"iterations": its, "wall_s": round(wall, 3), "u_max": umax, "iters_per_s": round(its / wall, 2)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
