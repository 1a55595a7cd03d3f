#!/usr/bin/env python
"""Face-connectivity rate on the Kuhn cube of size n (surface faces + shared-face pairs in one pass).
    python tools/topo_rate.py --n 220"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from femb200 import meshgen, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=220)
ap.add_argument("--opts", default="", help="comma-separated FEMB_TOPO_OPT values to sweep (A/B switches of csrc/topology.cu)")
ap.add_argument("--check", action="store_true", help="all swept variants must return identical tensors")
a = ap.parse_args()
dev = torch.device("cuda:0")
_, tets = meshgen.kuhn_cube(a.n, device=dev)
ref = None
for opt in (a.opts.split(",") if a.opts else [None]):
    if opt is not None:
        os.environ["FEMB_TOPO_OPT"] = opt
    ops.entities(ops.ENT_TET_FACES, tets, dev)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        f, x, p = ops.entities(ops.ENT_TET_FACES, tets, dev)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    same = ""
    if a.check:
        if ref is None:
            ref = (f, x, p)
        else:
            same = f" identical={all(torch.equal(u, v) for u, v in zip(ref, (f, x, p)))}"
    print(f"n={a.n} opt={opt} tets={tets.shape[0]} surface={f.shape[0]} shared={p.shape[0]} ms={best:.2f} "
          f"elems_per_s={tets.shape[0] / best * 1e3:.3e}{same}", flush=True)
    del f, x, p
