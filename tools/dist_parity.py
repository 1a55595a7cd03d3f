#!/usr/bin/env python
"""Repeated N-GPU vs 1-GPU parity solves on the C4 mesh (torchrun, one rank per GPU): prints one line per repetition.
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/dist_parity.py --n 220 --reps 3"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=220)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--soak", type=int, default=1)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
from femb200 import dist_cg, meshgen  # noqa: E402
coords, tets = meshgen.kuhn_cube(a.n, device=dev)
N = coords.shape[0]
part, plan, op, cl = dist_cg.setup_poisson_p1(coords, tets, rank, world, dev)
del tets
no = part.n_owned
mask = (cl[:no, 2] != 0).to(torch.uint8).contiguous()
F = torch.full((no,), 1.0 / N, dtype=torch.float64, device=dev)
for rep in range(a.reps):
    for _ in range(a.soak):
        op.solve(F, mask, tol=0.0, max_iter=300, check_every=50)
        op.solve(F, mask, tol=0.0, max_iter=20, check_every=20)
    par = dist_cg.parity_check(coords, part, plan, op, cl, mask, rank, world, dev, N)
    if rank == 0:
        print("PARITY", json.dumps({k: par[k] for k in ("ok", "iterations_N", "iterations_1gpu", "rel_err_u", "true_residual_N_over_f")}),
              {k: v for k, v in os.environ.items() if k.startswith("FEMB_")}, flush=True)
op.close()
dist.destroy_process_group()
