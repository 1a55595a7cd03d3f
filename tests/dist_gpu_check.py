"""Launched by torchrun (2+ ranks, one GPU each): the peer-memory distributed CG must reproduce the single-process
oracle solve (1e-8 relative, iterations +-1).  Used by tests/test_dist_gpu.py and by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    from femb200 import dist_cg, meshgen
    n = int(os.environ.get("FEMB_CHECK_N", "14"))
    coords, tets = meshgen.kuhn_cube(n, jitter=0.15)
    coords, tets = coords.to(dev), tets.to(dev)
    part, plan, op, cl = dist_cg.setup_poisson_p1(coords, tets, rank, world, dev)
    no = part.n_owned
    vol = torch.zeros(part.n_local, dtype=torch.float64, device=dev)
    import element as el
    v = el.compute_tetrahedral_volumes(cl, part.elements_local, device=dev, dtype=torch.float64) / 4
    vol.index_add_(0, part.elements_local.reshape(-1), v.repeat_interleave(4))
    F = vol[:no]
    mask = (cl[:no, 2] != 0).to(torch.uint8).contiguous()
    u, info = op.solve(F, mask, tol=1e-9, max_iter=3000)
    # run it twice: flags/epochs must survive a second solve, and the result must be bit-identical (deterministic)
    u2, info2 = op.solve(F, mask, tol=1e-9, max_iter=3000)
    assert torch.equal(u, u2) and info2["iterations"] == info["iterations"], (info, info2)
    full = torch.zeros(coords.shape[0], dtype=torch.float64, device=dev)
    full[part.owned_global] = u
    dist.all_reduce(full)
    ok = True
    if rank == 0:
        from oracle import fem_oracle as O
        c, t = coords.cpu().numpy(), tets.cpu().numpy()
        Ke = O.c3d4_poisson_K(c, t)
        load = np.bincount(t.reshape(-1), weights=np.repeat(O.tet_volumes(c, t) / 4, 4), minlength=c.shape[0]).reshape(-1, 1)
        fixed = np.flatnonzero(c[:, 2] == 0)
        uo, ito, st = O.stable_cg(Ke, t, load, fixed, tol=1e-9, max_iter=3000, ndof=1)
        err = np.abs(full.cpu().numpy() - uo[:, 0]).max() / np.abs(uo).max()
        ok = st == "converged" and info["status"] == "converged" and abs(info["iterations"] - ito) <= 1 and err < 1e-8
        print(f"dist_gpu_check world={world}: iterations {info['iterations']} (oracle {ito}), rel err {err:.2e}, "
              f"neighbors {part.neighbors}, ghosts {part.n_ghost}, loop {info['loop_ms']:.2f} ms -> {'OK' if ok else 'FAIL'}", flush=True)
    op.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
