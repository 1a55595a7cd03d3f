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
    ok = ok and check_elasticity_routes(rank, world, dev)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


def check_elasticity_routes(rank, world, dev):
    """3-dof operators over several GPUs through the solver API (distributed=True): 3x3 block-CSR rows per rank, projected CG
    and Jacobi-PCG, P1 (node-id partition) and P2 (RCB on coordinates), against the single-process oracle / single-GPU loop."""
    import element as el
    import solver as sv
    from femb200 import meshgen
    from oracle import fem_oracle as O
    ok = True
    n = int(os.environ.get("FEMB_CHECK_NE", "6"))
    c, t = meshgen.kuhn_cube(n, jitter=0.15)
    K = el.compute_c3d4_K_matrix(c, t, 1.0, 0.3, device=dev, dtype=torch.float64)
    fixed = torch.nonzero(c[:, 2] == 0).reshape(-1)
    F = torch.zeros(c.shape[0], 3, dtype=torch.float64)
    F[c[:, 2] == 1, 2] = 1.0 / float((c[:, 2] == 1).sum())
    u, info = sv.stable_conjugate_gradient_solver(K, t, F, fixed, tol=1e-9, max_iter=5000, device=dev, return_info=True, verbose=False,
                                                  distributed=True)
    u2, info2 = sv.stable_conjugate_gradient_solver(K, t, F, fixed, tol=1e-9, max_iter=5000, device=dev, return_info=True, verbose=False,
                                                    distributed=True)
    assert torch.equal(u, u2) and info["iterations"] == info2["iterations"]          # deterministic, flags survive a second solve
    if rank == 0:
        uo, ito, st = O.stable_cg(K.cpu().numpy(), t.numpy(), F.numpy(), fixed.numpy(), tol=1e-9, max_iter=5000)
        err = np.abs(u.cpu().numpy() - uo).max() / np.abs(uo).max()
        good = st == "converged" and info["status"] == "converged" and abs(info["iterations"] - ito) <= 1 and err < 1e-8
        print(f"dist_gpu_check world={world} P1 elasticity (block-CSR, node-id partition): iterations {info['iterations']} (oracle {ito}), "
              f"rel err {err:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        ok = ok and good
    # Jacobi-PCG with the corrected diagonal, P2 elements, RCB partition on the P2 coordinates
    c2, e10, _, _ = el.c3d4_to_c3d10(c, meshgen.swap01(t), dtype=torch.float64)
    e10 = e10.long()
    K2 = el.compute_c3d10_K_matrix(c2, e10, 1.0, 0.3, device=dev, dtype=torch.float64)
    fixed2 = torch.nonzero(c2[:, 2] == 0).reshape(-1)
    F2 = torch.zeros(c2.shape[0], 3, dtype=torch.float64)
    F2[c2[:, 2] == 1, 2] = 1.0 / float((c2[:, 2] == 1).sum())
    Minv = sv.compute_diagonal_preconditioner(K2, e10, c2.shape[0], device=dev, dtype=torch.float64)
    Minv[fixed2.to(Minv.device)] = 0.0              # fixed rows: M_inv = 0 keeps them at zero in the BC-less PCG loop
    up, ip = sv.preconditioned_conjugate_gradient_solver(K2, e10, F2, Minv, tol=1e-9, max_iter=5000, device=dev, dtype=torch.float64,
                                                         return_info=True, verbose=False, distributed=True, coords=c2)
    if rank == 0:
        u1, i1 = sv.preconditioned_conjugate_gradient_solver(K2, e10, F2, Minv, tol=1e-9, max_iter=5000, device=dev, dtype=torch.float64,
                                                             return_info=True, verbose=False)
        err = float((up - u1).abs().max() / u1.abs().max())
        good = ip["status"] == i1["status"] == "converged" and abs(ip["iterations"] - i1["iterations"]) <= 1 and err < 1e-8
        print(f"dist_gpu_check world={world} P2 elasticity Jacobi-PCG (block-CSR, RCB): iterations {ip['iterations']} (1 GPU {i1['iterations']}), "
              f"rel err {err:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
        ok = ok and good
    dist.barrier()
    return ok


if __name__ == "__main__":
    main()
