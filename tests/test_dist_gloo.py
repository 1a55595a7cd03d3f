"""World-size 2/3 CPU (gloo) tests of the multi-GPU host logic: RCB partition, local numbering, halo plan.
Each rank assembles its local operator with the CPU oracle and runs the reference CG loop with a real halo exchange and
all-reduce over torch.distributed; the result must equal the single-process oracle solve (1e-8, iterations +-1)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT

for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def _problem(n=6):
    from femb200 import meshgen
    c, t = meshgen.kuhn_cube(n, jitter=0.15)
    fixed = torch.nonzero(c[:, 2] == 0).reshape(-1)
    return c, t, fixed


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from femb200 import partition
    from oracle import fem_oracle as O
    c, t, fixed = _problem()
    labels = partition.rcb_labels(c, world)
    part = partition.build_local_part(t, labels, rank, world)
    cl = partition.localize(c, part).numpy()
    el = part.elements_local.numpy()
    Ke = O.c3d4_poisson_K(cl, el)
    crow, col, val, _ = O.assemble_csr(Ke, el, 1, part.n_local)
    no = part.n_owned
    crow, col, val = crow[:no + 1], col[:crow[no]], val[:crow[no]]        # owned rows are a prefix of the local CSR
    load = np.bincount(el.reshape(-1), weights=np.repeat(O.tet_volumes(cl, el) / 4, 4), minlength=part.n_local)
    # loads of owned nodes are complete (all their elements are local)
    F = load[:no].copy()
    isfixed = torch.zeros(c.shape[0], dtype=torch.bool)
    isfixed[fixed] = True
    free = (~isfixed[part.owned_global]).numpy().astype(np.float64)

    def halo(x_owned):
        xl = np.zeros(part.n_local)
        xl[:no] = x_owned
        reqs, bufs = [], {}
        for q in part.neighbors:
            bufs[q] = torch.empty(part.recv_cnt[q], dtype=torch.float64)
            reqs.append(dist.irecv(bufs[q], src=q))
            reqs.append(dist.isend(torch.from_numpy(x_owned[part.send_idx[q].numpy()]).contiguous(), dst=q))
        for r in reqs:
            r.wait()
        for q in part.neighbors:
            o = part.ghost_base + part.recv_off[q]
            xl[o:o + part.recv_cnt[q]] = bufs[q].numpy()
        return xl

    def gsum(v):
        t_ = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t_)
        return float(t_.item())

    def apply(x_owned):
        return O.csr_matvec(crow, col, val, halo(x_owned)) * free

    # the reference loop (solver.py:144-229) with global reductions
    u = np.zeros(no)
    r = (F - apply(u)) * free
    p = r.copy()
    rs_old = gsum(float(r @ r))
    its, tol, eps = 0, 1e-9, 1e-30
    for i in range(2000):
        Ap = apply(p)
        pAp = gsum(float(p @ Ap))
        alpha = rs_old / (pAp + eps)
        u += alpha * p
        r -= alpha * Ap
        rs_new = gsum(float(r @ r))
        its = i + 1
        if np.sqrt(rs_new) < tol:
            break
        p = r + (rs_new / (rs_old + eps)) * p
        rs_old = rs_new
    out.put((rank, part.owned_global.numpy(), u, its, part.n_ghost, sorted(part.neighbors)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_cg_matches_single(world):
    from oracle import fem_oracle as O
    c, t, fixed = _problem()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cn, tn = c.numpy(), t.numpy()
    Ke = O.c3d4_poisson_K(cn, tn)
    load = np.bincount(tn.reshape(-1), weights=np.repeat(O.tet_volumes(cn, tn) / 4, 4), minlength=cn.shape[0]).reshape(-1, 1)
    u_ref, it_ref, st = O.stable_cg(Ke, tn, load, fixed.numpy(), tol=1e-9, ndof=1)
    assert st == "converged"
    u = np.zeros(cn.shape[0])
    seen = np.zeros(cn.shape[0], dtype=int)
    for rank, owned, ul, its, n_ghost, nbrs in res:
        u[owned] = ul
        seen[owned] += 1
        assert abs(its - it_ref) <= 1
        assert n_ghost > 0 and len(nbrs) >= 1
    assert (seen == 1).all()                       # every node owned exactly once
    assert np.abs(u - u_ref[:, 0]).max() <= 1e-8 * np.abs(u_ref).max()


def test_rcb_is_balanced_and_deterministic():
    from femb200 import partition
    c, t, _ = _problem(8)
    for P in (2, 3, 4, 8):
        lab = partition.rcb_labels(c, P)
        cnt = torch.bincount(lab, minlength=P)
        assert cnt.max() - cnt.min() <= 1 and torch.equal(lab, partition.rcb_labels(c, P))
    # halo plans agree pairwise: what r sends to q is what q expects from r, in the same (global id) order
    lab = partition.rcb_labels(c, 4)
    parts = [partition.build_local_part(t, lab, r, 4) for r in range(4)]
    for r in range(4):
        for q in parts[r].neighbors:
            sent = parts[r].owned_global[parts[r].send_idx[q]]
            o = parts[q].recv_off[r]
            assert torch.equal(sent, parts[q].ghost_global[o:o + parts[q].recv_cnt[r]])


def test_block_partition_halo_plans_agree():
    """The coordinate-free partition of the solver API's multi-GPU route (node-id ranges): balanced, every node owned once,
    and the pairwise halo plans agree like the RCB ones."""
    from femb200 import partition
    c, t, _ = _problem(7)
    N = c.shape[0]
    for P in (2, 3, 5):
        lab = partition.block_labels(N, P)
        cnt = torch.bincount(lab, minlength=P)
        assert cnt.sum() == N and cnt.max() - cnt.min() <= 1 and bool((lab[1:] >= lab[:-1]).all())
        parts = [partition.build_local_part(t, lab, r, P) for r in range(P)]
        owned = torch.cat([p.owned_global for p in parts])
        assert torch.equal(torch.sort(owned).values, torch.arange(N))
        for r in range(P):
            assert parts[r].n_interior <= parts[r].n_owned
            for q in parts[r].neighbors:
                sent = parts[r].owned_global[parts[r].send_idx[q]]
                o = parts[q].recv_off[r]
                assert torch.equal(sent, parts[q].ghost_global[o:o + parts[q].recv_cnt[r]])
                assert int(parts[r].send_idx[q].min()) >= parts[r].n_interior      # only boundary rows are sent


@pytest.mark.parametrize("block", [1, 3])
@pytest.mark.parametrize("kind", ["rcb", "ids"])
def test_halo_tables_deliver_every_ghost(block, kind):
    """partition.halo_tables (the flat tables femb_dist_cg_solve receives) emulated on the host: every rank "pushes" through
    send_idx / ghost_off (the separate push kernel) and through the boundary-row table bptr / bk / boff (the push folded into the
    vector kernel); both must fill every ghost dof of every rank with the owner's value, for 1 and 3 dofs per node."""
    from femb200 import partition
    c, t, _ = _problem(6)
    N, P, B = c.shape[0], 4, block
    lab = partition.rcb_labels(c, P) if kind == "rcb" else partition.block_labels(N, P)
    parts = [partition.build_local_part(t, lab, r, P) for r in range(P)]
    sizes = [{"ghost_base": p.ghost_base, "recv_off": p.recv_off, "n_local": p.n_local} for p in parts]
    tabs = [partition.halo_tables(p, sizes, B) for p in parts]
    g = torch.Generator().manual_seed(3)
    xg = torch.randn(N, B, dtype=torch.float64, generator=g)                 # a global vector
    for mode in ("push", "folded"):
        # local p vectors: owned part filled, ghosts poisoned
        x = [torch.full((p.n_local * B,), float("nan"), dtype=torch.float64) for p in parts]
        for r, p in enumerate(parts):
            x[r][:p.n_owned * B] = xg[p.owned_global].reshape(-1)
        for r, (p, tb) in enumerate(zip(parts, tabs)):
            if mode == "push" or tb["bptr"] is None:      # (no interior rows on this rank: the library keeps the push kernel)
                for k, q in enumerate(tb["nbr"]):
                    a, b = tb["send_ptr"][k], tb["send_ptr"][k + 1]
                    x[q][tb["ghost_off"][k]: tb["ghost_off"][k] + (b - a)] = x[r][tb["send_idx"][a:b].long()]
            else:
                ni = p.n_interior * B
                assert tb["bptr"].numel() == (p.n_owned - p.n_interior) * B + 1
                for row in range(tb["bptr"].numel() - 1):
                    for e in range(int(tb["bptr"][row]), int(tb["bptr"][row + 1])):
                        k = int(tb["bk"][e])
                        x[tb["nbr"][k]][tb["ghost_off"][k] + int(tb["boff"][e])] = x[r][ni + row]
        for r, p in enumerate(parts):
            gb = p.ghost_base * B
            want = xg[p.ghost_global].reshape(-1)
            assert torch.equal(x[r][gb: gb + p.n_ghost * B], want), (mode, r)
            assert gb % 16 == 0                                              # ghosts start on their own 128-byte line
