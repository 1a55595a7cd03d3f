#!/usr/bin/env python
"""Per-iteration time of the peer-memory CG for several cube sizes (run under torchrun): latency floor vs compute."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
from femb200 import dist_cg, meshgen, ops  # noqa: E402

for n in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32,110,220").split(",")]:
    coords, tets = meshgen.kuhn_cube(n, device=dev)
    part, plan, op, cl = dist_cg.setup_poisson_p1(coords, tets, rank, world, dev)
    no = part.n_owned
    mask = (cl[:no, 2] != 0).to(torch.uint8).contiguous()
    F = torch.full((no,), 1.0, dtype=torch.float64, device=dev)
    op.solve(F, mask, tol=0.0, max_iter=10, check_every=10)
    K = 200
    u, info = op.solve(F, mask, tol=0.0, max_iter=K, check_every=50)
    ms = torch.tensor([info["loop_ms"]], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    line = f"world={world} n={n} rows/rank={no} ghosts={part.n_ghost} nbrs={len(part.neighbors)}: {1e3 * ms.item() / K:.1f} us/iter"
    if world == 1:
        crow, col = plan.pattern(1)
        val = plan.assemble_c3d4(cl, "poisson")
        nz = int(crow[no].item())     # owned rows are a prefix of the (line-padded) local pattern
        u1, i1 = ops.cg_solve(crow[:no + 1].contiguous(), col[:nz].contiguous(), val[:nz].contiguous(), F.reshape(-1, 1), mask=mask,
                              tol=0.0, max_iter=K, check_every=50)
        line += f"   (single-GPU path: {1e3 * i1['loop_ms'] / K:.1f} us/iter)"
    if rank == 0:
        print(line, flush=True)
    op.close()
    del op, plan, part, coords, tets
    torch.cuda.empty_cache()
dist.destroy_process_group()
