#!/usr/bin/env python
"""C3D10 element-stiffness rate on the BASELINE config-2 mesh (Kuhn n=69: 1,971,054 P2 tets, 14.2 GB of K out), one line per
kernel variant (FEMB_SOLID_WARP: 0 = CTA-phased solid_K_kernel, 2/3/4 = warp-per-element kernel with that many CTAs per SM).

    python tools/c3d10_k_rate.py [--n 69] [--variants 0,3,4,2]
The variant switch is read once per process, so every variant runs in a child process; the parent checks that all variants agree
to 1e-13 on a checksum vector (row sums of |K| of 4,096 sampled elements)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200")
for p in (ROOT, PKG, os.path.join(PKG, "solver")):
    sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=69)
ap.add_argument("--variants", default="0,3,4,2")
ap.add_argument("--child", default=None)
a = ap.parse_args()

if a.child is None:
    sums = {}
    for v in a.variants.split(","):
        env = dict(os.environ, FEMB_SOLID_WARP=v)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--n", str(a.n), "--child", v], env=env, capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print(f"variant {v}: failed\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}", flush=True)
            continue
        d = json.loads(line[-1])
        sums[v] = d.pop("checksum")
        print(json.dumps(d), flush=True)
    vs = list(sums)
    for v in vs[1:]:
        worst = max(abs(x - y) / max(abs(x), 1e-300) for x, y in zip(sums[vs[0]], sums[v]))
        print(f"variant {v} vs {vs[0]}: max rel diff of the sampled checksums {worst:.2e}", flush=True)
    sys.exit(0)

import torch  # noqa: E402
import element as el  # noqa: E402
from femb200 import meshgen  # noqa: E402

dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
c1, t1 = meshgen.kuhn_cube(a.n, device=dev)
coords, e10 = meshgen.p1_to_p2_lattice(a.n, meshgen.swap01(t1), device=dev)
del c1, t1
M, N = e10.shape[0], coords.shape[0]
out = torch.empty((M, 30, 30), dtype=torch.float64, device=dev)
for _ in range(3):
    el.compute_c3d10_K_matrix(coords, e10, 1.0, 0.3, device=dev, dtype=torch.float64, out=out)
torch.cuda.synchronize()
times = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    el.compute_c3d10_K_matrix(coords, e10, 1.0, 0.3, device=dev, dtype=torch.float64, out=out)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
times.sort()
ms = times[len(times) // 2]
bytes_K = M * (10 * 8 + 900 * 8) + N * 24
idx = torch.linspace(0, M - 1, 4096, device=dev).long()
chk = out[idx].abs().sum(dim=(1, 2)).cpu().tolist()
sym = float((out[idx] - out[idx].transpose(1, 2)).abs().max())
print(json.dumps({"variant": a.child, "n": a.n, "elements": M, "ms": round(ms, 3), "min_ms": round(times[0], 3), "hbm_frac": round(bytes_K / ms / 1e6 / HBM, 3),
                  "GBps": round(bytes_K / ms / 1e6, 1), "max_asymmetry": sym, "checksum": chk}), flush=True)
