"""Synthetic structured meshes used by the tests and by bench.py (SURVEY.md section 8 recipe).

Everything is built with torch ops so the same generator runs on the CPU (tests) and on the
GPU (64 M-tet benchmark meshes); this is input synthesis, not part of the timed hot path.

Lattice node (i,j,k) -> id (i*(n+1)+j)*(n+1)+k, coords (i,j,k)/n.  Hex connectivity follows the
reference's C3D8 node order (reference solver/element.py:1537-1544).
"""
from __future__ import annotations

import torch

_HEX_OFF = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))
KUHN = ((0, 1, 2, 6), (0, 2, 3, 6), (0, 3, 7, 6), (0, 7, 4, 6), (0, 4, 5, 6), (0, 5, 1, 6))
WEDGES = ((0, 1, 2, 4, 5, 6), (0, 2, 3, 4, 6, 7))
P2_EDGES = ((0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3))


def lattice_coords(nx, ny=None, nz=None, device="cpu", dtype=torch.float64, jitter=0.0, seed=0):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    ax = torch.arange(nx + 1, device=device, dtype=dtype) / nx
    ay = torch.arange(ny + 1, device=device, dtype=dtype) / ny
    az = torch.arange(nz + 1, device=device, dtype=dtype) / nz
    X, Y, Z = torch.meshgrid(ax, ay, az, indexing="ij")
    c = torch.stack([X, Y, Z], dim=-1).reshape(-1, 3).contiguous()
    if jitter:
        g = torch.Generator(device="cpu").manual_seed(seed)
        d = (torch.rand(c.shape, generator=g, dtype=dtype) * 2 - 1).to(device)
        # keep the outer boundary planar: only interior lattice nodes move
        inner = ((c > 0) & (c < 1)).all(dim=1, keepdim=True)
        c = c + d * inner * (jitter / max(nx, ny, nz))
    return c


def hex_connectivity(nx, ny=None, nz=None, device="cpu", i0=0, i1=None):
    """[M,8] int64 hexes of an (nx,ny,nz) lattice; optional slab i0<=i<i1 of cells along x."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    i1 = nx if i1 is None else i1
    I, J, K = torch.meshgrid(torch.arange(i0, i1, device=device), torch.arange(ny, device=device),
                             torch.arange(nz, device=device), indexing="ij")
    cols = [((I + a) * (ny + 1) + (J + b)) * (nz + 1) + (K + c) for (a, b, c) in _HEX_OFF]
    return torch.stack(cols, dim=-1).reshape(-1, 8).to(torch.int64).contiguous()


def _split(hexes, table):
    t = torch.tensor(table, device=hexes.device, dtype=torch.int64)
    return hexes[:, t].reshape(-1, t.shape[1]).contiguous()


def kuhn_cube(n, device="cpu", dtype=torch.float64, jitter=0.0):
    """Conforming 6-tets-per-hex cube: (coords [N,3], tets [6 n^3, 4])."""
    return lattice_coords(n, device=device, dtype=dtype, jitter=jitter), _split(hex_connectivity(n, device=device), KUHN)


def hex_cube(n, device="cpu", dtype=torch.float64, jitter=0.0):
    return lattice_coords(n, device=device, dtype=dtype, jitter=jitter), hex_connectivity(n, device=device)


def wedge_cube(n, device="cpu", dtype=torch.float64, jitter=0.0):
    return lattice_coords(n, device=device, dtype=dtype, jitter=jitter), _split(hex_connectivity(n, device=device), WEDGES)


def mixed_box(n, device="cpu", dtype=torch.float64, jitter=0.0):
    """One lattice split along x into three slabs: hexes | wedges | Kuhn tets (BASELINE config 3).
    Interfaces are conforming on their triangulated side only where both sides triangulate the
    shared quad the same way; hex|wedge and wedge|tet slabs meet on x-planes, which wedges and Kuhn
    tets both leave as quads/triangles split along the 0-2 style diagonal.  Returns
    (coords, {"c3d8":..., "c3d6":..., "c3d4":...})."""
    a, b = n // 3, 2 * n // 3
    coords = lattice_coords(n, device=device, dtype=dtype, jitter=jitter)
    return coords, {
        "c3d8": hex_connectivity(n, device=device, i0=0, i1=a),
        "c3d6": _split(hex_connectivity(n, device=device, i0=a, i1=b), WEDGES),
        "c3d4": _split(hex_connectivity(n, device=device, i0=b, i1=n), KUHN),
    }


def swap01(tets):
    """Swap local nodes 0<->1 so the reference's C3D10 natural coordinates give detJ>0."""
    return tets[:, [1, 0, 2, 3]].contiguous()


def p1_to_p2_lattice(n, tets, device="cpu", dtype=torch.float64):
    """Closed-form P2 numbering on the (2n+1)^3 lattice: corner (i,j,k) -> (2i,2j,2k), mid-edge
    node = midpoint lattice site.  Used for the large benchmarks (not the reference's
    first-encounter numbering).  Returns (coords2 [(2n+1)^3,3], elems [M,10])."""
    m = n + 1
    k = tets % m
    j = (tets // m) % m
    i = tets // (m * m)
    m2 = 2 * n + 1

    def nid(a, b):
        return ((i[:, a] + i[:, b]) * m2 + (j[:, a] + j[:, b])) * m2 + (k[:, a] + k[:, b])

    cols = [nid(a, a) for a in range(4)] + [nid(a, b) for (a, b) in P2_EDGES]
    coords2 = lattice_coords(2 * n, device=device, dtype=dtype)
    return coords2, torch.stack(cols, dim=-1).to(torch.int64).contiguous()


def tri_sheet(n, device="cpu", dtype=torch.float64, warp=0.0):
    """n x n quads of the z=warp*sin surface split into 2 triangles each: (coords [N,3], s3 [2n^2,3])."""
    c, q = quad_sheet(n, device, dtype, warp)
    t = torch.tensor(((0, 1, 2), (0, 2, 3)), device=device)
    return c, q[:, t].reshape(-1, 3).contiguous()


def quad_sheet(n, device="cpu", dtype=torch.float64, warp=0.0):
    ax = torch.arange(n + 1, device=device, dtype=dtype) / n
    X, Y = torch.meshgrid(ax, ax, indexing="ij")
    Z = warp * torch.sin(3.0 * X) * torch.cos(2.0 * Y)
    c = torch.stack([X, Y, Z], dim=-1).reshape(-1, 3).contiguous()
    I, J = torch.meshgrid(torch.arange(n, device=device), torch.arange(n, device=device), indexing="ij")
    q = torch.stack([I * (n + 1) + J, (I + 1) * (n + 1) + J, (I + 1) * (n + 1) + J + 1, I * (n + 1) + J + 1], dim=-1)
    return c, q.reshape(-1, 4).to(torch.int64).contiguous()


def mixed_sheet(n, device="cpu", dtype=torch.float64, warp=0.0):
    """The warped n x n sheet with every second quad split into triangles (0,1,2), (0,2,3): (coords [N,3], tri [T,3], quad [S,4]),
    the mid-surface input of shell_extrude."""
    c, q = quad_sheet(n, device, dtype, warp)
    odd = (torch.arange(q.shape[0], device=device) % 2) == 1
    t = torch.tensor(((0, 1, 2), (0, 2, 3)), device=device)
    return c, q[odd][:, t].reshape(-1, 3).contiguous(), q[~odd].contiguous()


# ---- quadratic (mid-edge) versions on the (2n+1)^3 lattice -----------------------------------------------------------
HEX20_EDGES = ((0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7))   # VTK / Abaqus order
WEDGE15_EDGES = ((0, 1), (1, 2), (2, 0), (3, 4), (4, 5), (5, 3), (0, 3), (1, 4), (2, 5))


def to_quadratic_lattice(n, coords, conn, edges, compact=False):
    """Mid-edge nodes for any linear lattice connectivity: corner (i,j,k) -> site (2i,2j,2k) of the (2n+1)^3 lattice, the
    mid-edge node of corners a,b -> the site at the sum of their (i,j,k).  Coordinates of mid-edge nodes are the mean of
    the (possibly jittered) corner coordinates, so edges stay straight.  compact=True renumbers the used sites
    0..N'-1 in ascending site order (the full lattice keeps face- and body-centre sites as isolated nodes)."""
    m, m2 = n + 1, 2 * n + 1
    k, j, i = conn % m, (conn // m) % m, conn // (m * m)

    def nid(a, b):
        return ((i[:, a] + i[:, b]) * m2 + (j[:, a] + j[:, b])) * m2 + (k[:, a] + k[:, b])

    nen = conn.shape[1]
    pairs = [(a, a) for a in range(nen)] + list(edges)
    q = torch.stack([nid(a, b) for (a, b) in pairs], dim=-1).to(torch.int64).contiguous()
    c2 = torch.zeros((m2 ** 3, 3), device=coords.device, dtype=coords.dtype)
    for col, (a, b) in enumerate(pairs):
        c2[q[:, col]] = (coords[conn[:, a]] + coords[conn[:, b]]) / 2
    if compact:
        used, inv = torch.unique(q.reshape(-1), return_inverse=True)
        return c2[used].contiguous(), inv.reshape(q.shape).contiguous()
    return c2, q


def hex20_cube(n, device="cpu", dtype=torch.float64, jitter=0.0, compact=True):
    c, h = hex_cube(n, device, dtype, jitter)
    return to_quadratic_lattice(n, c, h, HEX20_EDGES, compact)


def wedge15_cube(n, device="cpu", dtype=torch.float64, jitter=0.0, compact=True):
    c, w = wedge_cube(n, device, dtype, jitter)
    return to_quadratic_lattice(n, c, w, WEDGE15_EDGES, compact)


def tet10_cube(n, device="cpu", dtype=torch.float64, jitter=0.0, compact=True):
    """Kuhn tets with local nodes 0<->1 swapped (reference detJ > 0) and mid-edge nodes in the reference order."""
    c, t = kuhn_cube(n, device, dtype, jitter)
    return to_quadratic_lattice(n, c, swap01(t), P2_EDGES, compact)


def mixed_box_quadratic(n, device="cpu", dtype=torch.float64, jitter=0.0):
    """BASELINE config 3: the hex | wedge | tet slabs of mixed_box with mid-edge nodes, one shared node numbering.
    Returns (coords [N',3], {"c3d20": [.,20], "c3d15": [.,15], "c3d10": [.,10]}); numbering is compact."""
    coords, lin = mixed_box(n, device, dtype, jitter)
    m2 = 2 * n + 1
    parts, c2 = {}, None
    for name, key, edges in (("c3d20", "c3d8", HEX20_EDGES), ("c3d15", "c3d6", WEDGE15_EDGES), ("c3d10", "c3d4", P2_EDGES)):
        conn = swap01(lin[key]) if key == "c3d4" else lin[key]
        cc, q = to_quadratic_lattice(n, coords, conn, edges, compact=False)
        parts[name] = q
        if c2 is None:
            c2 = cc
        else:
            touched = torch.zeros(m2 ** 3, dtype=torch.bool, device=cc.device)
            touched[q.reshape(-1)] = True
            c2[touched] = cc[touched]
    used, inv = torch.unique(torch.cat([p.reshape(-1) for p in parts.values()]), return_inverse=True)
    out, off = {}, 0
    for name, q in parts.items():
        out[name] = inv[off:off + q.numel()].reshape(q.shape).contiguous()
        off += q.numel()
    return c2[used].contiguous(), out
