"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's three Python files are imported unmodified through a stub shim for its unused
plotting imports and called with device="cpu", dtype=float64.  Nothing
from the reference is copied; only its OUTPUTS on small seeded meshes are stored (.npz).
`c3d4_to_c3d10` is called through a one-identifier patch of its globals (`elems` -> the argument),
because the reference iterates an undefined name (solver/element.py:824).
"""
import contextlib
import io
import json
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "cuda-powered-mesh-handling-and-iterative-solvers_b200"))
sys.path.insert(0, ROOT)
from femb200 import meshgen  # noqa: E402

for name in ("plotly", "plotly.graph_objects", "pyvista"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["plotly"].graph_objects = sys.modules["plotly.graph_objects"]
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/solver")
import element as R   # noqa: E402  (the reference)
import shell as RS    # noqa: E402
import solver as RV   # noqa: E402

KW = dict(device="cpu", dtype=torch.float64)
E, NU = 1.0, 0.3
MEMB = torch.tensor([1.0, 0.3, 0.1], dtype=torch.float64)
BEND = torch.tensor([1.0, 0.3, 0.1], dtype=torch.float64)


def npy(x):
    if isinstance(x, (list, tuple)):
        return [npy(v) for v in x]
    if not torch.is_tensor(x):
        return np.asarray(x)
    return x.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB, {len(arrs)} arrays")


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **k)
    return out, buf.getvalue()


def iters_of(text):
    m = re.search(r"Converged after (\d+) iterations", text) or re.search(r"Converged @ iter (\d+)", text)
    return int(m.group(1)) if m else -1


def gen_tets():
    coords, tets = meshgen.kuhn_cube(2, jitter=0.2)
    o = dict(coords=coords, tets=tets)
    o["vol"] = R.compute_tetrahedral_volumes(coords, tets, **KW)
    o["B"] = R.compute_c3d4_B_matrix(coords, tets, **KW)
    o["K"] = R.compute_c3d4_K_matrix(coords, tets, E, NU, **KW)
    sf, s4 = R.compute_tetrahedral_surface_faces_with_fourth_node(tets, device="cpu")
    o["surf_faces"], o["surf_fourth"] = sf, s4
    o["shared"] = R.identify_tetrahedral_shared_faces(tets, device="cpu")
    o["surf_normals"] = R.compute_tetrahdral_surface_normals(coords, tets, **KW)
    o["face_normals"] = R.compute_tetrahedral_normals_and_area(coords, tets, **KW)
    o["edges"] = R.element_to_edge(tets, device="cpu")
    # P1 -> P2 (patched identifier), on positively oriented (for the reference's convention) tets
    t01 = meshgen.swap01(tets)
    R.c3d4_to_c3d10.__globals__["elems"] = t01
    c2, e10, _, _ = R.c3d4_to_c3d10(coords, t01, dtype=torch.float64)
    del R.c3d4_to_c3d10.__globals__["elems"]
    e10 = e10.to(torch.int64)
    o["coords10"], o["elems10"] = c2, e10
    ip = torch.tensor([0.2, 0.3, 0.15], dtype=torch.float64)
    o["ip10"] = ip
    o["J10"] = R.compute_c3d10_Jacobian(c2, e10, ip, **KW)
    o["g10"] = R.compute_c3d10_shape_gradients(c2, e10, ip, **KW)
    o["B10"] = R.compute_c3d10_B_matrix(c2, e10, ip, **KW)
    o["K10"] = R.compute_c3d10_K_matrix(c2, e10, E, NU, **KW)
    o["K10_multi"] = R.compute_c3d10_K_matrix(c2, e10[:5], E, NU, single=False, **KW)
    pts = torch.tensor([[0.25, 0.25, 0.25, 1 / 6], [0.1, 0.2, 0.3, 0.05]], dtype=torch.float64)
    o["pts10_custom"] = pts
    o["K10_custom"] = R.compute_c3d10_K_matrix(c2, e10, E, NU, integral_point=pts, **KW)
    o["tets_from10"] = R.c3d10_to_c3d4(e10, device="cpu")
    p, w = R.c3d10_integration_points(**KW)
    o["pts10"], o["w10"] = p, w
    u = torch.randn(c2.shape[0], 3, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    o["u10"] = u
    o["f10"] = R.compute_nodal_forces(o["K10"], e10, u, **KW)
    save("tets", **{k: npy(v) for k, v in o.items()})


def gen_hex():
    coords, hexes = meshgen.hex_cube(2, jitter=0.2)
    o = dict(coords=coords, hexes=hexes)
    o["vol"] = R.compute_hexahedral_volumes(coords, hexes, **KW)
    ip = torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64)
    o["ip"] = ip
    o["J"] = R.compute_c3d8_Jacobian(coords, hexes, ip, **KW)
    o["g"] = R.compute_c3d8_shape_gradients(coords, hexes, ip, **KW)
    o["B"] = R.compute_c3d8_B_matrix(coords, hexes, ip, **KW)
    o["K"] = R.compute_c3d8_K_matrix(coords, hexes, E, NU, **KW)
    o["K_multi"] = R.compute_c3d8_K_matrix(coords, hexes[:3], E, NU, single=False, **KW)
    p, w = R.c3d8_integration_points(**KW)
    o["pts"], o["w"] = p, w
    sf, ex = R.compute_hexahedral_surface_faces_with_extra_node(hexes, device="cpu")
    o["surf_faces"], o["surf_extra"] = sf, ex
    o["shared"] = R.identify_hexahedral_shared_faces(hexes, device="cpu")
    o["surf_normals"] = R.compute_hexahedral_surface_normals(coords, hexes, **KW)
    o["face_normals"] = R.compute_hexahedral_normals_and_area(coords, hexes, **KW)
    o["tets"] = R.c3d8_to_c3d4(hexes, device="cpu")
    save("hexes", **{k: npy(v) for k, v in o.items()})


def gen_wedge():
    coords, w6 = meshgen.wedge_cube(2, jitter=0.2)
    o = dict(coords=coords, wedges=w6)
    o["vol"] = R.compute_wedge_volumes(coords, w6, **KW)
    ip = torch.tensor([0.2, 0.3, -0.4], dtype=torch.float64)
    o["ip"] = ip
    o["J"] = R.compute_c3d6_Jacobian(coords, w6, ip, **KW)
    o["g"] = R.compute_c3d6_shape_gradients(coords, w6, ip, **KW)
    o["B"] = R.compute_c3d6_B_matrix(coords, w6, ip, **KW)
    o["K_single"] = R.compute_c3d6_K_matrix(coords, w6, E, NU, single=True, **KW)
    o["K_full"] = R.compute_c3d6_K_matrix(coords, w6, E, NU, single=False, **KW)
    p, w = R.c3d6_integration_points(**KW)
    o["pts"], o["w"] = p, w
    (q, t), (qe, te) = R.compute_wedge_surface_faces_with_extra_node(w6, device="cpu")
    o["surf_quads"], o["surf_tris"], o["quad_extra"], o["tri_extra"] = q, t, qe, te
    nq, nt = R.compute_wedge_surface_normals(coords, w6, **KW)
    o["nq"], o["nt"] = nq, nt
    o["tets"] = R.c3d6_to_c3d4(w6, device="cpu")
    save("wedges", **{k: npy(v) for k, v in o.items()})


def gen_shells():
    c3, s3 = meshgen.tri_sheet(3, warp=0.15)
    c4, s4 = meshgen.quad_sheet(3, warp=0.15)
    o = dict(c3=c3, s3=s3, c4=c4, s4=s4, membrane=MEMB, bending=BEND)
    o["D"] = RS.compute_kirchoff_D_matrix(MEMB, BEND, **KW)
    o["unit3"] = RS.compute_s3_local_unitvector(c3, s3, device="cpu")
    o["J3"] = RS.compute_s3_jacobian(c3, s3, **KW)
    o["g3"] = RS.compute_s3_shape_gradient(c3, s3, **KW)
    o["B3"] = RS.compute_s3_B_matrix(c3, s3, **KW)
    o["K3"] = RS.compute_s3_K_matrix(c3, s3, MEMB, BEND, **KW)
    o["shared3"] = RS.identify_s3_shared_edges(s3, device="cpu")
    e, t = RS.compute_triangle_surface_faces_with_third_node(s3, device="cpu")
    o["bedges3"], o["bthird3"] = e, t
    o["unit4"] = RS.compute_s4_local_unitvector(c4, s4, device="cpu")
    xi, eta = 0.3, -0.6
    o["xieta"] = np.array([xi, eta])
    o["J4"] = RS.compute_s4_jacobian(c4, s4, xi, eta, **KW)
    o["g4"] = RS.compute_s4_shape_gradient(c4, s4, xi, eta, **KW)
    o["B4"] = RS.compute_s4_B_matrix_single(c4, s4, xi, eta, **KW)
    o["K4"] = RS.compute_s4_K_matrix(c4, s4, MEMB, BEND, **KW)
    o["K4_multi"] = RS.compute_s4_K_matrix(c4, s4, MEMB, BEND, single=False, **KW)
    p, w = RS.s4_integration_points(device="cpu")
    o["pts4"], o["w4"] = p.double(), w.double()
    o["shared4"] = RS.identify_s4_shared_edges(s4, device="cpu")
    e, t = RS.compute_square_surface_faces_with_fourth_node(s4, device="cpu")
    o["bedges4"], o["bfourth4"] = e, t
    g = torch.Generator().manual_seed(5)
    u3 = torch.randn(c3.shape[0], 6, dtype=torch.float64, generator=g)
    u4 = torch.randn(c4.shape[0], 6, dtype=torch.float64, generator=g)
    o["u3"], o["u4"] = u3, u4
    o["f3"] = RS.compute_shell_nodal_forces(o["K3"], s3, u3, o["unit3"], **KW)
    o["f4"] = RS.compute_shell_nodal_forces(o["K4"], s4, u4, o["unit4"], **KW)
    save("shells", **{k: npy(v) for k, v in o.items()})


def gen_units():
    """The known-answer table of SURVEY.md section 4 (unit elements, E=1, nu=0.3)."""
    o = {}
    c = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]], dtype=torch.float64)
    t = torch.tensor([[0, 1, 2, 3]])
    o["c3d4"] = R.compute_c3d4_K_matrix(c, t, E, NU, **KW)
    R.c3d4_to_c3d10.__globals__["elems"] = t
    c2, e10, _, _ = R.c3d4_to_c3d10(c, t, dtype=torch.float64)
    del R.c3d4_to_c3d10.__globals__["elems"]
    o["c3d10"] = R.compute_c3d10_K_matrix(c2, e10.long(), E, NU, **KW)
    ch, h = meshgen.hex_cube(1)
    o["c3d8"] = R.compute_c3d8_K_matrix(ch, h, E, NU, **KW)
    w = h[:, [0, 1, 2, 4, 5, 6]]
    o["c3d6_single"] = R.compute_c3d6_K_matrix(ch, w, E, NU, single=True, **KW)
    o["c3d6_full"] = R.compute_c3d6_K_matrix(ch, w, E, NU, single=False, **KW)
    cs = torch.tensor([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0.0]], dtype=torch.float64)
    o["s3"] = RS.compute_s3_K_matrix(cs, torch.tensor([[0, 1, 3]]), MEMB, BEND, **KW)
    o["s4"] = RS.compute_s4_K_matrix(cs, torch.tensor([[0, 1, 2, 3]]), MEMB, BEND, **KW)
    o["D"] = R.compute_elasticity_matrix(E, NU, **KW)
    save("units", **{k: npy(v) for k, v in o.items()})


def gen_solve():
    n = 4
    coords, tets = meshgen.kuhn_cube(n)
    N = coords.shape[0]
    K = R.compute_c3d4_K_matrix(coords, tets, E, NU, **KW)
    fixed = torch.nonzero(coords[:, 2] == 0).reshape(-1)
    top = torch.nonzero(coords[:, 2] == 1).reshape(-1)
    F = torch.zeros(N, 3, dtype=torch.float64)
    F[top, 2] = 1.0 / top.numel()
    o = dict(coords=coords, tets=tets, fixed=fixed, F=F)
    u, txt = quiet(RV.stable_conjugate_gradient_solver, K, tets, F, fixed, tol=1e-8, max_iter=1000, **KW)
    o["u_cg"], o["it_cg"] = u, iters_of(txt)
    u2, txt = quiet(RV.final_solver, K, tets, F, fixed, tol=1e-8, max_iter=1000, **KW)
    o["u_final"], o["it_final"] = u2.detach(), iters_of(txt)
    # reference (buggy) diagonal builder, stored for the oracle's reproduction of it
    o["Minv_ref"] = RV.compute_diagonal_preconditioner(K, tets, N, **KW)
    # PCG loop driven by the correct Jacobi diagonal with fixed rows zeroed
    diag = torch.zeros(N * 3, dtype=torch.float64)
    dofs = (tets.unsqueeze(-1) * 3 + torch.arange(3)).reshape(-1)
    diag.index_add_(0, dofs, K.diagonal(dim1=1, dim2=2).reshape(-1))
    Minv = (1.0 / diag).reshape(N, 3)
    Minv[fixed] = 0.0
    o["Minv"] = Minv
    u3, txt = quiet(RV.preconditioned_conjugate_gradient_solver, K, tets, F, Minv, tol=1e-8, max_iter=1000, **KW)
    o["u_pcg"], o["it_pcg"] = u3, iters_of(txt)
    # coalesced CSR of the notebook's COO (subdivision.ipynb cell 6), via torch semantics
    M = K.shape[0]
    dof = (tets.unsqueeze(-1).repeat(1, 1, 3) * 3 + torch.arange(3).view(1, 1, 3)).view(M, -1)
    row = dof.unsqueeze(2).repeat(1, 1, 12).view(-1)
    col = dof.unsqueeze(1).repeat(1, 12, 1).view(-1)
    csr = torch.sparse_coo_tensor(torch.stack([row, col]), K.view(-1), size=(3 * N, 3 * N)).coalesce().to_sparse_csr()
    o["crow"], o["col"], o["val"] = csr.crow_indices(), csr.col_indices(), csr.values()
    save("solve_c3d4", **{k: npy(v) if torch.is_tensor(v) else v for k, v in o.items()})

    # mixed static_structure_solver: hex | wedge | tet slabs plus a quad+tri skin on z=1
    n = 3
    coords, parts = meshgen.mixed_box(n)
    N = coords.shape[0]
    m = n + 1
    top_ids = torch.tensor([[(i * m + j) * m + n for j in range(m)] for i in range(m)])
    quads = torch.stack([top_ids[:-1, :-1], top_ids[1:, :-1], top_ids[1:, 1:], top_ids[:-1, 1:]], -1).reshape(-1, 4)
    s4 = quads[: quads.shape[0] // 2]
    tq = quads[quads.shape[0] // 2:]
    s3 = torch.cat([tq[:, [0, 1, 2]], tq[:, [0, 2, 3]]], 0)
    fixed = torch.nonzero(coords[:, 2] == 0).reshape(-1)
    force = torch.zeros(N, 6, dtype=torch.float64)
    force[top_ids.reshape(-1), 2] = -1.0 / top_ids.numel()
    material = {"E": E, "nu": NU, "membrane": MEMB, "bending": BEND}
    u, txt = quiet(RV.static_structure_solver, coords, force, fixed, c3d4=parts["c3d4"], c3d6=parts["c3d6"],
                   c3d8=parts["c3d8"], s3=s3, s4=s4, material=material, tol=1e-8, max_iter=2000, **KW)
    save("solve_mixed", coords=npy(coords), c3d4=npy(parts["c3d4"]), c3d6=npy(parts["c3d6"]), c3d8=npy(parts["c3d8"]),
         s3=npy(s3), s4=npy(s4), fixed=npy(fixed), force=npy(force), u=npy(u), it=iters_of(txt))

    # shell CG on a flat triangle sheet (local frame == global frame, so the load stays in the
    # range of the w/theta_z-free Kirchhoff operator and the reference loop converges)
    c3, s3 = meshgen.tri_sheet(4, warp=0.0)
    K3 = RS.compute_s3_K_matrix(c3, s3, MEMB, BEND, **KW)
    unit = RS.compute_s3_local_unitvector(c3, s3, device="cpu")
    fixed = torch.nonzero(c3[:, 0] == 0).reshape(-1)
    Fs = torch.zeros(c3.shape[0], 6, dtype=torch.float64)
    Fs[c3[:, 0] == 1, 0] = 0.2
    Fs[c3[:, 0] == 1, 4] = 0.01
    us, txt = quiet(RV.stable_conjugate_gradient_shell_solver, K3, s3, Fs, fixed, unit=unit, tol=1e-9, max_iter=3000, **KW)
    save("solve_shell", c3=npy(c3), s3=npy(s3), fixed=npy(fixed), F=npy(Fs), u=npy(us), it=iters_of(txt))


def gen_stress():
    """Stress recovery: reference outputs on the small seeded meshes."""
    g = torch.Generator().manual_seed(9)
    o = {}
    coords, tets = meshgen.kuhn_cube(2, jitter=0.2)
    u = torch.randn(coords.shape[0], 3, dtype=torch.float64, generator=g) * 0.01
    o["c"], o["t"], o["u"] = coords, tets, u
    s, v = R.compute_c3d4_element_stress(coords, tets, u, E, NU, **KW)
    o["s4"], o["v4"] = s, v
    o["node_vm4"] = R.compute_node_vm_stress(coords, tets, v, **KW)
    t01 = meshgen.swap01(tets)
    R.c3d4_to_c3d10.__globals__["elems"] = t01
    c2, e10, _, _ = R.c3d4_to_c3d10(coords, t01, dtype=torch.float64)
    del R.c3d4_to_c3d10.__globals__["elems"]
    e10 = e10.long()
    u10 = torch.randn(c2.shape[0], 3, dtype=torch.float64, generator=g) * 0.01
    o["c10"], o["e10"], o["u10"] = c2, e10, u10
    o["s10"], o["v10"] = R.compute_c3d10_element_stress(c2, e10, u10, E, NU, **KW)
    sm, vm = R.compute_c3d10_element_stress(c2, e10[:6], u10, E, NU, single=False, **KW)
    o["s10m"], o["v10m"] = sm, vm
    ch, h = meshgen.hex_cube(2, jitter=0.2)
    uh = torch.randn(ch.shape[0], 3, dtype=torch.float64, generator=g) * 0.01
    o["ch"], o["h"], o["uh"] = ch, h, uh
    o["s8"], o["v8"] = R.compute_c3d8_element_stress(ch, h, uh, E, NU, **KW)
    o["s8m"], o["v8m"] = R.compute_c3d8_element_stress(ch, h, uh, E, NU, single=False, **KW)
    cw, w6 = meshgen.wedge_cube(2, jitter=0.2)
    o["w6"] = w6
    o["s6"], o["v6"] = R.compute_c3d6_element_stress(cw, w6, uh, E, NU, **KW)
    o["s6m"], o["v6m"] = R.compute_c3d6_element_stress(cw, w6, uh, E, NU, single=False, **KW)
    save("stress", **{k: npy(v) for k, v in o.items()})


CONSTRAINTS = dict(
    spc=[{"node": 0, "dofs": [0, 1, 2], "value": 0.0}, {"node": 4, "dofs": [0, 1, 2], "value": 0.0},
         {"node": 20, "dofs": [0, 1, 2], "value": 0.0}, {"node": 24, "dofs": [0, 1, 2], "value": 0.0},
         {"node": 12, "dofs": [2], "value": 0.002}, {"node": 2, "dofs": [0, 2], "value": 0.0}],
    rbe2=[{"master": 124, "slaves": [123, 119, 118], "dofs": [0, 1, 2]}, {"master": 104, "slaves": [103], "dofs": [2]}],
    rbe3=[{"master": 62, "slaves": [61, 63, 57, 67], "dofs": [0, 1, 2], "weights": [1.0, 2.0, 1.0, 0.5]},
          {"master": 112, "slaves": [111, 113], "dofs": [2], "weights": [1.0, 3.0]}],
    loads=[{"node": 124, "force": [0.0, 0.0, -0.01]}, {"node": 100, "force": [0.002, 0.0, -0.005]},
           {"node": 62, "force": [0.0, 0.003, 0.0]}])


def gen_constrained():
    """SPC / RBE2 / RBE3 constrained CG on the n=4 Kuhn cube (125 nodes)."""
    coords, tets = meshgen.kuhn_cube(4, jitter=0.1)
    N = coords.shape[0]
    K = R.compute_c3d4_K_matrix(coords, tets, E, NU, **KW)
    C = CONSTRAINTS
    F = torch.zeros(N, 3, dtype=torch.float64)
    RV.apply_loads_to_F(F, C["loads"])
    u1, t1 = quiet(RV.constrained_conjugate_gradient_solver, K, tets, F, C["rbe2"], C["spc"], tol=1e-9, max_iter=2000, **KW)
    u2, t2 = quiet(RV.new_constrained_conjugate_gradient_solver, K, tets, N, C["rbe2"], C["rbe3"], C["spc"], C["loads"], tol=1e-9,
                   max_iter=2000, **KW)
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(N, 3, dtype=torch.float64, generator=g) * 1e-3
    u3, t3 = quiet(RV.constrained_conjugate_gradient_solver, K, tets, F, C["rbe2"], C["spc"], u_init=u0, tol=1e-9, max_iter=2000, **KW)
    sp = RV.parse_spc_list(C["spc"], device="cpu")
    r2 = RV.parse_rbe2_list(C["rbe2"], device="cpu")
    r3 = RV.parse_rbe3_list(C["rbe3"], device="cpu")
    save("constrained", coords=npy(coords), tets=npy(tets), F=npy(F), u_c=npy(u1), it_c=iters_of(t1), u_n=npy(u2), it_n=iters_of(t2),
         u0=npy(u0), u_c0=npy(u3), it_c0=iters_of(t3), spc_n=npy(sp[0]), spc_d=npy(sp[1]), spc_v=npy(sp[2]), r2_s=npy(r2[0]),
         r2_m=npy(r2[1]), r2_d=npy(r2[2]), r3_m=npy(r3[0]), r3_s=npy(r3[1]), r3_d=npy(r3[2]), r3_w=npy(r3[3]), r3_i=npy(r3[4]),
         r3_ws=npy(r3[5]), constraints_json=np.array(json.dumps(C)))


def gen_partition():
    """subdivision.ipynb cells 7-9 executed as written (CPU), with torch.randint pinned so the first seed is reproducible."""
    import math
    nb = json.load(open("/root/reference/subdivision.ipynb"))
    ns = {"torch": torch, "math": math}
    exec("".join(nb["cells"][7]["source"]).split("\nprint(")[0], ns)
    exec("".join(nb["cells"][9]["source"]), ns)
    coords, tets = meshgen.kuhn_cube(4, jitter=0.1)
    M = tets.shape[0]
    K = R.compute_c3d4_K_matrix(coords, tets, E, NU, **KW)
    sh = R.identify_tetrahedral_shared_faces(tets, device="cpu")
    edge = torch.cat([sh[:, 0, 0].unsqueeze(0), sh[:, 1, 0].unsqueeze(0)], dim=0)
    first = 37
    real_randint = torch.randint
    torch.randint = lambda *a, **k: torch.tensor([first])
    try:
        groups, seeds = ns["region_growing_partition"](edge, 5, M, device="cpu")
    finally:
        torch.randint = real_randint
    labels = torch.full((M,), -1, dtype=torch.long)
    for i, g in enumerate(groups):
        labels[g] = i
    Kl, gn = ns["build_sparse_K_local"](K, tets, groups[2], device="cpu")
    csr = Kl.coalesce().to_sparse_csr()
    save("partition", coords=npy(coords), tets=npy(tets), edge=npy(edge), first=first, labels=npy(labels), seeds=npy(seeds),
         group2=npy(groups[2]), nodes2=npy(gn), crow2=npy(csr.crow_indices()), col2=npy(csr.col_indices()), val2=npy(csr.values()),
         subdiv=np.array([ns["compute_subdivisions"](338619, 10), ns["compute_subdivisions"](1000, 1), ns["compute_subdivisions"](3000000, 4)]))


def gen_quadratic():
    """The two C3D20 functions of the reference that run: the 27-point rule and the 24-tet table."""
    p, w = R.c3d20_integration_points(**KW)
    h20 = torch.arange(40, dtype=torch.int64).reshape(2, 20) * 3 + 1
    save("quadratic", p20=npy(p), w20=npy(w), h20=npy(h20), h20_tets=npy(R.c3d20_to_c3d4(h20, device="cpu")))


def gen_subdomain_forces():
    """subdivision.ipynb cells 13 and 15 executed as written (CPU) on a 5-part partition of a small cube."""
    from collections import defaultdict
    nb = json.load(open("/root/reference/subdivision.ipynb"))
    ns = {"torch": torch, "defaultdict": defaultdict, "print": lambda *a, **k: None}
    exec("".join(nb["cells"][13]["source"]).split("\ngroup_to_nodes =")[0], ns)
    exec("".join(nb["cells"][15]["source"]), ns)
    g = np.load(os.path.join(HERE, "partition.npz"))
    tets, labels = torch.tensor(g["tets"]), torch.tensor(g["labels"])
    node_maps = [torch.unique(tets[labels == p]) for p in range(5)]
    g2n = ns["build_ordered_subdomain_map"](node_maps)
    keys = sorted(g2n)                                   # a fixed group order for the fixture
    g2n = {k: sorted(g2n[k]) for k in keys}
    n_free = sum((len(k) - 1) * len(v) for k, v in g2n.items())
    gen = torch.Generator().manual_seed(21)
    fv = torch.randn(n_free, 3, dtype=torch.float64, generator=gen)
    N = int(tets.max()) + 1
    F = torch.randn(N, 3, dtype=torch.float64, generator=gen)
    iface = sorted(n for v in g2n.values() for n in v)
    rbe2 = iface[::7]                                    # some interface nodes are fixed: they consume unknowns, receive nothing
    out = ns["make_sub_domain_forces"](fv, g2n, rbe2, F, 5, device="cpu")
    flat_keys = np.array([list(k) + [-1] * (5 - len(k)) for k in keys])
    flat_nodes = np.concatenate([np.array(g2n[k]) for k in keys])
    counts = np.array([len(g2n[k]) for k in keys])
    save("subdomain_forces", keys=flat_keys, nodes=flat_nodes, counts=counts, fv=npy(fv), F=npy(F), rbe2=np.array(rbe2),
         out=npy(torch.stack(out)), node_maps_flat=np.concatenate([npy(m) for m in node_maps]), node_maps_len=np.array([m.numel() for m in node_maps]))


def gen_modal():
    """vectorized_modal_solver (solver.py:1084-1312) on CPU / fp64 with its torch.randn start pinned to a stored X0.  The
    reference cannot finish: its Gauss-Jordan helper divides a row in place by a view of one of its own entries
    (`row_i /= pivot`, solver.py:1232-1234), which torch rejects -- recorded here; the oracle restates the evident intent."""
    from oracle import fem_oracle as O
    coords, tets = meshgen.kuhn_cube(3, jitter=0.15)
    K = R.compute_c3d4_K_matrix(coords, tets, E, NU, **KW)
    Mloc = torch.tensor(O.c3d4_mass(coords.numpy(), tets.numpy(), 2.5))      # the reference has no mass routine: M_local is an input
    N = coords.shape[0]
    fixed = torch.nonzero(coords[:, 2] == 0).reshape(-1)
    gen = torch.Generator().manual_seed(8)
    X0 = torch.randn(3 * N, 4, dtype=torch.float64, generator=gen)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: X0.clone()
    raised = ""
    try:
        RV.vectorized_modal_solver(K, Mloc, tets, fixed, N, num_eigs=4, max_iter=2, **KW)
    except RuntimeError as exc:
        raised = str(exc)[:120]
    finally:
        torch.randn = real_randn
    save("modal", coords=npy(coords), tets=npy(tets), Mloc=npy(Mloc), fixed=npy(fixed), X0=npy(X0), reference_raises=np.array(raised))


def gen_notebook_calls():
    """Names the reference's two notebooks call that its own solver modules define: what a drop-in has to offer."""
    import builtins
    called = set()
    for nbf in ("/root/reference/solver_example.ipynb", "/root/reference/subdivision.ipynb"):
        for c in json.load(open(nbf))["cells"]:
            if c["cell_type"] == "code":
                called |= set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", "".join(c["source"])))
    defined = set(dir(R)) | set(dir(RS)) | set(dir(RV))
    nb_defs = set()
    for c in json.load(open("/root/reference/subdivision.ipynb"))["cells"]:
        nb_defs |= set(re.findall(r"^def ([A-Za-z_][A-Za-z0-9_]*)", "".join(c["source"]), flags=re.M))
    names = sorted(n for n in called if (n in defined or n in nb_defs) and not hasattr(builtins, n) and not hasattr(torch, n))
    with open(os.path.join(HERE, "notebook_calls.json"), "w") as f:
        json.dump({"module_functions": [n for n in names if n in defined], "notebook_functions": [n for n in names if n in nb_defs]}, f, indent=1)
    print("notebook_calls:", len(names), "names")


def gen_widen():
    """Functions either side of the element path: shell frames / stress / post-processing, shell extrusion, wedge face normals,
    tet face-force balance and the operator-callback CG."""
    o = {}
    c3, s3 = meshgen.tri_sheet(3, warp=0.15)
    c4, s4 = meshgen.quad_sheet(3, warp=0.15)
    unit3 = RS.compute_s3_local_unitvector(c3, s3, device="cpu")
    unit4 = RS.compute_s4_local_unitvector(c4, s4, device="cpu")
    o["normal3"] = RS.compute_s3_normal(c3, s3, device="cpu")
    o["normal4"] = RS.compute_s4_normal(c4, s4, device="cpu")
    o["loc3"] = RS.compute_s3_global_to_local_coordinates(c3, s3, unit3, **KW)
    o["loc4"] = RS.compute_s4_global_to_local_coordinates(c4, s4, unit4, **KW)
    g = torch.Generator().manual_seed(11)
    u3 = torch.randn(c3.shape[0], 6, dtype=torch.float64, generator=g)
    u4 = torch.randn(c4.shape[0], 6, dtype=torch.float64, generator=g)
    o["u3"], o["u4"] = u3, u4
    o["ul3"] = RS.compute_global_to_local_displacement(s3, u3, unit3, device="cpu")
    o["ul4"] = RS.compute_global_to_local_displacement(s4, u4, unit4, device="cpu")
    o["B4_sum"] = RS.compute_s4_B_matrix(c4, s4, single=True, **KW)
    o["B4_all"] = RS.compute_s4_B_matrix(c4, s4, single=False, **KW)
    # the same single-point functions handed 0-dim float32 tensors (what compute_s4_B_matrix passes down, shell.py:813-814)
    p4, _ = RS.s4_integration_points(device="cpu")
    o["J4_t"] = RS.compute_s4_jacobian(c4, s4, p4[0, 0], p4[0, 1], **KW)
    o["g4_t"] = RS.compute_s4_shape_gradient(c4, s4, p4[0, 0], p4[0, 1], **KW)
    o["stress3"] = RS.compute_s3_shell_stress(c3, s3, MEMB, BEND, u3, **KW)
    o["stress4"] = RS.compute_s4_shell_stress(c4, s4, MEMB, BEND, u4, **KW)
    tz = (0.1, 0.03)
    o["post_tz"] = np.array(tz)
    post = RS.compute_shell_postprocess_values(o["stress4"], tz[0], z=tz[1], **KW)
    o["post_keys"] = np.array(list(post.keys()))
    o["post"] = torch.stack([post[k] for k in post])
    post32 = RS.compute_shell_postprocess_values(o["stress4"], tz[0], z=tz[1], device="cpu")
    o["post32"] = torch.stack([post32[k] for k in post32])
    # extrusion of a mixed tri / quad mid-surface (fp64 and the default fp32)
    cm, tri, quad = meshgen.mixed_sheet(4, warp=0.2)
    o["cm"], o["tri"], o["quad"], o["thickness"] = cm, tri, quad, np.array(0.07)
    x3, w6, h8 = RS.shell_extrude(cm, tri, quad, 0.07, **KW)
    o["ext_x"], o["ext_w"], o["ext_h"] = x3, w6, h8
    o["ext_x32"] = RS.shell_extrude(cm, tri, quad, 0.07, device="cpu")[0]
    o["ext_x_tri_only"] = RS.shell_extrude(cm, tri, quad[:0], 0.07, **KW)[0]
    o["ext_x_quad_only"] = RS.shell_extrude(cm, tri[:0], quad, 0.07, **KW)[0]
    # wedge face normals
    cw, w = meshgen.wedge_cube(2, jitter=0.2)
    o["cw"], o["w"] = cw, w
    try:   # the reference takes torch.cross(..., dim=3) of 3-d tensors (element.py:2409): IndexError on every input
        R.compute_wedge_normals_and_area(cw, w, **KW)
        o["wedge_normals_raises"] = np.array(0)
    except IndexError:
        o["wedge_normals_raises"] = np.array(1)
    # tet face forces and their balance over shared faces
    ct, tets = meshgen.kuhn_cube(2, jitter=0.2)
    o["ct"], o["tets"] = ct, tets
    nrm = R.compute_tetrahedral_normals_and_area(ct, tets, **KW)
    sig = torch.randn(tets.shape[0], 3, 3, dtype=torch.float64, generator=g)
    sig = sig + sig.transpose(1, 2)
    o["sigma"] = sig
    o["face_forces"] = R.compute_c3d4_surface_forces(nrm, sig, device="cpu")
    shared = R.identify_tetrahedral_shared_faces(tets, device="cpu")
    o["shared"] = shared
    o["shared_sum"] = R.compute_c3d4_shared_face_forces_sum(shared, o["face_forces"], device="cpu")
    # conjugate_gradient_solver_Ku on the SPD operator u -> K u + 0.1 u (no boundary conditions in that loop)
    K = R.compute_c3d4_K_matrix(ct, tets, E, NU, **KW)
    Rv = torch.randn(ct.shape[0], 3, dtype=torch.float64, generator=g)
    o["ku_R"] = Rv
    o["ku_shift"] = np.array(0.1)
    o["ku_u"], _ = quiet(RV.conjugate_gradient_solver_Ku, lambda v: R.compute_nodal_forces(K, tets, v, **KW) + 0.1 * v, Rv, tol=1e-10,
                         max_iter=500, **KW)
    save("widen", **{k: npy(v) for k, v in o.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_units()
    gen_tets()
    gen_hex()
    gen_wedge()
    gen_shells()
    gen_solve()
    gen_stress()
    gen_quadratic()
    gen_constrained()
    gen_partition()
    gen_widen()
    gen_subdomain_forces()
    gen_modal()
    gen_notebook_calls()
