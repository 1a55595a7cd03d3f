"""Pin the CPU oracle (oracle/fem_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import fem_oracle as O

TOL = 1e-12   # north_star: fp64 element matrices within 1e-12 relative
E, NU = 1.0, 0.3


def close(a, b, tol=TOL):
    assert np.asarray(a).shape == np.asarray(b).shape, (np.asarray(a).shape, np.asarray(b).shape)
    assert rel_err(a, b) <= tol, rel_err(a, b)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a.astype(np.int64), b.astype(np.int64))


def test_unit_known_answers():
    """SURVEY.md section 4 table, to the digits printed there."""
    g = load_golden("units")
    table = {
        "c3d4": (0.35256410256410248, 0.16025641025641024, 2.115384615384615, 1.1377076505960799),
        "c3d10": (-0.53053846153846151, -0.24115384615384616, -26.523538461538461, 8.9516789275141999),
        "c3d8": (0.23504273504273501, 0.080128205128205107, 5.6410256410256405, 1.7240292952089149),
        "c3d6_single": (0.18963675213675213, 0.0, 2.8205128205128207, 1.4326600409376284),
        "c3d6_full": (0.51282050879366603, 0.0, 7.7564102057870388, 3.0423403688209252),
        "s3": (0.074175824175824176, 0.035714285714285712, 0.29695054945054944, 0.19426018350482771),
        "s4": (0.049450549006774501, 0.017857142857142856, 0.39593406238090784, 0.19269984464953877),
    }
    c = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    t = np.array([[0, 1, 2, 3]])
    c2, e10 = O.c3d4_to_c3d10(c, t)
    ch = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], float)
    h = np.array([[0, 4, 6, 2, 1, 5, 7, 3]])
    w = h[:, [0, 1, 2, 4, 5, 6]]
    cs = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0.0]])
    mb = (1.0, 0.3, 0.1)
    mine = {
        "c3d4": O.c3d4_K(c, t, E, NU), "c3d10": O.c3d10_K(c2, e10, E, NU), "c3d8": O.c3d8_K(ch, h, E, NU),
        "c3d6_single": O.c3d6_K(ch, w, E, NU, single=True), "c3d6_full": O.c3d6_K(ch, w, E, NU, single=False),
        "s3": O.s3_K(cs, np.array([[0, 1, 3]]), mb, mb), "s4": O.s4_K(cs, np.array([[0, 1, 2, 3]]), mb, mb),
    }
    for name, (k00, k01, tr, fro) in table.items():
        K = mine[name][0]
        assert abs(K[0, 0] - k00) <= 1e-13 * max(1, abs(k00)), name
        assert abs(K[0, 1] - k01) <= 1e-13, name
        assert abs(np.trace(K) - tr) <= 1e-12 * abs(tr), name
        assert abs(np.linalg.norm(K) - fro) <= 1e-12 * fro, name
        close(mine[name], g[name])
    close(O.elasticity_matrix(E, NU), g["D"])


def test_tets():
    g = load_golden("tets")
    c, t = g["coords"], g["tets"]
    close(O.tet_volumes(c, t), g["vol"])
    close(O.c3d4_B(c, t), g["B"])
    close(O.c3d4_K(c, t, E, NU), g["K"])
    f, x = O.tet_surface_faces(t)
    same(f, g["surf_faces"]); same(x, g["surf_fourth"])
    same(O.tet_shared_faces(t), O.canonical_pairs(g["shared"]))
    close(O.tet_surface_normals(c, t), g["surf_normals"])
    close(O.tet_face_normals_area(c, t), g["face_normals"])
    same(O.element_to_edge(t), g["edges"])
    assert 2 * O.tet_shared_faces(t).shape[0] + f.shape[0] == 4 * t.shape[0]


def test_c3d10():
    g = load_golden("tets")
    t01 = g["tets"][:, [1, 0, 2, 3]]
    c2, e10 = O.c3d4_to_c3d10(g["coords"], t01)
    close(c2, g["coords10"], 1e-15); same(e10, g["elems10"])
    ip = g["ip10"]
    close(O.jacobian("c3d10", c2, e10, ip), g["J10"])
    close(O.shape_gradients("c3d10", c2, e10, ip), g["g10"])
    close(O.B_matrix("c3d10", c2, e10, ip), g["B10"])
    close(O.c3d10_K(c2, e10, E, NU), g["K10"])
    close(O.c3d10_K(c2, e10[:5], E, NU, single=False), g["K10_multi"])
    close(O.c3d10_K(c2, e10, E, NU, points=g["pts10_custom"]), g["K10_custom"])
    same(O.to_c3d4(e10), g["tets_from10"])
    p, w = O.c3d10_points()
    close(p, g["pts10"], 0); close(w, g["w10"], 0)
    close(O.nodal_forces(g["K10"], e10, g["u10"]), g["f10"])


def test_hexes():
    g = load_golden("hexes")
    c, h = g["coords"], g["hexes"]
    close(O.hex_volumes(c, h), g["vol"])
    close(O.jacobian("c3d8", c, h, g["ip"]), g["J"])
    close(O.shape_gradients("c3d8", c, h, g["ip"]), g["g"])
    close(O.B_matrix("c3d8", c, h, g["ip"]), g["B"])
    close(O.c3d8_K(c, h, E, NU), g["K"])
    close(O.c3d8_K(c, h[:3], E, NU, single=False), g["K_multi"])
    p, w = O.c3d8_points()
    close(p, g["pts"], 0); close(w, g["w"], 0)
    f, x = O.hex_surface_faces(h)
    same(f, g["surf_faces"]); same(x, g["surf_extra"])
    same(O.hex_shared_faces(h), O.canonical_pairs(g["shared"]))
    close(O.hex_surface_normals(c, h), g["surf_normals"])
    close(O.hex_face_normals_area(c, h), g["face_normals"])
    same(O.to_c3d4(h), g["tets"])


def test_wedges():
    g = load_golden("wedges")
    c, w6 = g["coords"], g["wedges"]
    close(O.wedge_volumes(c, w6), g["vol"])
    close(O.jacobian("c3d6", c, w6, g["ip"]), g["J"])
    close(O.shape_gradients("c3d6", c, w6, g["ip"]), g["g"])
    close(O.B_matrix("c3d6", c, w6, g["ip"]), g["B"])
    close(O.c3d6_K(c, w6, E, NU, single=True), g["K_single"])
    close(O.c3d6_K(c, w6, E, NU, single=False), g["K_full"])
    p, w = O.c3d6_points()
    close(p, g["pts"], 0); close(w, g["w"], 0)
    (q, t), (qe, te) = O.wedge_surface_faces(w6)
    same(q, g["surf_quads"]); same(t, g["surf_tris"]); same(qe, g["quad_extra"]); same(te, g["tri_extra"])
    nq, nt = O.wedge_surface_normals(c, w6)
    close(nq, g["nq"]); close(nt, g["nt"])
    same(O.to_c3d4(w6), g["tets"])


def test_shells():
    g = load_golden("shells")
    mb, bd = g["membrane"], g["bending"]
    close(O.kirchhoff_D(mb, bd), g["D"])
    c3, s3 = g["c3"], g["s3"]
    close(O.s3_unit(c3, s3), g["unit3"])
    close(O.s3_jacobian(c3, s3), g["J3"])
    close(O.s3_shape_gradient(c3, s3), g["g3"])
    close(O.s3_B(c3, s3), g["B3"])
    close(O.s3_K(c3, s3, mb, bd), g["K3"])
    same(O.tri_shared_edges(s3), O.canonical_pairs(g["shared3"]))
    e, t = O.tri_boundary_edges(s3)
    same(e, g["bedges3"]); same(t, g["bthird3"])
    c4, s4 = g["c4"], g["s4"]
    xi, eta = g["xieta"]
    close(O.s4_unit(c4, s4), g["unit4"])
    close(O.s4_jacobian(c4, s4, xi, eta), g["J4"])
    close(O.s4_shape_gradient(c4, s4, xi, eta), g["g4"])
    close(O.s4_B_single(c4, s4, xi, eta), g["B4"])
    close(O.s4_K(c4, s4, mb, bd), g["K4"])
    close(O.s4_K(c4, s4, mb, bd, single=False), g["K4_multi"])
    p, w = O.s4_points()
    close(p, g["pts4"], 0); close(w, g["w4"], 0)
    same(O.quad_shared_edges(s4), O.canonical_pairs(g["shared4"]))
    e, t = O.quad_boundary_edges(s4)
    same(e, g["bedges4"]); same(t, g["bfourth4"])
    close(O.shell_nodal_forces(g["K3"], s3, g["u3"], g["unit3"]), g["f3"])
    close(O.shell_nodal_forces(g["K4"], s4, g["u4"], g["unit4"]), g["f4"])


def test_cg_family():
    g = load_golden("solve_c3d4")
    c, t, fixed, F = g["coords"], g["tets"], g["fixed"], g["F"]
    K = O.c3d4_K(c, t, E, NU)
    u, it, st = O.stable_cg(K, t, F, fixed, tol=1e-8)
    assert st == "converged" and abs(it - int(g["it_cg"])) <= 1
    close(u, g["u_cg"], 1e-8); close(u, g["u_final"], 1e-8)
    assert int(g["it_final"]) == int(g["it_cg"])
    # the reference's column-0 "diagonal" sums to round-off noise on some dofs; compare where it is
    # a well-defined number (|1/d| modest) -- elsewhere both are 1/noise
    mine, ref = O.reference_diagonal_preconditioner(K, t, c.shape[0]), g["Minv_ref"]
    ok = (np.abs(ref) < 1e6) & (np.abs(mine) < 1e6) & (ref != 0)
    assert ok.sum() > 0.5 * ok.size
    close(mine[ok], ref[ok], 1e-9)
    Minv = O.jacobi_preconditioner(K, t, c.shape[0], fixed)
    close(Minv, g["Minv"])
    u, it, st = O.pcg_solve(lambda v: O.nodal_forces(K, t, v), F, Minv, tol=1e-8)
    assert st == "converged" and abs(it - int(g["it_pcg"])) <= 1
    close(u, g["u_pcg"], 1e-8)
    # coalesced CSR: pattern bit-exact, values 1e-12
    crow, col, val, n = O.assemble_csr(K, t, 3, c.shape[0])
    same(crow, g["crow"]); same(col, g["col"]); close(val, g["val"])
    # CSR operator == element-by-element operator, and CG through it matches
    v = np.random.default_rng(0).standard_normal(F.shape)
    close(O.csr_matvec(crow, col, val, v).reshape(-1, 3), O.nodal_forces(K, t, v), 1e-13)
    u2, it2, _ = O.cg_solve(lambda w: O.csr_matvec(crow, col, val, w).reshape(-1, 3), F, fixed, tol=1e-8)
    assert abs(it2 - int(g["it_cg"])) <= 1
    close(u2, g["u_cg"], 1e-8)


def test_static_structure_mixed():
    g = load_golden("solve_mixed")
    mat = {"E": E, "nu": NU, "membrane": (1.0, 0.3, 0.1), "bending": (1.0, 0.3, 0.1)}
    u, it, st = O.static_structure(g["coords"], g["force"], g["fixed"], c3d4=g["c3d4"], c3d6=g["c3d6"], c3d8=g["c3d8"],
                                   s3=g["s3"], s4=g["s4"], material=mat, tol=1e-8, max_iter=2000)
    assert st == "converged" and abs(it - int(g["it"])) <= 1
    close(u, g["u"], 1e-8)


def test_shell_cg():
    g = load_golden("solve_shell")
    mb = (1.0, 0.3, 0.1)
    K = O.s3_K(g["c3"], g["s3"], mb, mb)
    unit = O.s3_unit(g["c3"], g["s3"])
    u, it, st = O.cg_solve(lambda v: O.shell_nodal_forces(K, g["s3"], v, unit), g["F"], g["fixed"], tol=1e-9, max_iter=3000)
    assert st == "converged"
    # 150 unknowns, ~360 iterations: finite-precision CG on an ill-conditioned plate; the count
    # is rounding-sensitive, the solution is not
    assert abs(it - int(g["it"])) <= 25
    close(u, g["u"], 1e-6)


def test_unpinned_invariants():
    """Poisson and mass restatements have no reference counterpart: invariants only."""
    g = load_golden("tets")
    c, t = g["coords"], g["tets"]
    Kp = O.c3d4_poisson_K(c, t)
    assert np.abs(Kp.sum(axis=2)).max() < 1e-13          # constants in the null space
    assert np.abs(Kp - Kp.transpose(0, 2, 1)).max() < 1e-15
    Mm = O.c3d4_mass(c, t, 2.0)
    V = O.tet_volumes(c, t)
    close(Mm[:, 0::3, 0::3].sum(axis=(1, 2)), 2.0 * V, 1e-14)


def test_stress_recovery():
    """element stress and nodal averaging against the reference's outputs."""
    g = load_golden("stress")
    s, v = O.c3d4_element_stress(g["c"], g["t"], g["u"], E, NU)
    close(s, g["s4"]); close(v, g["v4"])
    close(O.node_vm_stress(g["c"].shape[0], g["t"], g["v4"]), g["node_vm4"])
    s, v = O.solid_element_stress("c3d10", g["c10"], g["e10"], g["u10"], E, NU)
    close(s, g["s10"]); close(v, g["v10"])
    s, v = O.solid_element_stress("c3d10", g["c10"], g["e10"][:6], g["u10"], E, NU, single=False)
    close(s, g["s10m"]); close(v, g["v10m"])
    for kind, conn, cc in (("c3d8", g["h"], g["ch"]), ("c3d6", g["w6"], g["ch"])):
        tag = kind[-1]
        s, v = O.solid_element_stress(kind, cc, conn, g["uh"], E, NU)
        close(s, g["s" + tag]); close(v, g["v" + tag])
        s, v = O.solid_element_stress(kind, cc, conn, g["uh"], E, NU, single=False)
        close(s, g["s" + tag + "m"]); close(v, g["v" + tag + "m"])


# --------------------------------------------------------------------------------------------------------------------
# C3D20 / C3D15 / consistent mass: not computable with the reference -- the oracle is pinned where
# the reference runs (27-point rule, 24-tet table) and otherwise held to invariants.
# --------------------------------------------------------------------------------------------------------------------

def _rigid_modes(c):
    n = c.shape[0]
    R = np.zeros((6, n, 3))
    R[0, :, 0] = R[1, :, 1] = R[2, :, 2] = 1
    R[3, :, 0], R[3, :, 1] = -c[:, 1], c[:, 0]
    R[4, :, 1], R[4, :, 2] = -c[:, 2], c[:, 1]
    R[5, :, 0], R[5, :, 2] = c[:, 2], -c[:, 0]
    return R


def test_c3d20_reference_pins():
    g = load_golden("quadratic")
    p, w = O.c3d20_points()
    assert np.array_equal(p, g["p20"]) and np.abs(w - g["w20"]).max() < 1e-16
    assert np.array_equal(O.to_c3d4(g["h20"]), g["h20_tets"])


@pytest.mark.parametrize("kind", ["c3d20", "c3d15"])
def test_quadratic_shape_functions(kind):
    shape = {"c3d20": O.c3d20_shape, "c3d15": O.c3d15_shape}[kind]
    nodes = O.HEX20_NAT if kind == "c3d20" else np.array(
        [(0, 0, -1), (1, 0, -1), (0, 1, -1), (0, 0, 1), (1, 0, 1), (0, 1, 1), (.5, 0, -1), (.5, .5, -1), (0, .5, -1), (.5, 0, 1),
         (.5, .5, 1), (0, .5, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0)], float)
    for a, na in enumerate(nodes):                      # Kronecker property
        N, _ = shape(na)
        assert abs(N[a] - 1) < 1e-14 and np.abs(np.delete(N, a)).max() < 1e-14
    rng = np.random.default_rng(3)
    for _ in range(5):
        p = rng.uniform(-0.6, 0.6, 3) if kind == "c3d20" else np.array([0.25, 0.3, 0.0]) + rng.uniform(-0.2, 0.2, 3)
        N, dN = shape(p)
        assert abs(N.sum() - 1) < 1e-14 and np.abs(dN.sum(axis=0)).max() < 1e-14
        h = 1e-6
        fd = np.stack([(shape(p + h * np.eye(3)[d])[0] - shape(p - h * np.eye(3)[d])[0]) / (2 * h) for d in range(3)], axis=1)
        assert np.abs(fd - dN).max() < 1e-9
        # isoparametric completeness: sum_a dN_a (x) x_a = I for the natural node coordinates themselves
        assert np.abs(dN.T @ nodes - np.eye(3)).max() < 1e-13


@pytest.mark.parametrize("kind,gen", [("c3d20", "hex20_cube"), ("c3d15", "wedge15_cube")])
def test_quadratic_stiffness_invariants(kind, gen):
    from femb200 import meshgen
    c, e = getattr(meshgen, gen)(2, jitter=0.2)
    c, e = c.numpy(), e.numpy()
    K = O.solid_K(kind, c, e, E, NU)
    scale = np.abs(K).max()
    assert np.abs(K - K.transpose(0, 2, 1)).max() < 1e-14 * scale
    for r in _rigid_modes(c):
        assert np.abs(np.einsum("mij,mj->mi", K, r[e].reshape(e.shape[0], -1))).max() < 1e-13 * scale
    ev = np.linalg.eigvalsh(K[0])
    assert (ev > -1e-12 * scale).all() and (ev > 1e-8 * scale).sum() == K.shape[1] - 6     # exactly six zero-energy modes
    p, w = O._POINTS[kind]()
    vol = sum(w[q] * np.linalg.det(O.jacobian(kind, c, e, p[q])) for q in range(len(w)))
    assert abs(vol.sum() - 1.0) < 1e-12
    # patch test: a linear displacement field leaves no residual force on interior nodes
    A = np.array([[0.01, 0.02, -0.01], [0.0, 0.03, 0.01], [0.02, -0.01, 0.015]])
    u = c @ A.T
    f = np.zeros_like(c)
    fe = np.einsum("mij,mj->mi", K, u[e].reshape(e.shape[0], -1)).reshape(-1, 3)
    np.add.at(f, e.reshape(-1), fe)
    inner = ((c > 1e-9) & (c < 1 - 1e-9)).all(axis=1)
    assert inner.any() and np.abs(f[inner]).max() < 1e-13 and np.abs(f).max() > 1e-4
    # constant-strain stress: the recovered stress equals D eps everywhere
    eps = 0.5 * (A + A.T)
    voigt = np.array([eps[0, 0], eps[1, 1], eps[2, 2], 2 * eps[0, 1], 2 * eps[1, 2], 2 * eps[0, 2]])
    S, _ = O.solid_element_stress(kind, c, e, u, E, NU, single=False)
    close(S, np.broadcast_to(O.stress_tensor((O.elasticity_matrix(E, NU) @ voigt)[None])[0], S.shape), 1e-12)


@pytest.mark.parametrize("kind,gen", [("c3d10", "tet10_cube"), ("c3d8", "hex_cube"), ("c3d6", "wedge_cube"), ("c3d20", "hex20_cube"),
                                      ("c3d15", "wedge15_cube")])
def test_consistent_mass_invariants(kind, gen):
    from femb200 import meshgen
    c, e = getattr(meshgen, gen)(2, jitter=0.15)
    c, e = c.numpy(), e.numpy()
    rho = 2.5
    Mm = O.solid_mass(kind, c, e, rho)
    assert np.abs(Mm - Mm.transpose(0, 2, 1)).max() < 1e-17 + 1e-15 * np.abs(Mm).max()
    assert np.linalg.eigvalsh(Mm[0]).min() > 0
    pm, wm = O.mass_points(kind)
    vol = sum(wm[q] * np.abs(np.linalg.det(O.jacobian(kind, c, e, pm[q]))) for q in range(len(wm)))
    assert abs(vol.sum() - 1.0) < 1e-12
    for d in range(3):
        close(Mm[:, d::3, d::3].sum(axis=(1, 2)), rho * vol, 1e-13)
    assert np.abs(Mm[:, 0::3, 1::3]).max() == 0
    if kind == "c3d10":   # straight-sided P2 tets: the degree-5 rule is exact -> closed form (V/420)[6 on corners, ...]
        c0, e0 = meshgen.tet10_cube(1)
        M0 = O.solid_mass(kind, c0.numpy(), e0.numpy(), 1.0)[0, 0::3, 0::3]
        V = 1.0 / 6
        assert abs(M0[0, 0] - 6 * V / 420) < 1e-15 and abs(M0[0, 1] - V / 420) < 1e-15 and abs(M0[4, 4] - 32 * V / 420) < 1e-15


def test_mass_rules_are_exact():
    from math import factorial
    p, w = O.mass_points("c3d10")
    for i in range(6):
        for j in range(6 - i):
            for k in range(6 - i - j):
                exact = factorial(i) * factorial(j) * factorial(k) / factorial(i + j + k + 3)
                assert abs((w * p[:, 0] ** i * p[:, 1] ** j * p[:, 2] ** k).sum() - exact) < 1e-16
    p, w = O.mass_points("c3d15")
    for i in range(5):
        for j in range(5 - i):
            for k in range(6):
                exact = factorial(i) * factorial(j) / factorial(i + j + 2) * (0 if k % 2 else 2 / (k + 1))
                assert abs((w * p[:, 0] ** i * p[:, 1] ** j * p[:, 2] ** k).sum() - exact) < 1e-15


def test_constrained_cg():
    """SPC / RBE2 / RBE3 constrained CG loops against the reference's outputs."""
    import json
    g = load_golden("constrained")
    C = json.loads(str(g["constraints_json"]))
    sp, r2, r3 = O.parse_spc(C["spc"]), O.parse_rbe2(C["rbe2"]), O.parse_rbe3(C["rbe3"])
    for a, b in zip(sp + r2 + r3, [g[k] for k in ("spc_n", "spc_d", "spc_v", "r2_s", "r2_m", "r2_d", "r3_m", "r3_s", "r3_d", "r3_w", "r3_i", "r3_ws")]):
        assert np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))
    K = O.c3d4_K(g["coords"], g["tets"], E, NU)
    ap = lambda v: O.nodal_forces(K, g["tets"], v)  # noqa: E731
    u, it, st = O.constrained_cg(ap, g["F"], C["rbe2"], C["spc"], tol=1e-9, max_iter=2000)
    assert st == "converged" and abs(it - int(g["it_c"])) <= 1
    close(u, g["u_c"], 1e-8)
    u, it, st = O.constrained_cg(ap, g["F"], C["rbe2"], C["spc"], rbe3_list=C["rbe3"], tol=1e-9, max_iter=2000)
    assert st == "converged" and abs(it - int(g["it_n"])) <= 1
    close(u, g["u_n"], 1e-8)
    u, it, st = O.constrained_cg(ap, g["F"], C["rbe2"], C["spc"], u_init=g["u0"], tol=1e-9, max_iter=2000)
    assert st == "converged" and abs(it - int(g["it_c0"])) <= 1
    close(u, g["u_c0"], 1e-8)
    # the product's host-side parsers (no GPU needed) produce the reference's tensors
    import sys
    from conftest import PKG
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import solver as sv
    got = sv.parse_spc_list(C["spc"], device="cpu") + sv.parse_rbe2_list(C["rbe2"], device="cpu") + sv.parse_rbe3_list(C["rbe3"], device="cpu")
    keys = ("spc_n", "spc_d", "spc_v", "r2_s", "r2_m", "r2_d", "r3_m", "r3_s", "r3_d", "r3_w", "r3_i", "r3_ws")
    for t, k in zip(got, keys):
        assert np.array_equal(t.numpy().astype(np.float64), g[k].astype(np.float64)), k
    assert got[0].dtype.is_floating_point is False and str(got[0].dtype) == "torch.int32" and str(got[10].dtype) == "torch.int64"


def test_partition_oracle():
    """subdivision.ipynb cells 7-9: the oracle reproduces the notebook code's partition exactly."""
    g = load_golden("partition")
    lab, seeds = O.region_growing_partition(g["edge"], 5, g["tets"].shape[0], int(g["first"]))
    assert np.array_equal(seeds, g["seeds"]) and np.array_equal(lab, g["labels"])
    assert [O.compute_subdivisions(338619, 10), O.compute_subdivisions(1000, 1), O.compute_subdivisions(3000000, 4)] == g["subdiv"].tolist()
    from conftest import PKG
    import sys
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import subdivision as sd
    assert sd.compute_subdivisions(338619, 10) == int(g["subdiv"][0])
    nm = [np.array([0, 1, 2, 5]), np.array([2, 3, 5]), np.array([5, 6])]
    import torch
    assert sd.build_ordered_subdomain_map([torch.tensor(a) for a in nm]) == {(0, 1): [2], (0, 1, 2): [5]}
