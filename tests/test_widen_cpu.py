"""CPU checks for the rows either side of the element path (SURVEY section 8a complements a14/a15/a18/a23 and 8f #1/#3):
the oracle restatements against reference-generated outputs (tests/golden/widen.npz), and the library's host-side
legacy-VTK parser (no GPU involved) against files written here and against the oracle's pure-python parser."""
import ctypes
import os
import struct

import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import fem_oracle as O

TOL = 1e-12


def close(a, b, tol=TOL):
    assert np.asarray(a).shape == np.asarray(b).shape, (np.asarray(a).shape, np.asarray(b).shape)
    assert rel_err(a, b) <= tol, rel_err(a, b)


def test_shell_frames_stress_postprocess_oracle():
    d, g = load_golden("widen"), load_golden("shells")
    c3, s3, c4, s4 = g["c3"], g["s3"], g["c4"], g["s4"]
    close(O.shell_normal(c3, s3), d["normal3"])
    close(O.shell_normal(c4, s4), d["normal4"])
    # the global-axes element operator reproduces the reference's rotate / apply / rotate-back force routine
    Kg = O.shell_global_K(g["K3"], g["unit3"])
    dofs = (s3[:, :, None] * 6 + np.arange(6)).reshape(s3.shape[0], -1)
    f = np.bincount(dofs.reshape(-1), weights=np.einsum("mij,mj->mi", Kg, g["u3"].reshape(-1)[dofs]).reshape(-1), minlength=g["u3"].size)
    close(f.reshape(-1, 6), g["f3"])
    close(O.shell_local_coords(c3, s3, O.s3_unit(c3, s3)), d["loc3"])
    close(O.shell_local_coords(c4, s4, O.s4_unit(c4, s4)), d["loc4"])
    close(O.shell_local_displacement(s3, d["u3"], O.s3_unit(c3, s3)), d["ul3"])
    close(O.shell_local_displacement(s4, d["u4"], O.s4_unit(c4, s4)), d["ul4"])
    close(O.s4_B(c4, s4), d["B4_sum"])
    close(O.s4_B(c4, s4, single=False), d["B4_all"])
    p, _ = O.s4_points()
    close(O.s4_jacobian_t(c4, s4, p[0, 0], p[0, 1]), d["J4_t"])
    close(O.s4_shape_gradient_t(c4, s4, p[0, 0], p[0, 1]), d["g4_t"])
    close(O.s3_shell_stress(c3, s3, g["membrane"], g["bending"], d["u3"]), d["stress3"])
    close(O.s4_shell_stress(c4, s4, g["membrane"], g["bending"], d["u4"]), d["stress4"])
    t, z = d["post_tz"]
    assert tuple(d["post_keys"]) == O.POST_KEYS
    close(O.shell_postprocess(d["stress4"], t, z), d["post"])
    close(O.shell_postprocess(d["stress4"], t, z, np.float32), d["post32"], 1e-6)


def test_shell_extrude_oracle():
    d = load_golden("widen")
    th = float(d["thickness"])
    x, w6, h8 = O.shell_extrude(d["cm"], d["tri"], d["quad"], th)
    close(x, d["ext_x"])
    assert np.array_equal(w6, d["ext_w"]) and np.array_equal(h8, d["ext_h"])
    close(O.shell_extrude(d["cm"], d["tri"], d["quad"], th, np.float32)[0], d["ext_x32"], 1e-6)
    close(O.shell_extrude(d["cm"], d["tri"], d["quad"][:0], th)[0], d["ext_x_tri_only"])
    close(O.shell_extrude(d["cm"], d["tri"][:0], d["quad"], th)[0], d["ext_x_quad_only"])
    # nodes without elements keep their position in both layers (0/(0+eps) normal)
    N = d["cm"].shape[0]
    lone = np.setdiff1d(np.arange(N), d["tri"].reshape(-1))
    assert lone.size and np.array_equal(d["ext_x_tri_only"][lone], d["cm"][lone]) and np.array_equal(d["ext_x_tri_only"][lone + N], d["cm"][lone])
    # the extruded solids are positively oriented for the reference's wedge / hex Jacobians
    assert (np.linalg.det(O.jacobian("c3d6", x, w6, np.array([1 / 3, 1 / 3, 0.0]))) > 0).all()
    assert (np.linalg.det(O.jacobian("c3d8", x, h8, np.zeros(3))) > 0).all()


def test_face_forces_and_operator_cg_oracle():
    d = load_golden("widen")
    nrm = O.tet_face_normals_area(d["ct"], d["tets"])
    close(O.c3d4_surface_forces(nrm, d["sigma"]), d["face_forces"])
    close(O.c3d4_shared_face_forces_sum(d["shared"], d["face_forces"]), d["shared_sum"])
    # a uniform stress field is in equilibrium across every shared face
    uni = np.broadcast_to(np.array([[1.0, 0.2, 0.0], [0.2, -0.5, 0.3], [0.0, 0.3, 2.0]]), (d["tets"].shape[0], 3, 3))
    bal = O.c3d4_shared_face_forces_sum(d["shared"], O.c3d4_surface_forces(nrm, uni))
    assert np.abs(bal).max() < 1e-14
    K = O.c3d4_K(d["ct"], d["tets"], 1.0, 0.3)
    shift = float(d["ku_shift"])
    u, its = O.cg_solve_Ku(lambda v: O.nodal_forces(K, d["tets"], v) + shift * v, d["ku_R"], tol=1e-10, max_iter=500)
    assert its < 500
    close(u, d["ku_u"], 1e-10)


def test_wedge_face_normals_invariants():
    """The reference's compute_wedge_normals_and_area raises on every input (recorded in the fixture); the restated intent is
    checked by invariants: unit length, orthogonal to both generating edges, right-handed w.r.t. them."""
    d = load_golden("widen")
    assert int(d["wedge_normals_raises"]) == 1
    n = O.wedge_face_normals(d["cw"], d["w"])
    x = d["cw"][d["w"]]
    tab = [(0, 1, 3), (1, 2, 4), (2, 0, 5), (0, 2, 1), (3, 4, 5)]
    assert np.abs(np.linalg.norm(n, axis=-1) - 1).max() < 1e-14
    for f, (o, a, b) in enumerate(tab):
        e1, e2 = x[:, a] - x[:, o], x[:, b] - x[:, o]
        assert np.abs((n[:, f] * e1).sum(-1)).max() < 1e-14 and np.abs((n[:, f] * e2).sum(-1)).max() < 1e-14
        assert ((np.cross(e1, e2) * n[:, f]).sum(-1) > 0).all()


def test_s4_factor_rows_follow_the_argument_type():
    """Host logic of the S4 wrappers: python floats give 1 -/+ x in double, 0-dim float32 tensors give it rounded to float32
    first (what the reference's own arithmetic does when compute_s4_B_matrix walks its float32 rule)."""
    import sys
    import torch
    from conftest import PKG
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import shell as sh
    g32 = np.float32(1.0) / np.sqrt(np.float32(3.0))
    row = sh._s4_factor_row(torch.tensor(g32), torch.tensor(-g32), torch.tensor(1.0))
    one = np.float32(1.0)
    assert row == [float(one - g32), float(one + g32), float(one + g32), float(one - g32), 1.0]
    row = sh._s4_factor_row(0.3, -0.6)
    assert row == [1 - 0.3, 1 + 0.3, 1 + 0.6, 1 - 0.6, 1.0]
    x = 0.1234567891234
    assert sh._s4_factor_row(torch.tensor(x, dtype=torch.float32), 0.0)[1] == float(np.float32(1) + np.float32(x)) != 1 + x
    rows = sh._s4_rule_rows(None)
    p, w = O.s4_points()
    assert len(rows) == 4 and [r[4] for r in rows] == [1.0] * 4
    for q in range(4):
        dxi, deta = O._s4_d32(p[q, 0], p[q, 1])
        assert np.array_equal(0.25 * np.array([-rows[q][2], rows[q][2], rows[q][3], -rows[q][3]]), dxi)
        assert np.array_equal(0.25 * np.array([-rows[q][0], -rows[q][1], rows[q][1], rows[q][0]]), deta)


def load_groups(d):
    """{sorted tuple of parts: [nodes]} in the fixture's order."""
    out, off = {}, 0
    for row, c in zip(d["keys"], d["counts"]):
        out[tuple(int(v) for v in row if v >= 0)] = [int(v) for v in d["nodes"][off:off + c]]
        off += c
    return out


def test_subdomain_forces_oracle():
    """subdivision.ipynb cell 15 run as written (fixture) against the oracle, bit for bit (only additions and subtractions)."""
    d = load_golden("subdomain_forces")
    g2n = load_groups(d)
    s = O.sub_domain_forces(d["fv"], g2n, d["rbe2"], d["F"], 5)
    assert np.array_equal(np.stack(s), d["out"])
    # interface forces cancel in the sum over subdomains; fixed interface nodes and interior nodes keep F everywhere
    assert np.abs(np.stack(s).sum(0) - 5 * d["F"]).max() < 1e-13
    for nd in d["rbe2"]:
        assert all(np.array_equal(si[nd], d["F"][nd]) for si in s)
    # the index form the product builds on the host addresses every (subdomain, interface node) once
    import sys
    from conftest import PKG
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import subdivision as sd
    plan = sd.SubdomainForcePlan(g2n, d["rbe2"], d["F"].shape[0], 5, "cpu")
    assert plan.n_free == d["fv"].shape[0] == sd.count_free_variables(g2n)
    free_nodes = [n for v in g2n.values() for n in v if n not in set(d["rbe2"].tolist())]
    assert plan.tgt.numel() == sum(len(k) for k, v in g2n.items() for n in v if n in free_nodes)
    ref = np.broadcast_to(d["F"], (5,) + d["F"].shape).copy().reshape(-1, 3)
    tgt, plus, minus = plan.tgt.numpy(), plan.plus.numpy(), plan.minus.numpy()
    for t, a, b in zip(tgt, plus, minus):
        f = d["F"][t % d["F"].shape[0]]
        ref[t] = f + (d["fv"][a] - d["fv"][b]) if a >= 0 and b >= 0 else (f + d["fv"][a] if a >= 0 else f - d["fv"][b])
    assert np.array_equal(ref.reshape(5, -1, 3), d["out"])
    maps, off = [], 0
    import torch
    for n in d["node_maps_len"]:
        maps.append(torch.tensor(d["node_maps_flat"][off:off + n]))
        off += n
    got = sd.build_ordered_subdomain_map(maps)
    assert {k: sorted(v) for k, v in got.items()} == g2n
    with pytest.raises(ValueError):
        sd.SubdomainForcePlan({(0, 1): [3], (1, 2): [3]}, [], 10, 3, "cpu")


def test_modal_solver_oracle_and_host_steps():
    """The reference's vectorized_modal_solver raises on every input (recorded in the fixture).  The oracle restates its steps;
    with the small pencil solved as a symmetric-definite one the iteration converges to the largest eigenvalues of the
    constrained problem (dense scipy solve), and the product's host-side k x k routines agree with the oracle's."""
    import scipy.linalg as sl
    d = load_golden("modal")
    assert "single memory location" in str(d["reference_raises"])
    N = d["coords"].shape[0]
    K = O.c3d4_K(d["coords"], d["tets"], 1.0, 0.3)
    lam, modes = O.modal_solver(K, d["Mloc"], d["tets"], d["fixed"], N, d["X0"], max_iter=300, as_written=False)
    e = d["tets"]
    dofs = (e[:, :, None] * 3 + np.arange(3)).reshape(-1)
    Md = np.bincount(dofs, weights=np.diagonal(d["Mloc"], axis1=1, axis2=2).reshape(-1), minlength=3 * N)
    A = np.stack([O.nodal_forces(K, e, np.eye(3 * N)[c].reshape(N, 3)).reshape(-1) for c in range(3 * N)], axis=1)
    fix = (d["fixed"].reshape(-1, 1) * 3 + np.arange(3)).reshape(-1)
    free = np.setdiff1d(np.arange(3 * N), fix)
    exact = sl.eigh(A[np.ix_(free, free)], np.diag(Md[free]), eigvals_only=True)
    assert np.abs(lam - exact[-4:]).max() < 1e-4 * exact[-1]
    v = modes[:, -1]
    Kv = A @ v
    Kv[fix] = 0
    assert np.linalg.norm(Kv - lam[-1] * Md * v) < 1e-3 * np.linalg.norm(Kv) and np.abs(modes[fix]).max() == 0
    # host-side small-matrix routines of the product against the oracle's restatement
    import sys
    import torch
    from conftest import PKG
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import solver as sv
    rng = np.random.default_rng(4)
    for k in (2, 4, 5):
        B = rng.standard_normal((k, k))
        B = B @ B.T + k * np.eye(k)
        Am = rng.standard_normal((k, k))
        Am = Am + Am.T
        close(sv._invert_small_matrix(B.copy()), O._gj_inverse(B))
        close(sv._invert_small_matrix(B.copy()), np.linalg.inv(B), 1e-10)
        l1, Z1 = sv._solve_small_gevp(torch.tensor(Am), torch.tensor(B), np.float64)
        l2, Z2 = O._jacobi_as_written(O._gj_inverse(B) @ Am)
        close(l1, l2)
        close(Z1, Z2)
        l3, Z3 = sv._solve_small_gevp_sym(torch.tensor(Am), torch.tensor(B))
        close(l3, sl.eigh(Am, B, eigvals_only=True), 1e-10)
        close(Z3.T @ B @ Z3, np.eye(k), 1e-10)
    # the reference's rotation formulas are not a valid Jacobi method even for symmetric input (2x2: one "sweep" does not
    # diagonalise), which is why the product's default solves the small pencil with a proper symmetric eigen-solver
    S2 = np.array([[2.0, 1.0], [1.0, -1.0]])
    ls, _ = sv._naive_jacobi(S2.copy(), max_sweeps=200, tol=1e-14)
    assert np.abs(ls - np.linalg.eigvalsh(S2)).max() > 1e-3


# ------------------------------------------------------------------------------------------------ legacy VTK files

def _mesh():
    rng = np.random.default_rng(3)
    pts = rng.standard_normal((9, 3))
    tets = np.array([[0, 1, 2, 3], [1, 2, 3, 4], [4, 5, 6, 7], [5, 6, 7, 8]], dtype=np.int64)
    return pts, tets


def _write_ascii(path, pts, cells, vtk_type, real="double", crlf=False):
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        f.write("# vtk DataFile Version 3.0" + nl + "written by tests" + nl + "ASCII" + nl + "DATASET UNSTRUCTURED_GRID" + nl)
        f.write(f"POINTS {len(pts)} {real}" + nl)
        for p in pts:
            f.write(" ".join(repr(float(v)) for v in p) + nl)
        f.write(f"CELLS {len(cells)} {cells.size + len(cells)}" + nl)
        for c in cells:
            f.write(f"{len(c)} " + " ".join(str(int(v)) for v in c) + nl)
        f.write(f"CELL_TYPES {len(cells)}" + nl)
        f.write(nl.join(str(vtk_type) for _ in cells) + nl)
        f.write(f"CELL_DATA {len(cells)}" + nl + "SCALARS mat int 1" + nl + "LOOKUP_TABLE default" + nl + nl.join("7" for _ in cells) + nl)


def _write_binary(path, pts, cells, vtk_type, real="float", v5=False):
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version " + (b"5.1" if v5 else b"4.2") + b"\nwritten by tests\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {len(pts)} {real}\n".encode())
        f.write(pts.astype(">f4" if real == "float" else ">f8").tobytes() + b"\n")
        if v5:
            off = np.arange(len(cells) + 1, dtype=np.int64) * cells.shape[1]
            f.write(f"CELLS {len(cells) + 1} {cells.size}\nOFFSETS vtktypeint64\n".encode() + off.astype(">i8").tobytes() + b"\n")
            f.write(b"CONNECTIVITY vtktypeint64\n" + cells.astype(">i8").tobytes() + b"\n")
        else:
            flat = np.concatenate([np.full((len(cells), 1), cells.shape[1]), cells], axis=1)
            f.write(f"CELLS {len(cells)} {flat.size}\n".encode() + flat.astype(">i4").tobytes() + b"\n")
        f.write(f"CELL_TYPES {len(cells)}\n".encode() + np.full(len(cells), vtk_type, ">i4").tobytes() + b"\n")


@pytest.mark.parametrize("variant", ["ascii", "ascii_float_crlf", "binary_float", "binary_double", "binary_v5", "ascii_v5"])
def test_vtk_host_parser(tmp_path, variant):
    from femb200 import ops
    pts, tets = _mesh()
    path = str(tmp_path / f"{variant}.vtk")
    expect = pts
    if variant == "ascii":
        _write_ascii(path, pts, tets, 10)
    elif variant == "ascii_float_crlf":
        _write_ascii(path, pts, tets, 10, real="float", crlf=True)
        expect = pts.astype(np.float32).astype(np.float64)
    elif variant == "binary_float":
        _write_binary(path, pts, tets, 10, real="float")
        expect = pts.astype(np.float32).astype(np.float64)
    elif variant == "binary_double":
        _write_binary(path, pts, tets, 10, real="double")
    elif variant == "binary_v5":
        _write_binary(path, pts, tets, 10, real="double", v5=True)
    else:
        with open(path, "w") as f:
            f.write("# vtk DataFile Version 5.1\nt\nASCII\nDATASET UNSTRUCTURED_GRID\n")
            f.write(f"POINTS {len(pts)} double\n" + "\n".join(" ".join(repr(float(v)) for v in p) for p in pts) + "\n")
            f.write(f"CELLS {len(tets) + 1} {tets.size}\nOFFSETS vtktypeint64\n" + " ".join(str(4 * k) for k in range(len(tets) + 1)) + "\n")
            f.write("CONNECTIVITY vtktypeint64\n" + " ".join(str(int(v)) for v in tets.reshape(-1)) + "\n")
            f.write(f"CELL_TYPES {len(tets)}\n" + "\n".join("10" for _ in tets) + "\n")
    p, cells, types, is_float = ops.vtk_read(path)
    flat = np.concatenate([np.full((len(tets), 1), 4), tets], axis=1).reshape(-1)
    assert np.array_equal(p, expect) and np.array_equal(cells, flat) and np.array_equal(types, np.full(len(tets), 10))
    assert is_float == ("float" in variant)
    po, co, to = O.vtk_read_legacy(path)
    assert np.array_equal(po, p) and np.array_equal(co, cells) and np.array_equal(to, types)


def test_vtk_host_parser_errors(tmp_path):
    from femb200 import _lib, ops
    pts, tets = _mesh()
    with pytest.raises(_lib.FembError, match="cannot open"):
        ops.vtk_read(str(tmp_path / "missing.vtk"))
    bad = tmp_path / "xml.vtk"
    bad.write_text("<?xml version=\"1.0\"?>\n<VTKFile/>\n")
    with pytest.raises(_lib.FembError, match="legacy VTK"):
        ops.vtk_read(str(bad))
    poly = tmp_path / "poly.vtk"
    poly.write_text("# vtk DataFile Version 3.0\nt\nASCII\nDATASET POLYDATA\nPOINTS 0 float\n")
    with pytest.raises(_lib.FembError, match="UNSTRUCTURED_GRID"):
        ops.vtk_read(str(poly))
    trunc = tmp_path / "trunc.vtk"
    _write_binary(str(trunc), pts, tets, 10, real="double")
    data = trunc.read_bytes()
    trunc.write_bytes(data[:200])
    with pytest.raises(_lib.FembError, match="truncated"):
        ops.vtk_read(str(trunc))
    oob = tmp_path / "oob.vtk"
    _write_ascii(str(oob), pts, tets + 6, 10)
    with pytest.raises(_lib.FembError, match="outside POINTS"):
        ops.vtk_read(str(oob))
    wrong = tmp_path / "count.vtk"
    wrong.write_text("# vtk DataFile Version 3.0\nt\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 4 double\n0 0 0 1 0 0 0 1 0 0 0 1\n"
                     "CELLS 1 6\n4 0 1 2 3 0\nCELL_TYPES 1\n10\n")
    with pytest.raises(_lib.FembError, match="per-cell counts"):
        ops.vtk_read(str(wrong))
    empty = tmp_path / "empty.vtk"
    empty.write_text("# vtk DataFile Version 3.0\nt\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS 0 double\nCELLS 0 0\nCELL_TYPES 0\n")
    p, c, t, _ = ops.vtk_read(str(empty))
    assert p.shape == (0, 3) and c.size == 0 and t.size == 0


def test_operator_cg_signature_without_gpu():
    """The callback type of femb_cg_solve_operator binds, and argument validation runs without a device."""
    from femb200 import _lib
    cb = _lib.APPLY_FN(lambda ctx, x, y, s: 0)
    res = _lib.CGResult()
    rc = _lib.lib.femb_cg_solve_operator(0, cb, None, None, None, None, 1e-8, 10, 8, ctypes.byref(res), None)
    assert rc == 1 and b"n <= 0" in _lib.lib.femb_last_error()
    assert _lib.lib.femb_shell_ex(103, 6, None, 8, None, 8, 0, None, 1, None, None, None, None) == 1   # what = 6 is not a shell_ex code
    assert _lib.lib.femb_shell_postprocess(None, 8, 0, 5, 0.1, 0.0, None, None) == 1                    # needs >= 6 columns
