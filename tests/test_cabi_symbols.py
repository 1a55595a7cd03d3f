"""CPU-side checks of the boundary: the shared library loads without a GPU and exports every symbol that
include/femb200.h declares; argument validation works without touching the device."""
import ctypes
import os
import re

from conftest import PKG, ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "femb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(femb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(PKG, "libfemb200.so"))
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/femb200.h but not exported"


def test_binding_table_matches_header():
    from femb200 import _lib
    assert set(_lib.SIGNATURES) | {"femb_last_error"} == set(declared_symbols())


def test_argument_validation_without_gpu():
    from femb200 import _lib
    rc = _lib.lib.femb_c3d4(2, None, 3, None, 8, 0, 1.0, 0.3, None, None, None)   # fp=3 is invalid
    assert rc == 1 and b"fp" in _lib.lib.femb_last_error()
    buf = (ctypes.c_double * 256)()
    assert _lib.lib.femb_default_points(10, buf) == 11 and abs(sum(buf[4 * q + 3] for q in range(11)) - 0.45) < 1e-15
    assert _lib.lib.femb_default_points(6, buf) == 6 and buf[2] == -0.5773502588272095   # fp32-rounded (quirk q1)


def test_round2_entry_points_validate_arguments_without_gpu():
    """femb_p2_* (c3d4_to_c3d10 kernels), femb_dist_cg_solve's `block` argument and femb_entities_create reject bad arguments before
    touching the device; the additive keyword arguments of the mirror keep the reference's positional order intact."""
    import inspect
    import sys
    from femb200 import _lib
    L = _lib.lib
    h, ne = ctypes.c_void_p(), ctypes.c_int64()
    assert L.femb_p2_create(None, 3, 0, 1, None, ctypes.byref(h), ctypes.byref(ne)) == 1 and b"ib" in L.femb_last_error()
    assert L.femb_p2_create(None, 8, 0, 0, None, ctypes.byref(h), ctypes.byref(ne)) == 1          # n_nodes >= 1
    assert L.femb_p2_create(None, 8, 400_000_000, 10, None, ctypes.byref(h), ctypes.byref(ne)) == 1 and b"int32" in L.femb_last_error()
    assert L.femb_p2_create(None, 8, 0, 5, None, ctypes.byref(h), ctypes.byref(ne)) == 0 and ne.value == 0   # empty mesh: no device work
    assert L.femb_p2_destroy(h) == 0
    res = _lib.CGResult()
    sym = (ctypes.c_void_p * 2)()
    rc = L.femb_dist_cg_solve(0, 2, 9, 0, 1, None, None, None, None, None, None, None, None, sym, 0, None, None, None, None, None, None, None,
                              1e-8, 10, 1e-30, 16, 2, ctypes.byref(res), None)
    assert rc == 1 and b"block" in L.femb_last_error()
    rc = L.femb_dist_cg_solve(0, 2, 10, 0, 1, None, None, None, None, None, None, None, None, sym, 0, None, None, None, None, None, None, None,
                              1e-8, 10, 1e-30, 16, 3, ctypes.byref(res), None)
    assert rc == 1                                                                                     # 10 rows are not 3 per node
    rc = L.femb_dist_cg_solve(5, 2, 9, 0, 1, None, None, None, None, None, None, None, None, sym, 0, None, None, None, None, None, None, None,
                              1e-8, 10, 1e-30, 16, 1, ctypes.byref(res), None)
    assert rc == 1 and b"rank" in L.femb_last_error()
    assert L.femb_dist_header_bytes() % 256 == 0
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import element, solver
    sig = inspect.signature(solver.stable_conjugate_gradient_solver)
    assert list(sig.parameters)[:10] == ["K", "elements", "F", "rbe2", "u_init", "tol", "max_iter", "device", "dtype", "eps"]
    assert sig.parameters["distributed"].default is False and sig.parameters["coords"].default is None
    sig = inspect.signature(solver.preconditioned_conjugate_gradient_solver)
    assert list(sig.parameters)[:9] == ["K", "elements", "F", "M_inv", "u_init", "tol", "max_iter", "device", "dtype"]
    assert sig.parameters["tol"].default == 1e-8 and sig.parameters["distributed"].default is False
    assert list(inspect.signature(element.c3d4_to_c3d10).parameters)[:5] == ["coords", "elements", "rbe2_ids", "rbe3_ids", "dtype"]
    assert inspect.signature(element.compute_c3d10_K_matrix).parameters["out"].default is None
    assert inspect.signature(solver.hybrid_subdivided_solver).parameters["mode"].default == "multilevel"


def test_mirror_api_names():
    """Every hot-path function of the reference's three modules exists with the reference's parameter names."""
    import inspect
    import sys
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import element, shell, solver
    expect = {
        element: ["compute_elasticity_matrix", "to_c3d4", "to_2nd_order", "integral_points", "compute_Jacobian", "compute_shape_gradients",
                  "compute_B_matrix", "compute_K_matrix", "compute_nodal_forces", "compute_tetrahedral_volumes",
                  "compute_tetrahedral_surface_faces_with_fourth_node", "compute_tetrahdral_surface_normals",
                  "compute_tetrahedral_normals_and_area", "identify_tetrahedral_shared_faces", "c3d4_to_c3d10", "compute_c3d4_B_matrix",
                  "compute_c3d4_K_matrix", "c3d10_to_c3d4", "c3d10_integration_points", "compute_c3d10_Jacobian",
                  "compute_c3d10_shape_gradients", "compute_c3d10_B_matrix", "compute_c3d10_K_matrix", "compute_hexahedral_volumes",
                  "compute_hexahedral_surface_faces_with_extra_node", "compute_hexahedral_surface_normals",
                  "compute_hexahedral_normals_and_area", "identify_hexahedral_shared_faces", "c3d8_to_c3d4", "c3d8_integration_points",
                  "compute_c3d8_Jacobian", "compute_c3d8_shape_gradients", "compute_c3d8_B_matrix", "compute_c3d8_K_matrix",
                  "compute_wedge_volumes", "compute_wedge_surface_faces_with_extra_node", "compute_wedge_surface_normals", "c3d6_to_c3d4",
                  "c3d6_integration_points", "compute_c3d6_Jacobian", "compute_c3d6_shape_gradients", "compute_c3d6_B_matrix",
                  "compute_c3d6_K_matrix", "element_to_edge", "compute_wedge_normals_and_area", "compute_c3d4_surface_forces",
                  "compute_c3d4_shared_face_forces_sum", "vtk_loader_to_torch", "compute_element_stress", "compute_node_vm_stress"],
        shell: ["compute_kirchoff_D_matrix", "compute_shell_nodal_forces", "identify_s3_shared_edges",
                "compute_triangle_surface_faces_with_third_node", "compute_s3_local_unitvector", "compute_s3_jacobian",
                "compute_s3_shape_gradient", "compute_s3_B_matrix", "compute_s3_K_matrix", "identify_s4_shared_edges",
                "compute_square_surface_faces_with_fourth_node", "compute_s4_local_unitvector", "s4_integration_points",
                "compute_s4_jacobian", "compute_s4_shape_gradient", "compute_s4_B_matrix_single", "compute_s4_K_matrix",
                "compute_global_to_local_displacement", "compute_shell_postprocess_values", "compute_s3_global_to_local_coordinates",
                "compute_s4_global_to_local_coordinates", "compute_s4_B_matrix", "compute_s3_shell_stress", "compute_s4_shell_stress",
                "shell_extrude"],
        solver: ["static_structure_solver", "stable_conjugate_gradient_solver", "final_solver", "stable_conjugate_gradient_shell_solver",
                 "preconditioned_conjugate_gradient_solver", "compute_diagonal_preconditioner", "compute_K_matrix",
                 "conjugate_gradient_solver_Ku", "vectorized_modal_solver", "constrained_conjugate_gradient_solver", "new_constrained_conjugate_gradient_solver"],
    }
    for mod, names in expect.items():
        for n in names:
            assert callable(getattr(mod, n)), n
    sig = inspect.signature(solver.stable_conjugate_gradient_solver)
    assert list(sig.parameters)[:10] == ["K", "elements", "F", "rbe2", "u_init", "tol", "max_iter", "device", "dtype", "eps"]
    assert sig.parameters["tol"].default == 1e-10 and sig.parameters["eps"].default == 1e-30
    assert [element.human_readable_number(v) for v in (5, 1000, 63888000, -2.5e9, 1.6e13, 3e15, 7e18)] == \
        ["5.0", "1.0K", "63.9M", "-2.5B", "16.0T", "3.0Quad", "7.0Quint"]
    sig = inspect.signature(solver.conjugate_gradient_solver_Ku)
    assert list(sig.parameters)[:6] == ["compute_Ku", "R", "tol", "max_iter", "device", "dtype"] and sig.parameters["tol"].default == 1e-8
    assert list(inspect.signature(shell.shell_extrude).parameters) == ["coords", "tri", "quad", "thickness", "device", "dtype"]
    assert list(inspect.signature(shell.compute_shell_postprocess_values).parameters) == ["NMQ", "t", "z", "device", "dtype"]
    assert list(inspect.signature(element.vtk_loader_to_torch).parameters) == ["file_path", "element_type", "device", "dtype"]
    sig = inspect.signature(element.compute_K_matrix)
    assert list(sig.parameters) == ["coords", "elements", "element_type", "E", "nu", "integral_point", "single", "device", "dtype"]


def test_host_tables_match_oracle():
    """Host-side tables of the library (no GPU needed): default / mass rules and the natural-coordinate shape tables the
    solid kernels consume, against the oracle's independent restatement."""
    import numpy as np
    from femb200 import _lib
    from oracle import fem_oracle as O
    kinds = {"c3d10": 10, "c3d8": 8, "c3d6": 6, "c3d20": 20, "c3d15": 15}
    buf = (ctypes.c_double * 256)()
    for name, k in kinds.items():
        n = _lib.lib.femb_default_points(k, buf)
        p, w = O._POINTS[name]()
        got = np.array(buf[:4 * n]).reshape(n, 4)
        assert n == len(w) and np.array_equal(got[:, :3], p) and np.abs(got[:, 3] - w).max() < 1e-16, name
        n = _lib.lib.femb_mass_points(k, buf)
        p, w = O.mass_points(name)
        got = np.array(buf[:4 * n]).reshape(n, 4)
        assert n == len(w) and np.abs(got[:, :3] - p).max() < 1e-16 and np.abs(got[:, 3] - w).max() < 1e-17, name
    rng = np.random.default_rng(5)
    for name, k in kinds.items():
        nq = 7
        pts = rng.uniform(-0.7, 0.7, (nq, 4)) if name in ("c3d8", "c3d20") else np.abs(rng.uniform(0.05, 0.3, (nq, 4)))
        Nh = (ctypes.c_double * (nq * k))()
        dNh = (ctypes.c_double * (nq * k * 3))()
        rc = _lib.lib.femb_shape_tables(k, pts.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nq, Nh, dNh)
        assert rc == 0
        Ng, dNg = np.array(Nh[:]).reshape(nq, k), np.array(dNh[:]).reshape(nq, k, 3)
        for q in range(nq):
            assert np.abs(O._N[name](pts[q, :3]) - Ng[q]).max() < 1e-14, name
            assert np.abs(O._DN[name][0](pts[q, :3]) - dNg[q]).max() < 1e-14, name


def test_notebook_call_sites_are_covered():
    """Every function of its own modules / cells that the reference's two notebooks call (fixture written by
    tests/golden/make_golden.py from the notebooks) exists in the mirror."""
    import json
    import sys
    sys.path.insert(0, os.path.join(PKG, "solver"))
    import element, shell, solver, subdivision   # noqa: E401
    calls = json.load(open(os.path.join(ROOT, "tests", "golden", "notebook_calls.json")))
    for n in calls["module_functions"]:
        assert callable(getattr(solver, n, None)), n      # solver re-exports element and shell, as the reference's does
    for n in calls["notebook_functions"]:
        assert callable(getattr(subdivision, n, None)), n
    assert len(calls["module_functions"]) >= 7 and len(calls["notebook_functions"]) >= 9
