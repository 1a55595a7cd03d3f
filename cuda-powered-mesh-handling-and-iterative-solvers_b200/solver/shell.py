"""Drop-in mirror of the reference's `solver/shell.py` (Kirchhoff S3/S4 shells) on sm_100a kernels.

Same names, argument order, defaults and shapes as the reference; CUDA only, no torch op chains.
Docstrings cite the reference lines (relative to its solver/ directory)."""
from __future__ import annotations

import os
import sys

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from femb200 import ops as _ops  # noqa: E402


def compute_kirchoff_D_matrix(membrane, bending, device="cuda:0", dtype=torch.float32):
    """6x6 membrane+bending constitutive matrix from (E, nu, t) triples (shell.py:15-39)."""
    Em, num, tm = (float(v) for v in membrane)
    Eb, nub, tb = (float(v) for v in bending)
    a = Em * tm / (1 - num ** 2)
    b = Eb * tb ** 3 / (12 * (1 - nub ** 2))
    rows = [[a, num * a, 0, 0, 0, 0], [num * a, a, 0, 0, 0, 0], [0, 0, a * (1 - num) / 2, 0, 0, 0],
            [0, 0, 0, b, nub * b, 0], [0, 0, 0, nub * b, b, 0], [0, 0, 0, 0, 0, b * (1 - nub) / 2]]
    return torch.tensor(rows, device=device, dtype=dtype)


def _D_host(membrane, bending, dtype):
    return compute_kirchoff_D_matrix(membrane, bending, device="cpu", dtype=dtype).to(torch.float64)


def compute_shell_nodal_forces(K, shell, displacement, unit, device="cuda:0", dtype=torch.float32):
    """Matrix-free shell operator: rotate nodal vectors into each element frame, apply K_e, rotate back, sum per node in
    ascending element order (shell.py:58-102; the reference scatters with atomics)."""
    dev = _ops.cuda_device(device)
    u = _ops.real(displacement, dev, dtype)
    plan = _ops.cached_plan(shell, u.shape[0], dev)
    return plan.ebe_apply(K, u, 6, unit, dtype)


def compute_global_to_local_displacement(shell, displacement, unit, device="cuda:0"):
    """[M,nen,6]: translations and rotations of every element's nodes rotated into the element frame (shell.py:41-56)."""
    return _ops.shell_local_displacement(shell, displacement, unit, device=device)


_POST_KEYS = ("sx", "sy", "txy", "s1", "s2", "theta_p", "tau_max", "vm_stress")


def compute_shell_postprocess_values(NMQ, t, z=0, device="cuda:0", dtype=torch.float32):
    """Stresses at through-thickness coordinate z from the resultants N, M: sx, sy, txy, principal stresses, principal angle,
    maximum in-plane shear and the plane-stress von Mises value, each [M] (shell.py:104-160; the dict keys are the
    reference's, including 'vm_stress').  One kernel writes all eight rows."""
    out = _ops.shell_postprocess(NMQ, t, z, device=device, dtype=dtype)
    return {k: out[i] for i, k in enumerate(_POST_KEYS)}


# ------------------------------------------------------------------------------------------- S3

def compute_s3_normal(coords, shell, device="cuda:0"):
    """cross(x1-x0, x2-x0)/2 (shell.py:184-203)."""
    return _ops.shell_normal(coords, shell, device)


def identify_s3_shared_edges(shell, device="cuda:0"):
    """[S,2,2] = ((shell,edge),(shell,edge)) for edges used twice (shell.py:205-259)."""
    return _ops.entities(_ops.ENT_TRI_EDGES, shell, device, want_surface=False)[2]


def compute_triangle_surface_faces_with_third_node(shell, device="cuda:0"):
    """Boundary edges (used once) in slot-major order and the off-edge node (shell.py:261-295)."""
    e, x, _ = _ops.entities(_ops.ENT_TRI_EDGES, shell, device, want_shared=False)
    return e, x


def _coords_dtype(coords):
    dt = torch.as_tensor(coords).dtype
    return dt if dt in (torch.float32, torch.float64) else torch.float32


def compute_s3_local_unitvector(coords, shell, device="cuda:0"):
    """[M,3,3] rows e1,e2,e3 in the dtype of `coords` (shell.py:297-321)."""
    return _ops.shell(_ops.S3, 0, coords, shell, device=device, dtype=_coords_dtype(coords))


def compute_s3_global_to_local_coordinates(coords, shell, unit, device="cuda:0", dtype=torch.float32):
    """[M,3,3] node coordinates relative to node 0 in the element frame (shell.py:323-347).  The reference ignores `dtype`
    here and works in the dtype of `coords`; so does this."""
    return _ops.shell_local_coordinates(coords, shell, unit, device=device, dtype=_coords_dtype(coords))


def compute_s3_jacobian(coords, shell, device="cuda:0", dtype=torch.float32):
    """[M,2,2] in the element frame (shell.py:349-384)."""
    return _ops.shell(_ops.S3, 1, coords, shell, device=device, dtype=dtype)


def compute_s3_shape_gradient(coords, shell, device="cuda:0", dtype=torch.float32):
    """[M,3,2]; contracts with Jinv exactly as the reference does (shell.py:386-402)."""
    return _ops.shell(_ops.S3, 2, coords, shell, device=device, dtype=dtype)


def compute_s3_B_matrix(coords, shell, device="cuda:0", dtype=torch.float32):
    """[M,6,18] (shell.py:404-438)."""
    return _ops.shell(_ops.S3, 3, coords, shell, device=device, dtype=dtype)


def compute_s3_K_matrix(coords, shell, membrane, bending, device="cuda:0", dtype=torch.float32):
    """K = B^T D B detJ / 2, [M,18,18] (shell.py:440-453)."""
    return _ops.shell(_ops.S3, 4, coords, shell, D=_D_host(membrane, bending, dtype), device=device, dtype=dtype)


def compute_s3_shell_stress(coords, shell, membrane, bending, displacement, device="cuda:0", dtype=torch.float32):
    """Stress resultants [M,6] = D B u_e with u_e = displacement[shell] ([N,6], taken in global axes as the reference does)
    (shell.py:455-481)."""
    return _ops.shell_ex(_ops.S3, 9, coords, shell, D=_D_host(membrane, bending, dtype), disp=displacement, device=device, dtype=dtype)


# ------------------------------------------------------------------------------------------- S4

def compute_s4_normal(coords, shell, device="cuda:0"):
    """cross(x1-x0, x3-x0) (shell.py:483-502)."""
    return _ops.shell_normal(coords, shell, device)


def identify_s4_shared_edges(shell, device="cuda:0"):
    """shell.py:504-559"""
    return _ops.entities(_ops.ENT_QUAD_EDGES, shell, device, want_surface=False)[2]


def compute_square_surface_faces_with_fourth_node(shell, device="cuda:0"):
    """shell.py:561-597"""
    e, x, _ = _ops.entities(_ops.ENT_QUAD_EDGES, shell, device, want_shared=False)
    return e, x


def compute_s4_local_unitvector(coords, shell, device="cuda:0"):
    """shell.py:597-622"""
    return _ops.shell(_ops.S4, 0, coords, shell, device=device, dtype=_coords_dtype(coords))


def compute_s4_global_to_local_coordinates(coords, shell, unit, device="cuda:0", dtype=torch.float32):
    """[M,4,3] node coordinates relative to node 0 in the element frame, in `dtype` (shell.py:625-649)."""
    return _ops.shell_local_coordinates(coords, shell, unit, device=device, dtype=dtype)


def s4_integration_points(device="cuda:0"):
    """2x2 rule, ALWAYS float32 (shell.py:651-672, quirk q1)."""
    rows = torch.tensor(_ops.default_points(_ops.S4), dtype=torch.float64)
    return rows[:, :2].to(torch.float32).to(device), rows[:, 3].to(torch.float32).to(device)


def _s4_pts(points, weights=None):
    def f64(v):  # python floats must not pass through torch's float32 default
        return v.detach().to("cpu", torch.float64) if torch.is_tensor(v) else torch.tensor(v, dtype=torch.float64)

    p = f64(points).reshape(-1, 2)
    w = torch.ones(p.shape[0], dtype=torch.float64) if weights is None else f64(weights).reshape(-1)
    return [[float(p[q, 0]), float(p[q, 1]), 0.0, float(w[q])] for q in range(p.shape[0])]


def _s4_factor_row(xi, eta, w=1.0):
    """(1-xi, 1+xi, 1-eta, 1+eta, w).  The reference evaluates 1 -/+ xi in the type it was handed: python floats on the K path
    (`.item()`, shell.py:841-842) but 0-dim float32 tensors when compute_s4_B_matrix iterates over the rows of its float32
    rule (shell.py:813-814, :683-695) -- the fp32 rounding of the sums is part of its numbers."""
    def pm(v):
        if torch.is_tensor(v):
            v = v.detach().to("cpu")
            return float(1 - v), float(1 + v)
        return 1 - float(v), 1 + float(v)

    return [*pm(xi), *pm(eta), float(w)]


def _s4_rule_rows(integration_points):
    """Factor rows of a user rule given as (points [n,2], weights [n]) -- or of the default rule iterated as tensors."""
    if integration_points is None:
        pts, wts = s4_integration_points(device="cpu")
    else:
        pts, wts = integration_points
        pts, wts = torch.as_tensor(pts), torch.as_tensor(wts)
    return [_s4_factor_row(pts[q, 0], pts[q, 1], wts[q]) for q in range(pts.shape[0])]


def compute_s4_B_matrix(coords, shell, integration_points=None, single=True, device="cuda:0", dtype=torch.float32):
    """sum_q w_q B(xi_q, eta_q) [M,6,24], or the weighted per-point matrices [M,6,24,4] when single=False (shell.py:802-823).
    `integration_points`, if given, is (points [n,2], weights [n]) as for compute_s4_K_matrix -- the reference only runs with
    the default (its `weights` is unbound otherwise, shell.py:810-816)."""
    return _ops.shell_ex(_ops.S4, 7 if single else 8, coords, shell, factors=_s4_rule_rows(integration_points), device=device, dtype=dtype)


def compute_s4_shell_stress(coords, shell, membrane, bending, displacement, device="cuda:0", dtype=torch.float32):
    """Stress resultants [M,6] = D (sum_q B_q) u_e over the default 2x2 rule, u_e = displacement[shell] ([N,6], global axes)
    (shell.py:863-879)."""
    return _ops.shell_ex(_ops.S4, 9, coords, shell, factors=_s4_rule_rows(None), D=_D_host(membrane, bending, dtype), disp=displacement,
                         device=device, dtype=dtype)


def compute_s4_jacobian(coords, shell, xi, eta, device="cuda:0", dtype=torch.float32):
    """shell.py:674-721"""
    return _ops.shell_ex(_ops.S4, 1, coords, shell, factors=[_s4_factor_row(xi, eta)], device=device, dtype=dtype)


def compute_s4_shape_gradient(coords, shell, xi, eta, device="cuda:0", dtype=torch.float32):
    """shell.py:723-746"""
    return _ops.shell_ex(_ops.S4, 2, coords, shell, factors=[_s4_factor_row(xi, eta)], device=device, dtype=dtype)


def compute_s4_B_matrix_single(coords, shell, xi, eta, device="cuda:0", dtype=torch.float32):
    """shell.py:748-800"""
    return _ops.shell_ex(_ops.S4, 3, coords, shell, factors=[_s4_factor_row(xi, eta)], device=device, dtype=dtype)


def compute_s4_K_matrix(coords, shell, membrane, bending, integration_points=None, single=True, device="cuda:0", dtype=torch.float32):
    """Sum over the 2x2 rule of B^T D B detJ w; single=False -> [M,24,24,4] (shell.py:825-861)."""
    if integration_points is None:
        pts = _ops.default_points(_ops.S4)
    else:
        pts = _s4_pts(integration_points[0], integration_points[1])
    return _ops.shell(_ops.S4, 4 if single else 5, coords, shell, points=pts, D=_D_host(membrane, bending, dtype), device=device, dtype=dtype)


# ------------------------------------------------------------------------------------------- shell -> wedge / hexahedron

def shell_extrude(coords, tri, quad, thickness, device="cuda:0", dtype=torch.float32):
    """Extrudes a mid-surface mesh along its node normals: (coords_3d [2N,3] = bottom layer then top layer, wedges [T,6] =
    (tri | tri+N), hexahedra [S,8] = (quad | quad+N)) (shell.py:885-983).  Node normals are the normalised mean of the unit
    normals of the incident triangles and of both halves (0,1,2), (0,2,3) of the incident quads; every node sums them over
    its incidence list in the reference's order instead of scattering with atomics."""
    return _ops.shell_extrude(coords, tri, quad, thickness, device=device, dtype=dtype)
