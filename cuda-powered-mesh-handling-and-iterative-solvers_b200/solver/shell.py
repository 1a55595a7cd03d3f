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


# ------------------------------------------------------------------------------------------- S3

def compute_s3_normal(coords, shell, device="cuda:0"):
    """cross(x1-x0, x2-x0)/2 (shell.py:184-203) = area * third row of the local frame."""
    dev = _ops.cuda_device(device)
    x = torch.as_tensor(coords).to(dev)
    dt = x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32
    unit = _ops.shell(_ops.S3, 0, x, shell, device=dev, dtype=dt)
    J = _ops.shell(_ops.S3, 1, x, shell, device=dev, dtype=dt)
    return unit[:, 2, :] * (0.5 * (J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0])).unsqueeze(1)


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


# ------------------------------------------------------------------------------------------- S4

def compute_s4_normal(coords, shell, device="cuda:0"):
    """cross(x1-x0, x3-x0) (shell.py:483-502)."""
    dev = _ops.cuda_device(device)
    x = torch.as_tensor(coords).to(dev)
    s = _ops.index(shell, dev).long()
    a, b = x[s[:, 1]] - x[s[:, 0]], x[s[:, 3]] - x[s[:, 0]]
    return torch.linalg.cross(a, b, dim=1)


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


def compute_s4_jacobian(coords, shell, xi, eta, device="cuda:0", dtype=torch.float32):
    """shell.py:674-721"""
    return _ops.shell(_ops.S4, 1, coords, shell, points=_s4_pts([[float(xi), float(eta)]]), device=device, dtype=dtype)


def compute_s4_shape_gradient(coords, shell, xi, eta, device="cuda:0", dtype=torch.float32):
    """shell.py:723-746"""
    return _ops.shell(_ops.S4, 2, coords, shell, points=_s4_pts([[float(xi), float(eta)]]), device=device, dtype=dtype)


def compute_s4_B_matrix_single(coords, shell, xi, eta, device="cuda:0", dtype=torch.float32):
    """shell.py:748-800"""
    return _ops.shell(_ops.S4, 3, coords, shell, points=_s4_pts([[float(xi), float(eta)]]), device=device, dtype=dtype)


def compute_s4_K_matrix(coords, shell, membrane, bending, integration_points=None, single=True, device="cuda:0", dtype=torch.float32):
    """Sum over the 2x2 rule of B^T D B detJ w; single=False -> [M,24,24,4] (shell.py:825-861)."""
    if integration_points is None:
        pts = _ops.default_points(_ops.S4)
    else:
        pts = _s4_pts(integration_points[0], integration_points[1])
    return _ops.shell(_ops.S4, 4 if single else 5, coords, shell, points=pts, D=_D_host(membrane, bending, dtype), device=device, dtype=dtype)
