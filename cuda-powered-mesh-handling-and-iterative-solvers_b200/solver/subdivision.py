"""Drop-in for the helper functions the reference defines inside `subdivision.ipynb` (cells 6-9, 13): memory-driven
subdivision count, element face-adjacency graph, farthest-seed region-growing partition and per-part local operators, and of
its closing fragments (cells 12, 14, 15): dense subdomain inverses, interface-force unknowns and per-subdomain load vectors.

Same names and argument order as the notebook cells.  The notebook runs every BFS level as a `torch.sparse.mm` with a dense
[n_parts, M] frontier and renumbers nodes through a Python dict; here the adjacency is a CSR built by one radix sort
(`femb_graph_from_pairs`), a BFS level is one kernel over the unlabelled elements (`femb_graph_bfs`), node renumbering is a
sort-based unique, and the local operators are deterministic coalesced CSR (the notebook keeps an un-coalesced COO).
The notebook's first seed is `torch.randint` without a seed; pass `first_seed` (default 0) to make the partition
reproducible.  Elements reached by several regions in the same BFS level go to the highest part index (the notebook's CPU
behaviour; on CUDA its duplicate-index assignment is unspecified).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys
from collections import defaultdict

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from femb200 import ops as _ops  # noqa: E402
from femb200._lib import check, lib  # noqa: E402


def compute_subdivisions(matrix_size, gpu_memory_gb):
    """subdivision.ipynb cell 7: how many float32 dense blocks of at most `gpu_memory_gb` a matrix_size^2 operator needs."""
    max_dim = int(math.floor(math.sqrt((gpu_memory_gb * (1024 ** 3)) // 4)))
    return math.ceil(matrix_size / max_dim)


class ElementGraph:
    """CSR face-adjacency of the elements: crow int32 [M+1], col int32 [2S] (neighbours ascending)."""

    def __init__(self, crow, col):
        self.crow, self.col, self.M, self.dev = crow, col, crow.numel() - 1, crow.device

    def bfs(self, sources, want_labels=False):
        """(dist int32 [M] (-1 = unreachable), labels int32 [M] or None, levels)."""
        src = torch.as_tensor(sources, device=self.dev).to(torch.int64).reshape(-1).contiguous()
        dist = torch.empty(self.M, device=self.dev, dtype=torch.int32)
        lab = torch.empty(self.M, device=self.dev, dtype=torch.int32) if want_labels else None
        levels = C.c_int32()
        with torch.cuda.device(self.dev):
            check(lib.femb_graph_bfs(_ops._p(self.crow), _ops._p(self.col), self.M, _ops._p(src), src.numel(), _ops._p(dist), _ops._p(lab),
                                     C.byref(levels), _ops._stream(self.dev)), "femb_graph_bfs")
        return dist, lab, levels.value


def build_adjacency_matrix(edge, num_elements, device):
    """cell 9: `edge` is [2,S] (the two element ids of every shared face, as cell 8 builds it from
    identify_tetrahedral_shared_faces) or the [S,2,2] shared-face tensor itself.  Returns an ElementGraph."""
    dev = _ops.cuda_device(device)
    e = torch.as_tensor(edge).to(dev)
    if e.dim() == 3:                       # [S,2,2]: (element, local face) x 2
        pairs, stride = e.to(torch.int64).contiguous(), 4
    else:
        pairs, stride = e.to(torch.int64).t().contiguous(), 2
    S = pairs.shape[0]
    crow = torch.empty(num_elements + 1, device=dev, dtype=torch.int32)
    col = torch.empty(2 * S, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        check(lib.femb_graph_from_pairs(_ops._p(pairs), S, stride, num_elements, _ops._p(crow), _ops._p(col), _ops._stream(dev)),
              "femb_graph_from_pairs")
    return ElementGraph(crow, col)


def pick_distant_seeds(adj, n_parts, first_seed=0):
    """cell 9: farthest-point seeds by repeated multi-source BFS (an unreachable element counts as infinitely far)."""
    seeds = [int(first_seed)]
    for _ in range(n_parts - 1):
        dist, _, _ = adj.bfs(seeds)
        far = torch.where(dist < 0, torch.iinfo(torch.int32).max, dist)
        seeds.append(int(torch.argmax(far).item()))
    return torch.tensor(seeds, device=adj.dev)


def region_growing_partition(edge, n_parts, num_elements, device="cuda:0", first_seed=0):
    """cell 9 -> (groups: list of element-id tensors per part, seeds).  Elements of components that contain no seed are
    left out of every group (the notebook loops forever on a disconnected mesh)."""
    adj = edge if isinstance(edge, ElementGraph) else build_adjacency_matrix(edge, num_elements, device)
    seeds = pick_distant_seeds(adj, n_parts, first_seed)
    _, labels, _ = adj.bfs(seeds, want_labels=True)
    order = torch.argsort(labels.long(), stable=True)          # ascending element id inside a part
    counts = torch.bincount(labels[labels >= 0].long(), minlength=n_parts)
    skip = int((labels < 0).sum().item())
    groups = list(torch.split(order[skip:], counts.tolist()))
    return groups, seeds


def build_sparse_K_local(K, elements, element_indices, device="cuda:0"):
    """cell 9 -> (K_local, global_nodes): the part's operator in its own node numbering (ascending global id) as a
    coalesced torch CSR tensor [3 n_local, 3 n_local], and the global ids of its nodes."""
    dev = _ops.cuda_device(device)
    idx = torch.as_tensor(element_indices).to(dev).long()
    Kp = torch.as_tensor(K).to(dev)[idx]
    elems = torch.as_tensor(elements).to(dev).long()[idx]
    global_nodes, inv = torch.unique(elems, return_inverse=True)
    local = inv.reshape(elems.shape).contiguous()
    ndof = Kp.shape[1] // elems.shape[1]
    plan = _ops.CsrPlan(local, int(global_nodes.numel()), dev)
    crow, col = plan.pattern(ndof)
    vals = plan.assemble(Kp, ndof)
    n = plan.n_nodes * ndof
    return torch.sparse_csr_tensor(crow, col, vals, size=(n, n), device=dev), global_nodes


def partition_and_build_sparse_K(K, elements, edge, n_parts, device="cuda:0", first_seed=0):
    """cell 9 -> (K_parts, node_maps, groups, seeds)."""
    groups, seeds = region_growing_partition(edge, n_parts, torch.as_tensor(K).shape[0], device=device, first_seed=first_seed)
    K_parts, node_maps = [], []
    for g in groups:
        K_local, global_nodes = build_sparse_K_local(K, elements, g, device=device)
        K_parts.append(K_local)
        node_maps.append(global_nodes)
    return K_parts, node_maps, groups, seeds


def build_ordered_subdomain_map(node_maps):
    """cell 13: {sorted tuple of parts sharing a node: [node ids]} for interface nodes.  Computed from one sort of the
    (node, part) pairs instead of a Python set per node; the dict itself is host data as in the notebook."""
    nodes = torch.cat([nm.reshape(-1) for nm in node_maps]).cpu()
    parts = torch.cat([torch.full((nm.numel(),), i, dtype=torch.long) for i, nm in enumerate(node_maps)])
    order = torch.argsort(nodes * len(node_maps) + parts)
    nodes, parts = nodes[order].tolist(), parts[order].tolist()
    out = defaultdict(list)
    i = 0
    while i < len(nodes):
        j = i
        while j < len(nodes) and nodes[j] == nodes[i]:
            j += 1
        if j - i > 1:
            out[tuple(parts[i:j])].append(nodes[i])
        i = j
    return dict(out)


# ------------------------------------------------------------------------------------------- cells 12, 14, 15 (fragments)

def invert_K_parts(K_parts):
    """cell 12: dense inverse of every subdomain operator (a library call in the notebook too: torch.linalg.inv of
    K.to_dense()).  Floating subdomains are singular -- the notebook inverts them regardless, and so does this."""
    return [torch.linalg.inv(K.to_dense()) for K in K_parts]


def count_free_variables(group_to_nodes):
    """Number of interface-force unknowns: (parts sharing the node - 1) per interface node (cell 14)."""
    return sum((len(k) - 1) * len(v) for k, v in group_to_nodes.items())


def make_free_variables(group_to_nodes, device="cuda:0", dtype=torch.float64):
    """cell 14: zero-initialised interface-force unknowns [n_free, 3] (the notebook builds them from a Python list of
    [0,0,0] rows, which makes an int64 tensor; pass the dtype of the load vector instead)."""
    return torch.zeros((count_free_variables(group_to_nodes), 3), device=device, dtype=dtype)


class SubdomainForcePlan:
    """Index form of cell 15's double loop over (group, node): for every (subdomain, interface node) the unknown it gains and
    the one it loses.  Built once per partition on the host (the notebook walks the dict on every call)."""

    def __init__(self, group_to_nodes, rbe2, N, n_sub, device):
        fixed = set(int(v) for v in (rbe2.reshape(-1).tolist() if torch.is_tensor(rbe2) else (rbe2 if rbe2 is not None else ())))
        tgt, plus, minus = [], [], []
        off = 0
        for k, v in group_to_nodes.items():
            m = len(k) - 1
            sd = sorted(k)
            idx = off
            for nd in v:
                if nd not in fixed:
                    for j in range(m + 1):
                        tgt.append(sd[j] * N + nd)
                        plus.append(idx + j if j < m else -1)
                        minus.append(idx + j - 1 if j > 0 else -1)
                idx += m
            off += m * len(v)
        if len(set(tgt)) != len(tgt):
            raise ValueError("a node appears in more than one interface group")
        if tgt and (max(tgt) >= n_sub * N or min(tgt) < 0):
            raise IndexError("interface group refers to a subdomain / node outside (n_sub, N)")
        self.n_free, self.N, self.n_sub, self.dev = off, N, n_sub, device
        self.tgt = torch.tensor(tgt, device=device, dtype=torch.int64)
        self.plus = torch.tensor(plus, device=device, dtype=torch.int32)
        self.minus = torch.tensor(minus, device=device, dtype=torch.int32)

    def apply(self, free_vars, F):
        F = _ops.real(F, self.dev, F.dtype if F.dtype in (torch.float32, torch.float64) else torch.float64)
        fv = _ops.real(free_vars, self.dev, F.dtype).reshape(-1, 3)
        if fv.shape[0] < self.n_free or tuple(F.shape) != (self.N, 3):
            raise ValueError(f"expected free_vars [>={self.n_free},3] and F [{self.N},3], got {tuple(fv.shape)} and {tuple(F.shape)}")
        out = torch.empty((self.n_sub, self.N, 3), device=self.dev, dtype=F.dtype)
        with torch.cuda.device(self.dev):
            check(lib.femb_subdomain_forces(_ops._p(F), F.element_size(), self.N, self.n_sub, _ops._p(fv), self.tgt.numel(), _ops._p(self.tgt),
                                            _ops._p(self.plus), _ops._p(self.minus), _ops._p(out), _ops._stream(self.dev)),
                  "femb_subdomain_forces")
        return list(out.unbind(0))


def make_sub_domain_forces(free_vars, group_to_nodes, rbe2, F, n_sub, device="cuda", plan=None):
    """cell 15: the load vector of every subdomain = F plus the interface forces it exchanges with the other subdomains that
    share its interface nodes (chained +f_j / -f_j so that they cancel in the sum over subdomains); nodes in `rbe2` carry
    none.  Returns n_sub tensors [N,3] in the global node numbering, as the notebook does.  Pass `plan`
    (a SubdomainForcePlan) to reuse the index arrays across calls of an interface iteration."""
    dev = _ops.cuda_device("cuda:0" if str(device) == "cuda" else device)
    F = torch.as_tensor(F).to(dev)
    if plan is None:
        plan = SubdomainForcePlan(group_to_nodes, rbe2, F.shape[0], n_sub, dev)
    return plan.apply(free_vars, F)
