"""Row/element partition of a mesh over P ranks and the halo-exchange plan (host-side plumbing, torch ops, any device).

The reference has no distributed code; its only partition is the notebook's element region-growing on one GPU
(reference subdivision.ipynb cell 9, random first seed -> not reproducible).  Here nodes (= operator rows) are split by a
deterministic recursive coordinate bisection (a geometric graph partition); every rank also takes each element that
touches one of its nodes, so assembly needs no communication (ghost elements are recomputed redundantly).

Local numbering on a rank: [owned nodes, interior rows first | padding to a 128-byte line | ghost nodes grouped by owner
rank, ascending id].
Both sides derive the send/recv lists from the same rule (nodes of rank r that share an element with a node of rank q,
ascending global id), so the plan needs no communication either.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import torch


def rcb_labels(coords: torch.Tensor, nparts: int) -> torch.Tensor:
    """Recursive coordinate bisection: label in [0,nparts) per node; nparts need not be a power of two.
    Splits along the longest extent at the weighted median; ties are broken by node id (stable sort) -> deterministic."""
    N = coords.shape[0]
    labels = torch.zeros(N, dtype=torch.int64, device=coords.device)
    work = [(torch.arange(N, device=coords.device), 0, nparts)]
    while work:
        ids, base, k = work.pop()
        if k == 1:
            labels[ids] = base
            continue
        kl = k // 2
        c = coords[ids]
        ext = c.max(dim=0).values - c.min(dim=0).values if ids.numel() else torch.zeros(3)
        ax = int(torch.argmax(ext).item()) if ids.numel() else 0
        order = torch.sort(c[:, ax], stable=True).indices
        cut = (ids.numel() * kl) // k
        work.append((ids[order[:cut]], base, kl))
        work.append((ids[order[cut:]], base + kl, k - kl))
    return labels


def block_labels(n_nodes: int, nparts: int, device="cpu") -> torch.Tensor:
    """Coordinate-free partition: contiguous node-id ranges of (almost) equal size.  What the solver API falls back to when the
    caller gives no coordinates (the reference's solvers take none); its quality is that of the mesh numbering."""
    return (torch.arange(n_nodes, device=device) * nparts // max(n_nodes, 1)).clamp_(max=nparts - 1)


@dataclass
class LocalPart:
    rank: int
    nparts: int
    n_owned: int
    n_ghost: int
    owned_global: torch.Tensor            # [n_owned] global ids, ascending
    ghost_global: torch.Tensor            # [n_ghost] global ids, grouped by owner
    elements_local: torch.Tensor          # [Ml, nen] in local numbering
    element_ids: torch.Tensor             # [Ml] global element ids
    neighbors: List[int] = field(default_factory=list)
    send_idx: Dict[int, torch.Tensor] = field(default_factory=dict)   # q -> local owned indices to push to q
    recv_off: Dict[int, int] = field(default_factory=dict)            # q -> offset of q's block inside the ghost region
    recv_cnt: Dict[int, int] = field(default_factory=dict)
    n_interior: int = 0                   # owned rows [0, n_interior) reference no ghost column

    @property
    def ghost_base(self):
        """Local index of the first ghost: n_owned rounded up to 16 doubles (one 128-byte line), so that no cache line holds
        both owned entries (read early by interior rows) and ghost entries (written by peers later in the same kernel)."""
        return (self.n_owned + 15) // 16 * 16

    @property
    def n_local(self):
        return self.ghost_base + self.n_ghost


def build_local_part(elements: torch.Tensor, labels: torch.Tensor, rank: int, nparts: int) -> LocalPart:
    """Everything rank `rank` needs, computed from the replicated (elements, labels)."""
    dev = elements.device
    lab = labels[elements]                                            # [M, nen]
    touch = (lab == rank).any(dim=1)
    eids = torch.nonzero(touch).reshape(-1)
    el = elements[eids]
    ll = lab[eids]
    owned = torch.nonzero(labels == rank).reshape(-1)                 # ascending
    # interior rows first: an owned node is "boundary" when one of its elements holds a node of another rank, i.e. its
    # operator row references ghost columns; interior rows can be multiplied while the halo is still in flight
    mixed = (ll != rank).any(dim=1)
    bflag = torch.zeros(labels.numel(), dtype=torch.bool, device=dev)
    bflag[el[mixed].reshape(-1)] = True
    is_b = bflag[owned]
    owned = torch.cat([owned[~is_b], owned[is_b]])
    n_interior = int((~is_b).sum().item())
    nodes = torch.unique(el)
    gh = nodes[labels[nodes] != rank]
    gh_owner = labels[gh]
    order = torch.sort(gh_owner * (labels.numel() + 1) + gh).indices  # by owner, then id
    gh, gh_owner = gh[order], gh_owner[order]
    g2l = torch.full((labels.numel(),), -1, dtype=torch.int64, device=dev)
    g2l[owned] = torch.arange(owned.numel(), device=dev)
    g2l[gh] = (owned.numel() + 15) // 16 * 16 + torch.arange(gh.numel(), device=dev)
    part = LocalPart(rank, nparts, int(owned.numel()), int(gh.numel()), owned, gh, g2l[el], eids)
    part.n_interior = n_interior
    for q in sorted(set(gh_owner.tolist())):
        sel = gh_owner == q
        part.neighbors.append(q)
        part.recv_off[q] = int(torch.nonzero(sel)[0].item())
        part.recv_cnt[q] = int(sel.sum().item())
        # what q needs from me: my nodes in elements that also hold a node of q (ascending global id)
        has_q = (ll == q).any(dim=1)
        mine = torch.unique(el[has_q][ll[has_q] == rank])
        part.send_idx[q] = g2l[mine]
    return part


def localize(vec_global: torch.Tensor, part: LocalPart) -> torch.Tensor:
    """Rows of a replicated global [N, ...] array in local numbering (owned, zero padding up to ghost_base, ghost)."""
    pad = torch.zeros((part.ghost_base - part.n_owned,) + tuple(vec_global.shape[1:]), dtype=vec_global.dtype, device=vec_global.device)
    return torch.cat([vec_global[part.owned_global], pad, vec_global[part.ghost_global]], dim=0)


def halo_tables(part: LocalPart, sizes, block: int = 1):
    """The halo plan of one rank as flat tables for the device loop (pure torch, no device needed; femb_dist_cg_solve's arguments).
    `sizes[q]` = {"ghost_base", "recv_off"} of every rank q.  `block` dofs per node: rows, offsets and counts become dof-level
    (`block*node + c`, component fastest -- the ghost blocks of the receiver use the same interleaving).  Returns a dict:
      nbr        neighbour ranks (ascending)
      send_ptr   [nnbr+1] prefix into send_idx per neighbour
      send_idx   owned dof rows to push to each neighbour, in the receiver's ghost order
      ghost_off  per neighbour: dof index in ITS p vector where my block starts
      bptr/bk/boff  CSR over the boundary dof rows [n_interior*block, n_owned*block): neighbour index and offset inside that
                 neighbour's block for every destination of the row (None when the rows are not ordered interior-first)
    """
    B = int(block)
    nb = list(part.neighbors)

    def dofs(t):
        t = torch.as_tensor(t)
        return t if B == 1 else (t.reshape(-1, 1) * B + torch.arange(B, device=t.device)).reshape(-1)

    ptrs, idx = [0], []
    for q in nb:
        idx.append(dofs(part.send_idx[q]).to(torch.int32))
        ptrs.append(ptrs[-1] + B * int(part.send_idx[q].numel()))
    out = {"nbr": nb, "send_ptr": ptrs, "send_idx": torch.cat(idx) if idx else torch.zeros(1, dtype=torch.int32),
           "ghost_off": [B * (sizes[q]["ghost_base"] + sizes[q]["recv_off"][part.rank]) for q in nb],
           "bptr": None, "bk": None, "boff": None}
    ni, no = int(getattr(part, "n_interior", 0)), part.n_owned
    if nb and ni > 0:
        rows = torch.cat([part.send_idx[q] for q in nb])
        ks = torch.cat([torch.full((part.send_idx[q].numel(),), k, dtype=torch.int64) for k, q in enumerate(nb)]).to(rows.device)
        offs = torch.cat([torch.arange(part.send_idx[q].numel()) for q in nb]).to(rows.device)
        assert int(rows.min().item()) >= ni, "send rows must be boundary rows"
        if B > 1:       # every boundary node row becomes B dof rows with the same destinations, offsets scaled
            c = torch.arange(B, device=rows.device)
            nrow = rows.numel()
            rows = (rows.reshape(-1, 1) * B + c).reshape(-1)
            ks = ks.reshape(-1, 1).expand(nrow, B).reshape(-1)
            offs = (offs.reshape(-1, 1) * B + c).reshape(-1)
        order = torch.sort(rows, stable=True).indices
        rows, ks, offs = rows[order], ks[order], offs[order]
        ni, no = ni * B, no * B
        cnt = torch.bincount(rows - ni, minlength=no - ni)
        bptr = torch.zeros(no - ni + 1, dtype=torch.int64, device=rows.device)
        bptr[1:] = torch.cumsum(cnt, 0)
        out["bptr"], out["bk"], out["boff"] = bptr.to(torch.int32), ks.to(torch.uint8), offs.to(torch.int32)
    return out

