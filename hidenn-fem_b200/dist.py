"""Multi-GPU: element-block partition with halo nodes + one packed all-reduce per evaluation.

New capability (the reference is single-process); design of SURVEY.md §8(e):
each rank owns a contiguous block of elements and holds every node those elements touch.  Nodes
touched by more than one rank ("shared") receive partial gradients on each rank; ONE
`all_reduce(sum)` over the packed buffer  [loss, (gx,gy,gu,gv) of every shared node]  completes
them, so the collective moves O(sqrt(Ne)) values, not the full gradient.  The packing order is the
ascending global node id, identical on every rank.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import meshgen
from .loss import EnergyLoss2D


# ------------------------------------------------------------------------------------------------
# host-side index plan (pure integer work; unit-tested on CPU with gloo)
# ------------------------------------------------------------------------------------------------
@dataclass
class HaloPlan:
    shared_gid: np.ndarray      # [S] ascending global ids of all shared nodes (same on every rank)
    local_pos: np.ndarray       # [s] positions in shared_gid of the shared nodes this rank holds
    local_node: np.ndarray      # [s] their local node index
    x_rows: np.ndarray          # [sx] rows of node_coords_free (only free ones)
    x_pos: np.ndarray           # [sx] matching positions in shared_gid
    u_rows: np.ndarray          # [su] rows of u_free
    u_pos: np.ndarray           # [su]

    @property
    def buffer_len(self):
        return 2 + 4 * self.shared_gid.shape[0]


def shared_ids_from_candidates(cands: Sequence[np.ndarray]) -> np.ndarray:
    """Global ids that appear in the candidate lists of at least two ranks."""
    allc = np.concatenate([np.unique(c) for c in cands]) if len(cands) else np.zeros(0, np.int64)
    ids, cnt = np.unique(allc, return_counts=True)
    return ids[cnt >= 2]


def shared_owner_from_candidates(cands: Sequence[np.ndarray], shared_gid: np.ndarray) -> np.ndarray:
    """Owner rank of every shared node = the lowest rank whose candidate list holds it ([S] int64)."""
    owner = np.full(shared_gid.shape, len(cands), np.int64)
    for r in range(len(cands) - 1, -1, -1):
        owner[np.isin(shared_gid, cands[r])] = r
    return owner


def owner_weights(plan: "HaloPlan", shared_owner: np.ndarray, rank: int, n_free_x: int, n_free_u: int):
    """Per-row weights for global reductions over sharded Parameters (optim.ShardedLBFGS): 1 for the rows this rank
    owns, 0 for its copies of shared rows owned by a lower rank.  Returns (wx [n_free_x], wu [n_free_u]) float64."""
    wx, wu = np.ones(n_free_x), np.ones(n_free_u)
    wx[plan.x_rows[shared_owner[plan.x_pos] != rank]] = 0.0
    wu[plan.u_rows[shared_owner[plan.u_pos] != rank]] = 0.0
    return wx, wu


def build_halo_plan(global_node_id: np.ndarray, free_mask: np.ndarray, u_free_mask: np.ndarray,
                    shared_gid: np.ndarray) -> HaloPlan:
    order = np.argsort(global_node_id, kind="stable")
    sorted_gid = global_node_id[order]
    pos_in_local = np.searchsorted(sorted_gid, shared_gid)
    pos_in_local = np.clip(pos_in_local, 0, max(sorted_gid.size - 1, 0))
    have = sorted_gid[pos_in_local] == shared_gid if sorted_gid.size else np.zeros(shared_gid.shape, bool)
    local_pos = np.nonzero(have)[0]
    local_node = order[pos_in_local[have]]
    xrow_of_node = np.cumsum(free_mask) - 1
    urow_of_node = np.cumsum(u_free_mask) - 1
    fx = free_mask[local_node]
    fu = u_free_mask[local_node]
    return HaloPlan(shared_gid, local_pos, local_node,
                    xrow_of_node[local_node][fx].astype(np.int32), local_pos[fx].astype(np.int64),
                    urow_of_node[local_node][fu].astype(np.int32), local_pos[fu].astype(np.int64))


def gather_candidates(local_cand: np.ndarray, group=None) -> list:
    """all_gather of variable-length int64 id lists (setup time only)."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n = torch.tensor([local_cand.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.full((max(mx, 1),), -1, dtype=torch.int64, device=dev)
    pad[:local_cand.size] = torch.from_numpy(local_cand.astype(np.int64)).to(dev)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:int(s.item())].cpu().numpy() for b, s in zip(bufs, sizes)]


# ------------------------------------------------------------------------------------------------
# partitioners
# ------------------------------------------------------------------------------------------------
def strip_mesh(nx: int, ny: int, rank: int, world: int, **kw) -> meshgen.PlateMesh:
    """Rank's contiguous block of cell columns of the global nx x ny plate (bit-identical to the
    corresponding part of the global mesh; ordering applied inside the strip)."""
    ncol = nx - 1
    c0 = rank * ncol // world
    c1 = (rank + 1) * ncol // world
    return meshgen.plate_mesh(nx, ny, col_range=(c0, c1), **kw)


def strip_candidates(mesh: meshgen.PlateMesh) -> np.ndarray:
    """Nodes on the strip's first / last node column (only they can be shared with a neighbour)."""
    ny = mesh.meta["ny"]
    c0, c1 = mesh.meta["col_range"]
    gix = mesh.global_node_id // ny
    return mesh.global_node_id[(gix == c0) | (gix == c1)]


def partition_elements(mesh: meshgen.PlateMesh, world: int, rank: int) -> meshgen.PlateMesh:
    """General partition of an in-memory mesh: contiguous blocks of the (locality-ordered) element list;
    the rank keeps the nodes its elements touch, renumbered ascending.  Neumann edges follow the rank
    that owns both end nodes' element (an edge on the boundary belongs to exactly one element)."""
    Ne = mesh.connectivity.shape[0]
    e0, e1 = rank * Ne // world, (rank + 1) * Ne // world
    conn = mesh.connectivity[e0:e1]
    used = np.unique(conn)
    new = -np.ones(mesh.node_coords.shape[0], np.int64)
    new[used] = np.arange(used.size)
    lconn = new[conn]
    # edges of my elements
    e_all = np.concatenate([conn[:, [0, 1]], conn[:, [1, 2]], conn[:, [2, 0]]], 0)
    e_all = np.sort(e_all, 1)
    key = e_all[:, 0] * mesh.node_coords.shape[0] + e_all[:, 1]
    # match on the sorted end nodes, keep the stored orientation (the reference's raw [-1,1] edge rule makes the edge
    # term orientation-dependent); renumbered / gmsh-style edges need not be stored ascending
    ns = np.sort(mesh.neumann_edges, axis=1)
    nk = ns[:, 0] * mesh.node_coords.shape[0] + ns[:, 1]
    mine = np.isin(nk, key)
    ledges = new[mesh.neumann_edges[mine]]
    return meshgen.PlateMesh(mesh.node_coords[used], lconn, mesh.boundary_mask[used], mesh.dirichlet_mask[used],
                             mesh.neumann_mask[used], ledges, mesh.global_node_id[used],
                             meta=dict(mesh.meta, part=(rank, world)))


# ------------------------------------------------------------------------------------------------
# device-side exchange
# ------------------------------------------------------------------------------------------------
class HaloExchange:
    """[loss, shared-node gradients]: one pack kernel, one NCCL all-reduce, one unpack kernel per evaluation.

    Message layout  [loss, 0 | gx pairs (S) | gu pairs (S)]  in ascending global node id (same on every rank).
    The message buffer holds zeros at the positions of shared nodes this rank does not hold (or holds as fixed),
    so the in-place sum over ranks is exactly the sum of the partial gradients."""

    def __init__(self, plan: HaloPlan, device, dtype, group=None):
        self.plan = plan
        self.group = group
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.HidennError("HaloExchange runs on CUDA tensors only (no CPU fallback)")
        self.dtype = dtype
        S = plan.shared_gid.shape[0]
        self.S = S
        # two alternating message buffers: the pack kernel of one step clears the buffer of the next step
        self.bufs = [torch.zeros(2 + 4 * S, device=self.device, dtype=dtype) for _ in range(2)]
        self.parity = 0
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.device)
        self.x_rows, self.x_pos = t(plan.x_rows), t(plan.x_pos)
        self.u_rows, self.u_pos = t(plan.u_rows), t(plan.u_pos)

    def exchange(self, out, gx, gu):
        """out: device [4] (loss first); gx/gu: parameter-layout gradients or None. In place."""
        s = _lib.stream_ptr()
        dt = self.dtype
        nx, nu = C.c_int64(self.x_rows.numel()), C.c_int64(self.u_rows.numel())
        buf, other = self.bufs[self.parity], self.bufs[self.parity ^ 1]
        self.parity ^= 1
        _lib.check(_lib.fn("hidenn_halo_pack_all", dt)(
            _lib.ptr(gx), _lib.ptr(self.x_rows), _lib.ptr(self.x_pos), nx, _lib.ptr(gu), _lib.ptr(self.u_rows), _lib.ptr(self.u_pos), nu,
            _lib.ptr(out), C.c_int64(self.S), _lib.ptr(buf), _lib.ptr(other), s))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        _lib.check(_lib.fn("hidenn_halo_unpack_all", dt)(
            _lib.ptr(gx), _lib.ptr(self.x_rows), _lib.ptr(self.x_pos), nx, _lib.ptr(gu), _lib.ptr(self.u_rows), _lib.ptr(self.u_pos), nu,
            _lib.ptr(out), C.c_int64(self.S), _lib.ptr(buf), s))


class DistributedEnergyLoss2D(EnergyLoss2D):
    """EnergyLoss2D over a partitioned mesh: local fused kernels, then one packed all-reduce.
    Every rank returns the global loss; shared-node gradients are complete (and bit-identical) on
    all ranks that hold the node, so element-wise optimisers keep halo copies consistent."""

    def __init__(self, *args, halo: Optional[HaloExchange] = None, **kw):
        super().__init__(*args, **kw)
        self.halo = halo

    def _post_forward(self, model, out, gx, gu):
        if self.halo is not None:
            self.halo.exchange(out, gx, gu)


def setup_strip_halo(mesh: meshgen.PlateMesh, boundary_mask: np.ndarray, dirichlet_mask: np.ndarray, device, dtype,
                     group=None) -> HaloExchange:
    cands = gather_candidates(strip_candidates(mesh), group)
    shared = shared_ids_from_candidates(cands)
    plan = build_halo_plan(mesh.global_node_id, ~boundary_mask, ~dirichlet_mask, shared)
    halo = HaloExchange(plan, device, dtype, group)
    rank = dist.get_rank(group)
    wx, wu = owner_weights(plan, shared_owner_from_candidates(cands, shared), rank, int((~boundary_mask).sum()),
                           int((~dirichlet_mask).sum()))
    # row weights for optim.ShardedLBFGS(model.parameters(), weights=halo.row_weights): [node_coords_free, u_free] order
    halo.row_weights = [torch.from_numpy(wx).to(halo.device, dtype), torch.from_numpy(wu).to(halo.device, dtype)]
    return halo
