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


@dataclass
class PeerTables:
    """Index tables of the peer-memory exchange (HaloP2P), all int32 numpy arrays; built by `build_peer_tables`.
    Two ranks that share nodes list them in ascending global id: the k-th common node uses slot k in both directions."""
    smax: int                   # largest number of nodes any two ranks share (same on every rank)
    send_xrow: np.ndarray       # [n_send] row of node_coords_free (-1: coordinate fixed) ...
    send_urow: np.ndarray       # ... and of u_free of the node to put
    send_peer: np.ndarray       # destination rank
    send_k: np.ndarray          # slot in the destination's block of this sender
    node_xrow: np.ndarray       # [n_nodes] my shared nodes (ascending global id): gradient rows to complete
    node_urow: np.ndarray
    node_off: np.ndarray        # [n_nodes+1] sources of node j: [node_off[j], node_off[j+1])
    src_rank: np.ndarray        # holder ranks of the node, ascending, INCLUDING this rank
    src_k: np.ndarray           # slot of the node in the (this rank, src_rank) list; 0 for this rank itself
    wait_ranks: np.ndarray      # ranks this rank receives from
    local_node: np.ndarray      # [n_nodes] local node index of my shared nodes


def build_peer_tables(cands: Sequence[np.ndarray], shared_gid: np.ndarray, rank: int, global_node_id: np.ndarray,
                      free_mask: np.ndarray, u_free_mask: np.ndarray) -> PeerTables:
    world = len(cands)
    S = shared_gid.shape[0]
    member = np.stack([np.isin(shared_gid, c) for c in cands]) if S else np.zeros((world, 0), bool)      # [world, S]
    smax = 0
    for a in range(world):
        for b in range(a + 1, world):
            smax = max(smax, int((member[a] & member[b]).sum()))
    mine = member[rank]
    # local rows of my shared nodes
    order = np.argsort(global_node_id, kind="stable")
    sorted_gid = global_node_id[order]
    my_gid = shared_gid[mine]
    pos = np.searchsorted(sorted_gid, my_gid)
    assert my_gid.size == 0 or np.array_equal(sorted_gid[pos], my_gid), "a shared node of this rank is not in its local mesh"
    local_node = order[pos]
    xrow_of = np.where(free_mask, np.cumsum(free_mask) - 1, -1)
    urow_of = np.where(u_free_mask, np.cumsum(u_free_mask) - 1, -1)
    nx, nu = xrow_of[local_node], urow_of[local_node]
    my_index = np.cumsum(mine) - 1                      # index of shared node s in my list
    sx, su, sp, sk = [], [], [], []
    k_with = {}
    for q in range(world):
        if q == rank:
            continue
        common = mine & member[q]
        if not common.any():
            continue
        k = np.cumsum(common) - 1                       # slot of shared node s in the (rank, q) list
        k_with[q] = k
        idx = my_index[common]
        sx.append(nx[idx]); su.append(nu[idx]); sp.append(np.full(idx.size, q)); sk.append(k[common])
    cat = lambda a: np.concatenate(a).astype(np.int32) if a else np.zeros(0, np.int32)
    # sources per node, ascending rank including me
    holders = member[:, mine]                           # [world, n_mine]
    n_nodes = int(mine.sum())
    counts = holders.sum(0).astype(np.int64)
    node_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    src_rank = np.zeros(int(counts.sum()), np.int32)
    src_k = np.zeros(int(counts.sum()), np.int32)
    fill = node_off[:-1].astype(np.int64).copy()
    s_of_mine = np.nonzero(mine)[0]
    for q in range(world):
        has = holders[q]
        at = fill[has]
        src_rank[at] = q
        if q != rank and q in k_with:
            src_k[at] = k_with[q][s_of_mine[has]]
        fill[has] += 1
    return PeerTables(smax, cat(sx), cat(su), cat(sp), cat(sk), nx.astype(np.int32), nu.astype(np.int32), node_off, src_rank, src_k,
                      np.array(sorted(k_with.keys()), np.int32), local_node)


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


def x_dtype_is_f64(t):
    return t.dtype == torch.float64


class HaloP2P:
    """Halo exchange over NVLink peer memory (csrc/halo_p2p.cu): every rank's receive buffer is torch symmetric memory
    mapped by all ranks of the box; gradients of shared nodes and the loss partials are put into the peers' buffers by
    this rank's own kernels and completed by per-sender flags -- no NCCL call in the step, capturable in a CUDA graph.
    `exchange_grads` can run on a side stream while the tiles that own no shared node still compute."""

    capturable = True

    def __init__(self, tables: PeerTables, device, dtype, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.t = tables
        self.device = torch.device(device)
        self.dtype = dtype
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        rb = 8 if dtype == torch.float64 else 4
        nbytes = int(_lib.lib().hidenn_halo_p2p_bytes(C.c_int(self.world), C.c_int64(max(1, tables.smax)), C.c_int(rb)))
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=self.device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr(), "symmetric memory rendezvous returned unexpected pointers"
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.step = torch.ones(1, dtype=torch.int64, device=self.device)          # device-side step counters (start at 1):
        self.loss_step = torch.ones(1, dtype=torch.int64, device=self.device)     # gradient channel / loss channel
        self.first_done = torch.zeros(1, dtype=torch.int32, device=self.device)   # progress counter of the tile kernel (overlap)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.device)
        self.s_xrow, self.s_urow, self.s_peer, self.s_k = d(tables.send_xrow), d(tables.send_urow), d(tables.send_peer), d(tables.send_k)
        self.n_xrow, self.n_urow, self.n_off = d(tables.node_xrow), d(tables.node_urow), d(tables.node_off)
        self.src_rank, self.src_k, self.wait = d(tables.src_rank), d(tables.src_k), d(tables.wait_ranks)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)            # every rank's buffer is zeroed before anybody puts into it

    def exchange_grads(self, gx, gu, wait_counter=None, wait_target=0):
        """Complete the gradient rows of the shared nodes (in place); gx / gu may be None (frozen Parameter).
        wait_counter / wait_target: first spin on a device counter the tile kernel running on another stream advances
        (hidenn_tri_energy_overlap_*)."""
        t, s = self.t, _lib.stream_ptr(self.device)
        _lib.check(_lib.fn("hidenn_halo_p2p_push", self.dtype)(
            _lib.ptr(gx), _lib.ptr(gu), _lib.ptr(self.s_xrow if gx is not None else self._neg(self.s_xrow)),
            _lib.ptr(self.s_urow if gu is not None else self._neg(self.s_urow)), _lib.ptr(self.s_peer), _lib.ptr(self.s_k),
            C.c_int64(t.send_xrow.size), _lib.ptr(self.peer_ptrs), C.c_int(self.rank), C.c_int(self.world), C.c_int64(max(1, t.smax)),
            _lib.ptr(self.step), _lib.ptr(wait_counter), C.c_uint32(int(wait_target)), s))
        _lib.check(_lib.fn("hidenn_halo_p2p_pull", self.dtype)(
            _lib.ptr(gx), _lib.ptr(gu), _lib.ptr(self.n_xrow if gx is not None else self._neg(self.n_xrow)),
            _lib.ptr(self.n_urow if gu is not None else self._neg(self.n_urow)), _lib.ptr(self.n_off), _lib.ptr(self.src_rank),
            _lib.ptr(self.src_k), C.c_int64(t.node_xrow.size), _lib.ptr(self.wait), C.c_int(t.wait_ranks.size), _lib.ptr(self.buf),
            C.c_int(self.rank), C.c_int(self.world), C.c_int64(max(1, t.smax)), _lib.ptr(self.step), s))

    def _neg(self, rows):
        key = id(rows)
        c = self.__dict__.setdefault("_neg_cache", {})
        if key not in c:
            c[key] = torch.full_like(rows, -1)
        return c[key]

    def exchange_loss(self, out):
        """out[0..2] <- sum over ranks (ascending rank order, identical on every rank); closes the step."""
        _lib.check(_lib.fn("hidenn_halo_p2p_loss", self.dtype)(
            _lib.ptr(out), _lib.ptr(self.peer_ptrs), _lib.ptr(self.buf), C.c_int(self.rank), C.c_int(self.world),
            C.c_int64(max(1, self.t.smax)), _lib.ptr(self.loss_step), _lib.stream_ptr(self.device)))

    def exchange(self, out, gx, gu):
        self.exchange_grads(gx, gu)
        self.exchange_loss(out)


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

    def _launch(self, model, plan, inputs, flags, out, gx, gu, gt, scratch, post=None):
        """Overlapped form: the tiles that own shared nodes run first; while the remaining tiles compute, a side stream
        puts / completes the shared gradient rows over peer memory; the loss partials are exchanged after the final
        reduction.  Needs the peer-memory exchange, a tile-ordered FP64 plan built with the shared nodes as priority
        nodes, gradients requested and no x-dependent traction chain; otherwise the plain sequence (launch, exchange)."""
        halo = self.halo
        nb = plan.info.get("n_first_tiles", 0)
        grads = flags & 3
        if not (isinstance(halo, HaloP2P) and plan.info.get("tile_ordered") and nb > 0 and grads and post is None
                and x_dtype_is_f64(inputs[0]) and not (flags & 16)):
            return super()._launch(model, plan, inputs, flags, out, gx, gu, gt, scratch, post)
        x_free, xb, u_free, ub, consts, t_table = inputs
        n_tiles = plan.info["n_tiles"]
        dev = x_free.device
        main = torch.cuda.current_stream(dev)
        side = self.__dict__.get("_side_stream") or torch.cuda.Stream(device=dev)
        self._side_stream = side
        L = _lib.lib()
        target = int(L.hidenn_tri_plan_overlap_target(plan.handle)) if not self.__dict__.get("two_launch_overlap") else 0
        side.wait_stream(main)
        with _lib.nvtx("hidenn.tri_energy"):
            if target > 0:
                # ONE launch: the shared tiles come first and are counted as they finish; the put / complete kernels on the
                # side stream start when the count is reached and run on an SM the tile kernel leaves free
                _lib.check(L.hidenn_tri_energy_overlap_f64(
                    plan.handle, _lib.ptr(x_free), _lib.ptr(xb), _lib.ptr(u_free), _lib.ptr(ub), _lib.ptr(consts), _lib.ptr(t_table),
                    C.c_int(flags), _lib.ptr(out), _lib.ptr(gx), _lib.ptr(gu), _lib.ptr(gt), _lib.ptr(scratch),
                    _lib.ptr(halo.first_done), C.c_int(self.__dict__.get("reserve_sms", 1)), _lib.ptr(halo.peer_ptrs), _lib.ptr(halo.buf),
                    C.c_int(halo.rank), C.c_int(halo.world), C.c_int64(max(1, halo.t.smax)), _lib.ptr(halo.loss_step), _lib.stream_ptr(dev)))
                with torch.cuda.stream(side), _lib.nvtx("hidenn.halo_exchange"):
                    halo.exchange_grads(gx, gu, halo.first_done, target)
                main.wait_stream(side)
                return
            else:
                def run(t0, t1):
                    _lib.check(L.hidenn_tri_energy_range_f64(
                        plan.handle, _lib.ptr(x_free), _lib.ptr(xb), _lib.ptr(u_free), _lib.ptr(ub), _lib.ptr(consts), _lib.ptr(t_table),
                        C.c_int(flags), _lib.ptr(gx), _lib.ptr(gu), _lib.ptr(gt), _lib.ptr(scratch), C.c_int(t0), C.c_int(t1),
                        _lib.stream_ptr(dev)))
                run(0, nb)                               # tiles owning shared nodes: their rows are final after this launch
                side.wait_stream(main)
                with torch.cuda.stream(side), _lib.nvtx("hidenn.halo_exchange"):
                    halo.exchange_grads(gx, gu)          # overlaps the interior tiles
                if nb < n_tiles:
                    run(nb, n_tiles)
                _lib.check(L.hidenn_tri_energy_finish_f64(plan.handle, _lib.ptr(scratch), _lib.ptr(out), _lib.stream_ptr(dev)))
            main.wait_stream(side)
            halo.exchange_loss(out)


def setup_strip_halo(mesh: meshgen.PlateMesh, boundary_mask: np.ndarray, dirichlet_mask: np.ndarray, device, dtype,
                     group=None, backend: Optional[str] = None):
    """Halo exchange object for a strip mesh.  backend "p2p" (default; HIDENN_HALO=nccl overrides): peer-memory exchange
    (HaloP2P); "nccl": one packed all-reduce (HaloExchange).  `halo.priority_nodes` = local ids of the shared nodes: set
    `model.priority_nodes` to it before the first evaluation so that their tiles run first."""
    import os
    cands = gather_candidates(strip_candidates(mesh), group)
    shared = shared_ids_from_candidates(cands)
    plan = build_halo_plan(mesh.global_node_id, ~boundary_mask, ~dirichlet_mask, shared)
    rank = dist.get_rank(group)
    backend = backend or os.environ.get("HIDENN_HALO", "p2p")
    if backend == "p2p":
        tables = build_peer_tables(cands, shared, rank, mesh.global_node_id, ~boundary_mask, ~dirichlet_mask)
        halo = HaloP2P(tables, device, dtype, group)
    elif backend == "nccl":
        halo = HaloExchange(plan, device, dtype, group)
    else:
        raise ValueError(f"unknown halo backend {backend!r}")
    halo.backend = backend
    halo.priority_nodes = plan.local_node.astype(np.int64)
    wx, wu = owner_weights(plan, shared_owner_from_candidates(cands, shared), rank, int((~boundary_mask).sum()),
                           int((~dirichlet_mask).sum()))
    # row weights for optim.ShardedLBFGS(model.parameters(), weights=halo.row_weights): [node_coords_free, u_free] order
    halo.row_weights = [torch.from_numpy(wx).to(halo.device, dtype), torch.from_numpy(wu).to(halo.device, dtype)]
    return halo
