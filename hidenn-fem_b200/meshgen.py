"""Synthetic plate meshes (numpy, host side, seeded and strip-consistent).

The reference builds its example-4 mesh with gmsh / meshzoo
(/root/reference/src/mesh.py:8-153 and :155-276); neither is installed here and
mesh generation is outside the hot path.  These generators
produce the same 6-tuple the reference hands to the model
(node_coords, connectivity, geom_boundary_mask, bc_mask, mn_mask, neumann_edges;
/root/reference/src/mesh.py:146-153) for the 2x1 plate with three circular holes
of /root/reference/examples/example4.py:14-23, at any resolution.

All randomness is a counter-based hash of the *global* grid index, so a column
strip generated on one rank is bit-identical to the same columns of the global
mesh (needed by the multi-GPU weak-scaling runs).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

DEFAULT_HOLES = ((0.5, 0.7, 0.12), (1.0, 0.3, 0.15), (1.4, 0.6, 0.1))  # example4.py:16


@dataclass
class PlateMesh:
    """The reference's mesh 6-tuple plus bookkeeping for partitioned runs."""

    node_coords: np.ndarray        # [Nn,2] float64
    connectivity: np.ndarray       # [Ne,3] int64
    boundary_mask: np.ndarray      # [Nn] bool  (geometric boundary: rectangle + hole rims)
    dirichlet_mask: np.ndarray     # [Nn] bool  (x == 0)
    neumann_mask: np.ndarray       # [Nn] bool  (x == length)
    neumann_edges: np.ndarray      # [Ned,2] int64, each row sorted ascending
    global_node_id: np.ndarray     # [Nn] int64: ix*ny+iy of the generating grid
    meta: dict = field(default_factory=dict)

    def as_tuple(self):
        return (self.node_coords, self.connectivity, self.boundary_mask,
                self.dirichlet_mask, self.neumann_mask, self.neumann_edges)


def _hash_u01(a: np.ndarray, seed: int) -> np.ndarray:
    """splitmix64 finaliser -> uniform [0,1) doubles; vectorised, stateless."""
    with np.errstate(over="ignore"):
        z = a.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(seed + 1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _part1by1(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
    return v


def morton2(ix: np.ndarray, iy: np.ndarray) -> np.ndarray:
    return _part1by1(ix) | (_part1by1(iy) << np.uint64(1))


def plate_mesh(nx: int, ny: int, length: float = 2.0, height: float = 1.0,
               holes: Sequence[Tuple[float, float, float]] = DEFAULT_HOLES,
               jitter: float = 0.0, diag: str = "alt", seed: int = 0,
               ordering: str = "natural", perm_seed: int = 2,
               col_range: Optional[Tuple[int, int]] = None) -> PlateMesh:
    """nx x ny *nodes* on [0,length]x[0,height], every cell split in two triangles.

    jitter   : interior nodes moved by uniform +-jitter*cell (0 = structured case of SURVEY §8(d) C4)
    diag     : "alt" alternating diagonal (checkerboard), "random" hashed coin per cell
    ordering : "natural" (ix-major), "morton" (nodes and elements along a Z curve),
               "random" (seeded permutation of node ids and of element order)
    col_range: (c0,c1) keep only cells with c0 <= ix < c1 (a strip for one rank);
               node jitter / diagonals / hole removal are identical to the global mesh.
    Elements with any node strictly inside a hole are removed; nodes left without an
    element are dropped; rim nodes of removed elements join the geometric boundary mask
    (same rule as /root/reference/src/mesh.py:202-215).
    """
    assert nx >= 2 and ny >= 2
    c0, c1 = (0, nx - 1) if col_range is None else col_range
    assert 0 <= c0 < c1 <= nx - 1
    hx, hy = length / (nx - 1), height / (ny - 1)

    # build one extra cell column on each side so that hole-rim flags of interface nodes see the
    # neighbouring strip's removed elements (keeps boundary masks identical to the global mesh)
    own0, own1 = c0, c1
    c0, c1 = max(c0 - 1, 0), min(c1 + 1, nx - 1)
    ixs = np.arange(c0, c1 + 1, dtype=np.int64)
    iys = np.arange(ny, dtype=np.int64)
    IX, IY = np.meshgrid(ixs, iys, indexing="ij")           # [ncol+1, ny]
    gid = (IX * ny + IY).ravel()
    x = IX.ravel() * hx
    y = IY.ravel() * hy
    # exact end coordinates (so boundary tests are exact)
    x[IX.ravel() == nx - 1] = length
    y[IY.ravel() == ny - 1] = height
    if jitter > 0.0:
        interior = (IX.ravel() > 0) & (IX.ravel() < nx - 1) & (IY.ravel() > 0) & (IY.ravel() < ny - 1)
        jx = (2.0 * _hash_u01(gid * 2, seed) - 1.0) * jitter * hx
        jy = (2.0 * _hash_u01(gid * 2 + 1, seed) - 1.0) * jitter * hy
        x = np.where(interior, x + jx, x)
        y = np.where(interior, y + jy, y)

    ncol = c1 - c0
    nyl = ny
    # local (strip) node index of grid node (ix,iy)
    def lid(ix, iy):
        return (ix - c0) * nyl + iy

    CX, CY = np.meshgrid(np.arange(c0, c1, dtype=np.int64), np.arange(ny - 1, dtype=np.int64), indexing="ij")
    cx, cy = CX.ravel(), CY.ravel()
    n00, n10 = lid(cx, cy), lid(cx + 1, cy)
    n01, n11 = lid(cx, cy + 1), lid(cx + 1, cy + 1)
    if diag == "alt":
        flip = ((cx + cy) & 1).astype(bool)
    elif diag == "random":
        flip = _hash_u01(cx * (ny - 1) + cy + (1 << 40), seed) < 0.5
    elif diag == "none":
        flip = np.zeros(cx.shape, dtype=bool)
    else:
        raise ValueError(diag)
    # diagonal n00-n11 (flip False) or n10-n01 (flip True); both triangles counter-clockwise
    ta = np.where(flip[:, None], np.stack([n00, n10, n01], 1), np.stack([n00, n10, n11], 1))
    tb = np.where(flip[:, None], np.stack([n10, n11, n01], 1), np.stack([n00, n11, n01], 1))
    conn = np.empty((2 * cx.size, 3), dtype=np.int64)
    conn[0::2] = ta
    conn[1::2] = tb
    cell_of_elem = np.repeat(np.arange(cx.size), 2)

    # hole removal uses the *unjittered* grid position so strips agree and rims stay on the lattice
    x0g, y0g = IX.ravel() * hx, IY.ravel() * hy
    inside = np.zeros(x.size, dtype=bool)
    for (hcx, hcy, r) in holes:
        inside |= (x0g - hcx) ** 2 + (y0g - hcy) ** 2 <= r * r
    elem_ok = ~inside[conn].any(axis=1)
    rim = np.zeros(x.size, dtype=bool)
    bad = conn[~elem_ok].ravel()
    rim[bad] = True
    rim &= ~inside
    own = (cx[cell_of_elem] >= own0) & (cx[cell_of_elem] < own1)
    conn = conn[elem_ok & own]
    cell_of_elem = cell_of_elem[elem_ok & own]

    used = np.zeros(x.size, dtype=bool)
    used[conn.ravel()] = True
    keep = used
    new_id = -np.ones(x.size, dtype=np.int64)
    new_id[keep] = np.arange(int(keep.sum()))
    conn = new_id[conn]
    x, y, gid, rim = x[keep], y[keep], gid[keep], rim[keep]
    gix, giy = gid // ny, gid % ny

    boundary = rim | (gix == 0) | (gix == nx - 1) | (giy == 0) | (giy == ny - 1)
    dirichlet = gix == 0
    neumann = gix == nx - 1

    # ordering
    ecx, ecy = cx[cell_of_elem], cy[cell_of_elem]
    if ordering == "natural":
        pass
    elif ordering == "morton":
        nperm = np.argsort(morton2(gix, giy), kind="stable")
        inv = np.empty_like(nperm)
        inv[nperm] = np.arange(nperm.size)
        x, y, gid, boundary, dirichlet, neumann = (a[nperm] for a in (x, y, gid, boundary, dirichlet, neumann))
        conn = inv[conn]
        eperm = np.argsort(morton2(ecx, ecy), kind="stable")
        conn = conn[eperm]
    elif ordering == "random":
        rng = np.random.default_rng(perm_seed)
        nperm = rng.permutation(x.size)
        inv = np.empty_like(nperm)
        inv[nperm] = np.arange(nperm.size)
        x, y, gid, boundary, dirichlet, neumann = (a[nperm] for a in (x, y, gid, boundary, dirichlet, neumann))
        conn = inv[conn]
        conn = conn[rng.permutation(conn.shape[0])]
    else:
        raise ValueError(ordering)

    # Neumann edges = unique element edges with both ends on x == length
    # (/root/reference/src/mesh.py:125-134), restricted first to candidate edges for speed.
    e_all = np.concatenate([conn[:, [0, 1]], conn[:, [1, 2]], conn[:, [2, 0]]], axis=0)
    cand = neumann[e_all[:, 0]] & neumann[e_all[:, 1]]
    e_c = np.sort(e_all[cand], axis=1)
    neumann_edges = np.unique(e_c, axis=0) if e_c.size else np.zeros((0, 2), dtype=np.int64)

    coords = np.stack([x, y], axis=1).astype(np.float64)
    return PlateMesh(coords, np.ascontiguousarray(conn), boundary, dirichlet, neumann,
                     neumann_edges.astype(np.int64), gid,
                     meta=dict(nx=nx, ny=ny, length=length, height=height, jitter=jitter, diag=diag,
                               seed=seed, ordering=ordering, col_range=(own0, own1)))


def plate_dims_for_elements(n_elems: int, aspect: float = 2.0, hole_fraction: float = 0.0745):
    """(nx, ny) nodes such that the plate keeps >= n_elems triangles after hole removal."""
    cells = n_elems / 2.0 / (1.0 - hole_fraction)
    ncy = int(np.ceil(np.sqrt(cells / aspect)))
    ncx = int(np.ceil(aspect * ncy))
    return ncx + 1, ncy + 1


def invert_some_elements(conn: np.ndarray, fraction: float, seed: int = 0) -> np.ndarray:
    """Swap two corners of a hashed subset of elements (negative det J) -- parity edge case."""
    out = conn.copy()
    pick = _hash_u01(np.arange(conn.shape[0]) + (1 << 41), seed) < fraction
    out[pick, 0], out[pick, 1] = conn[pick, 1], conn[pick, 0]
    return out


def reorder_for_locality(node_coords: np.ndarray, connectivity: np.ndarray, boundary_mask: np.ndarray,
                         dirichlet_mask: np.ndarray, neumann_edges: Optional[np.ndarray] = None, bits: int = 20,
                         mode: str = "tiles", tile_nodes: int = 0):
    """Mesh ingestion helper for meshes with arbitrary numbering (gmsh / meshzoo output, the 6-tuple of
    /root/reference/src/mesh.py:125-153, 252-276): renumber the nodes for locality and list the elements by their
    smallest new node id.  The fused kernels are correct for any numbering but read Parameter rows through 32-byte
    sectors, so a numbering without locality runs ~3.5x slower (profiles/README.md).

    mode="tiles" (default): the native `hidenn_tri_locality_order` -- the plan's own recursive coordinate bisection,
    nodes numbered tile by tile (inside a tile by free/fixed class, then by valence).  The plan recognises this
    numbering and stages every tile's rows with bulk copies (the fastest path).  mode="morton": Z curve of the coordinates.

    Returns (node_coords, connectivity, boundary_mask, dirichlet_mask, neumann_edges, new_to_old, elem_new_to_old):
    `new_to_old[i]` is the original index of new node i (to map results back: u_old[new_to_old] = u_new); corner
    order inside every element is kept (the reference's results depend on it), and so is the orientation of every
    Neumann edge."""
    xy = np.ascontiguousarray(node_coords, dtype=np.float64)
    conn = np.ascontiguousarray(connectivity, dtype=np.int64)
    if mode == "tiles":
        import ctypes as C
        from . import _lib
        bm = np.ascontiguousarray(boundary_mask, dtype=np.uint8)
        dm = np.ascontiguousarray(dirichlet_mask, dtype=np.uint8)
        new_to_old = np.empty(xy.shape[0], np.int64)
        elem_order = np.empty(conn.shape[0], np.int64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        ed = np.zeros((0, 2), np.int64) if neumann_edges is None else np.ascontiguousarray(neumann_edges, dtype=np.int64).reshape(-1, 2)
        _lib.check(_lib.lib().hidenn_tri_locality_order(P(conn), C.c_int64(conn.shape[0]), C.c_int64(xy.shape[0]), P(xy), P(bm), P(dm),
                                                        P(ed), C.c_int64(ed.shape[0]), C.c_int(int(tile_nodes)), P(new_to_old),
                                                        P(elem_order)),
                   "hidenn_tri_locality_order")
        old_to_new = np.empty_like(new_to_old)
        old_to_new[new_to_old] = np.arange(new_to_old.size)
        conn_new = old_to_new[conn][elem_order]
    elif mode == "morton":
        lo, hi = xy.min(0), xy.max(0)
        span = np.where(hi > lo, hi - lo, 1.0)
        q = np.minimum(((xy - lo) / span * ((1 << bits) - 1)).astype(np.uint64), np.uint64((1 << bits) - 1))
        new_to_old = np.argsort(morton2(q[:, 0], q[:, 1]), kind="stable")
        old_to_new = np.empty_like(new_to_old)
        old_to_new[new_to_old] = np.arange(new_to_old.size)
        conn_new = old_to_new[conn]
        elem_order = np.argsort(conn_new.min(1), kind="stable")
        conn_new = conn_new[elem_order]
    else:
        raise ValueError(mode)
    edges = None if neumann_edges is None else old_to_new[np.asarray(neumann_edges, dtype=np.int64)]
    return (np.asarray(node_coords)[new_to_old], conn_new, np.asarray(boundary_mask)[new_to_old],
            np.asarray(dirichlet_mask)[new_to_old], edges, new_to_old, elem_order)


def reorder_mesh(m: PlateMesh, mode: str = "tiles", tile_nodes: int = 0) -> PlateMesh:
    """`reorder_for_locality` applied to a PlateMesh (all per-node arrays follow the new numbering)."""
    xy, conn, bm, dm, ed, n2o, _ = reorder_for_locality(m.node_coords, m.connectivity, m.boundary_mask, m.dirichlet_mask,
                                                        m.neumann_edges, mode=mode, tile_nodes=tile_nodes)
    return PlateMesh(xy, conn, bm, dm, m.neumann_mask[n2o], ed, m.global_node_id[n2o], meta=dict(m.meta, ordering=mode))


def ingest_mesh(node_coords, connectivity, length: float = 2.0, height: float = 1.0, boundaries: Optional[dict] = None,
                tol: float = 1e-6, reorder: Optional[str] = "tiles", tile_nodes: int = 0, dtype=None):
    """From a raw triangulation (any generator: gmsh, meshzoo, Delaunay; any numbering) to the reference's 6-tuple
    `(node_coords, connectivity, geom_boundary_mask, bc_mask, mn_mask, neumann_edges)` -- the part of
    /root/reference/src/mesh.py:70-153 and :202-276 that follows the mesher:

      geom_boundary_mask  nodes on the boundary of the domain incl. hole rims (topological: edges of exactly one
                          element, native `hidenn_mesh_boundary_nodes`; the reference asks gmsh / tests the hole radius)
      bc_mask / mn_mask   faces "up" / "down" / "left" / "right" with condition 1 (Dirichlet) / 2 (Neumann), |coordinate
                          - face| < tol as in mesh.py:108-124
      neumann_edges       the sorted-unique element edges whose two nodes are both in mn_mask (mesh.py:125-134)

    then (reorder="tiles" | "morton" | None) renumbered for locality with `reorder_for_locality`.  Returns torch tensors
    like the reference (`dtype` of the coordinates: float32 there; default keeps the input's), plus `new_to_old`
    (identity when reorder is None)."""
    import ctypes as C
    import torch
    from . import _lib
    xy = np.ascontiguousarray(np.asarray(node_coords), dtype=np.float64)
    conn = np.ascontiguousarray(np.asarray(connectivity), dtype=np.int64).reshape(-1, 3)
    boundaries = boundaries or {"left": 1, "right": 2}
    geom = np.zeros(xy.shape[0], np.uint8)
    _lib.check(_lib.lib().hidenn_mesh_boundary_nodes(conn.ctypes.data_as(C.c_void_p), C.c_int64(conn.shape[0]), C.c_int64(xy.shape[0]),
                                                     geom.ctypes.data_as(C.c_void_p)), "hidenn_mesh_boundary_nodes")
    geom = geom.astype(bool)
    bc = np.zeros(xy.shape[0], bool)
    mn = np.zeros(xy.shape[0], bool)
    face_of = {"up": (1, height), "down": (1, 0.0), "left": (0, 0.0), "right": (0, length)}
    for face, cond in boundaries.items():
        if cond == 0 or face not in face_of:
            continue
        ax, val = face_of[face]
        on = np.abs(xy[:, ax] - val) < tol
        if cond == 1:
            bc |= on
        elif cond == 2:
            mn |= on
    e_all = np.concatenate([conn[:, [0, 1]], conn[:, [1, 2]], conn[:, [2, 0]]], axis=0)
    cand = mn[e_all[:, 0]] & mn[e_all[:, 1]]                     # restrict first: the unique over all edges is the slow part
    e_c = np.sort(e_all[cand], axis=1)
    edges = np.unique(e_c, axis=0) if e_c.size else np.zeros((0, 2), np.int64)
    n2o = np.arange(xy.shape[0])
    if reorder:
        xy, conn, geom, bc, edges, n2o, _ = reorder_for_locality(xy, conn, geom, bc, edges, mode=reorder, tile_nodes=tile_nodes)
        mn = mn[n2o]
    out_dt = dtype or (torch.float32 if np.asarray(node_coords).dtype == np.float32 else torch.float64)
    T = torch.tensor
    return (T(xy, dtype=out_dt), T(conn), T(geom), T(bc), T(mn), T(np.ascontiguousarray(edges), dtype=torch.long), n2o)
