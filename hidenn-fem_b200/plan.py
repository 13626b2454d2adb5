"""TriPlan: Python handle of the static-topology plan (include/hidenn_b200.h, hidenn_tri_plan_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

INFO_KEYS = ("n_tiles", "elem_visits", "node_visits", "max_local", "max_entries", "scratch", "smem_f64", "smem_f32",
             "n_free_x", "n_free_u", "n_edges", "n_edge_nodes", "plan_bytes", "max_elem", "n_elems", "n_nodes")


def _np(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


class TriPlan:
    """Built once per mesh (connectivity, masks and Neumann edges never change during training;
    /root/reference/src/models.py:248-282 keeps them as buffers)."""

    def __init__(self, connectivity, n_nodes, coords_init, boundary_mask, dirichlet_mask, neumann_edges=None,
                 tile_nodes=0, real_bytes=8, device=None, first_nodes=None, options=0):
        L = _lib.lib()
        conn = _np(connectivity, np.int64).reshape(-1, 3)
        xy = _np(coords_init, np.float64).reshape(-1, 2)
        bm = _np(boundary_mask, np.uint8)
        dm = _np(dirichlet_mask, np.uint8)
        if neumann_edges is None:
            ed = np.zeros((0, 2), np.int64)
        else:
            ed = _np(neumann_edges, np.int64).reshape(-1, 2)
        assert xy.shape[0] == n_nodes and bm.shape[0] == n_nodes and dm.shape[0] == n_nodes
        if device is None or device == -1:
            dev_index = -1
        else:
            device = torch.device(device)
            if device.type != "cuda":
                raise _lib.HidennError("TriPlan needs a CUDA device: the B200 path has no CPU fallback")
            _lib.require_cuda()
            dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self.device_index = dev_index
        self._h = C.c_void_p(0)
        # first_nodes: nodes whose owner tiles are listed first (multi-GPU: the nodes shared with other ranks)
        fn = np.zeros(0, np.int64) if first_nodes is None else _np(first_nodes, np.int64).reshape(-1)
        rc = L.hidenn_tri_plan_create_ex(conn.ctypes.data_as(C.c_void_p), C.c_int64(conn.shape[0]), C.c_int64(n_nodes),
                                         xy.ctypes.data_as(C.c_void_p), bm.ctypes.data_as(C.c_void_p),
                                         dm.ctypes.data_as(C.c_void_p), ed.ctypes.data_as(C.c_void_p),
                                         C.c_int64(ed.shape[0]), fn.ctypes.data_as(C.c_void_p), C.c_int64(fn.shape[0]),
                                         C.c_int(int(options)), C.c_int(int(tile_nodes)), C.c_int(int(real_bytes)),
                                         C.c_int(dev_index), C.byref(self._h))
        _lib.check(rc, "hidenn_tri_plan_create_ex")
        info = (C.c_int64 * 16)()
        _lib.check(L.hidenn_tri_plan_info(self._h, info), "hidenn_tri_plan_info")
        self.info = dict(zip(INFO_KEYS, [int(v) for v in info]))
        lay = (C.c_int64 * 8)()
        _lib.check(L.hidenn_tri_plan_layout(self._h, lay), "hidenn_tri_plan_layout")
        self.info.update(tile_ordered=bool(lay[0] & 1), pairs_only=bool(lay[0] & 2), max_halo=int(lay[1]), edge_visits=int(lay[2]), smem_v8=int(lay[3]),
                         n_pairs=int(lay[4]), pair_entries=int(lay[5]), max_entries9=int(lay[6]), n_first_tiles=int(lay[7]))
        self.info["kernel"] = int(L.hidenn_tri_plan_kernel(self._h))
        loc = (C.c_double * 2)()
        _lib.check(L.hidenn_tri_plan_locality(self._h, loc), "hidenn_tri_plan_locality")
        self.info.update(runs_per_tile=float(loc[0]), local_per_tile=float(loc[1]))
        if self.info["n_tiles"] >= 64 and loc[0] > 0.5 * loc[1]:
            import warnings
            warnings.warn("HiDeNN B200: the node numbering of this mesh has no locality (%.0f contiguous row runs per tile of %.0f "
                          "nodes): the fused kernels run ~3x slower on it.  Pass the mesh through "
                          "hidenn_fem_b200.meshgen.ingest_mesh / reorder_for_locality (keeps elements, corner order and results)."
                          % (loc[0], loc[1]), RuntimeWarning, stacklevel=3)
        self.real_bytes = real_bytes
        self.n_nodes = n_nodes
        self.n_elems = conn.shape[0]

    @property
    def handle(self):
        return self._h

    def slots(self):
        xs = np.empty(self.n_nodes, np.int32)
        us = np.empty(self.n_nodes, np.int32)
        _lib.check(_lib.lib().hidenn_tri_plan_slots(self._h, xs.ctypes.data_as(C.c_void_p), us.ctypes.data_as(C.c_void_p)))
        return xs, us

    def decode(self):
        v = self.info["elem_visits"]
        el = np.empty(v, np.int64)
        nd = np.empty((v, 3), np.int64)
        ow = np.empty(v, np.uint8)
        _lib.check(_lib.lib().hidenn_tri_plan_decode(self._h, el.ctypes.data_as(C.c_void_p), nd.ctypes.data_as(C.c_void_p),
                                                     ow.ctypes.data_as(C.c_void_p)))
        return el, nd, ow

    def tiles(self):
        """(node_off [n_tiles+1], n_owned [n_tiles], nodes [node_visits]): tile membership, owned nodes first."""
        nt = self.info["n_tiles"]
        off = np.empty(nt + 1, np.int64)
        own = np.empty(nt, np.int32)
        nodes = np.empty(self.info["node_visits"], np.int32)
        _lib.check(_lib.lib().hidenn_tri_plan_tiles(self._h, off.ctypes.data_as(C.c_void_p), own.ctypes.data_as(C.c_void_p),
                                                    nodes.ctypes.data_as(C.c_void_p)))
        return off, own, nodes

    def fold_tables(self):
        """Raw fold tables: dict(elem_off, packs, elems, owned_off, entry_off, n_entries) -- see hidenn_tri_plan_fold_tables."""
        nt, ev = self.info["n_tiles"], self.info["elem_visits"]
        _, n_owned, _ = self.tiles()
        elem_off = np.empty(nt + 1, np.int64)
        packs = np.empty(ev, np.uint64)
        elems = np.empty(ev, np.int64)
        owned_off = np.empty(nt + 1, np.int64)
        entry_off = np.empty(int(n_owned.sum()), np.uint32)
        n_entries = np.empty(nt, np.int32)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().hidenn_tri_plan_fold_tables(self._h, P(elem_off), P(packs), P(elems), P(owned_off), P(entry_off), P(n_entries)))
        return dict(elem_off=elem_off, packs=packs, elems=elems, owned_off=owned_off, entry_off=entry_off, n_entries=n_entries)

    def pair_tables(self):
        """Paired layout (tile-ordered FP64 plans): dict(pent_off, packs [entries,2], owned_off, entry_off9, n_entries9, mate)."""
        nt = self.info["n_tiles"]
        _, n_owned, _ = self.tiles()
        pent_off = np.empty(nt + 1, np.int64)
        packs = np.empty((self.info["pair_entries"], 2), np.uint64)
        owned_off = np.empty(nt + 1, np.int64)
        entry_off9 = np.empty(int(n_owned.sum()), np.uint32)
        n_entries9 = np.empty(nt, np.int32)
        mate = np.empty(self.n_elems, np.int32)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().hidenn_tri_plan_pair_tables(self._h, P(pent_off), P(packs), P(owned_off), P(entry_off9), P(n_entries9), P(mate)))
        return dict(pent_off=pent_off, packs=packs, owned_off=owned_off, entry_off9=entry_off9, n_entries9=n_entries9, mate=mate)

    def pipeline(self):
        """Row-block tables of the host-buffer pipeline: dict(rows_x, rows_u, first_need_x, last_own_x, first_need_u, last_own_u)."""
        rows = np.empty(2, np.int32)
        t = [np.empty(64, np.int32) for _ in range(4)]
        _lib.check(_lib.lib().hidenn_tri_plan_pipeline(self._h, rows.ctypes.data_as(C.c_void_p), *[a.ctypes.data_as(C.c_void_p) for a in t]))
        return dict(rows_x=int(rows[0]), rows_u=int(rows[1]), first_need_x=t[0], last_own_x=t[1], first_need_u=t[2], last_own_u=t[3])

    def bank_stats(self, real_bytes=None):
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib().hidenn_tri_plan_bank_stats(self._h, C.c_int(real_bytes or self.real_bytes), out))
        return dict(gather=int(out[0]), gather_ideal=int(out[1]), store=int(out[2]), store_ideal=int(out[3]))

    def stage_stats(self):
        out = (C.c_int64 * 2)()
        _lib.check(_lib.lib().hidenn_tri_plan_stage_stats(self._h, out))
        return dict(passes=int(out[0]), ideal=int(out[1]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().hidenn_tri_plan_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
