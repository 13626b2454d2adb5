"""Model classes -- the drop-in mirror of /root/reference/src/models.py on the B200 kernels.

Same constructor signatures, attribute / Parameter / buffer names and registration order as the
reference, so `state_dict()` round-trips and the unchanged Adam / LBFGS loops of
/root/reference/examples/*.py run on them.  Every forward / energy evaluation calls the
hand-written CUDA kernels through the C-ABI (include/hidenn_b200.h); a CPU tensor raises --
there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .plan import TriPlan

c_i64 = C.c_int64


# ------------------------------------------------------------------------------------------------
# small autograd bridges over the generic kernels
# ------------------------------------------------------------------------------------------------
class _AssembleFn(torch.autograd.Function):
    """coords / u_full (reference models.py:292-305): full[n] = free[slot] or fixed[~slot]."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, free_vals, fixed_vals, model, which):
        plan = model._plan()
        full = torch.empty(model.Nnodes, 2, device=free_vals.device, dtype=free_vals.dtype)
        _lib.check(_lib.fn("hidenn_tri_assemble", free_vals.dtype)(
            plan.handle, C.c_int(which), _lib.ptr(free_vals), _lib.ptr(fixed_vals), _lib.ptr(full), _lib.stream_ptr()))
        ctx.mask = model.free_mask if which == 0 else model.u_free_mask
        return full

    @staticmethod
    def backward(ctx, g):
        return g[ctx.mask], None, None, None


class _TriEvalFn(torch.autograd.Function):
    """forward(x_ref, elem_id) of reference models.py:316-357 with a deterministic VJP."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, x_free, u_free, model, x_ref, elem_id):
        plan = model._plan()
        dt, dev = x_free.dtype, x_free.device
        M = elem_id.shape[0]
        x_ref = x_ref.to(dt).contiguous()
        # x_ref receives no gradient (the reference's forward is differentiable in x_eval; none of its losses use that)
        elem_id = _lib.check_ids(elem_id.to(torch.int64), model.Nelems, "forward(x_eval, elem_id)").contiguous()
        u_h = torch.empty(M, 2, device=dev, dtype=dt)
        det = torch.empty(M, device=dev, dtype=dt)
        G = torch.empty(M, 2, 2, device=dev, dtype=dt)
        xb, ub = model._fixed_pair()
        _lib.check(_lib.fn("hidenn_tri_eval_fwd", dt)(
            plan.handle, _lib.ptr(x_free), _lib.ptr(xb), _lib.ptr(u_free), _lib.ptr(ub), _lib.ptr(x_ref), _lib.ptr(elem_id),
            c_i64(M), _lib.ptr(u_h), _lib.ptr(det), _lib.ptr(G), _lib.stream_ptr()))
        ctx.model = model
        ctx.save_for_backward(x_free, u_free, x_ref, elem_id)
        return u_h, det, G

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, cu, cd, cG):
        x_free, u_free, x_ref, elem_id = ctx.saved_tensors
        model = ctx.model
        plan = model._plan()
        dt, dev = x_free.dtype, x_free.device
        M = elem_id.shape[0]
        Ne = model.Nelems
        cu = None if cu is None else cu.contiguous()
        cd = None if cd is None else cd.contiguous()
        cG = None if cG is None else cG.contiguous()
        row_gx = torch.empty(M, 3, 2, device=dev, dtype=dt)
        row_gu = torch.empty(M, 3, 2, device=dev, dtype=dt)
        xb, ub = model._fixed_pair()
        s = _lib.stream_ptr()
        _lib.check(_lib.fn("hidenn_tri_eval_bwd", dt)(
            plan.handle, _lib.ptr(x_free), _lib.ptr(xb), _lib.ptr(u_free), _lib.ptr(ub), _lib.ptr(x_ref), _lib.ptr(elem_id),
            c_i64(M), _lib.ptr(cu), _lib.ptr(cd), _lib.ptr(cG), _lib.ptr(row_gx), _lib.ptr(row_gu), s))
        # deterministic fold: rows grouped by element (stable sort), elements by node (plan CSR)
        sorted_ids, order = torch.sort(elem_id, stable=True)
        seg = torch.searchsorted(sorted_ids, torch.arange(Ne + 1, device=dev, dtype=torch.int64))
        tmp = torch.empty(Ne, 12, device=dev, dtype=dt)
        gx = torch.zeros_like(x_free)
        gu = torch.zeros_like(u_free)
        _lib.check(_lib.fn("hidenn_tri_fold_rows", dt)(
            plan.handle, _lib.ptr(row_gx), _lib.ptr(row_gu), _lib.ptr(order), _lib.ptr(seg), c_i64(M), _lib.ptr(tmp),
            _lib.ptr(gx), _lib.ptr(gu), s))
        return gx, gu, None, None, None


class NeumannEdgesWrapper:
    """reference models.py:214-226: endpoint coordinates of the selected Neumann edges."""

    def __init__(self, coords, edges):
        self.coords = coords
        self.edges = edges

    def __getitem__(self, idx):
        return self.coords[self.edges[idx, 0]], self.coords[self.edges[idx, 1]]

    def __len__(self):
        return self.edges.shape[0]


class ConnectivityWrapper:
    """reference models.py:228-238: coords[connectivity[idx]]."""

    def __init__(self, coords, connectivity):
        self.coords = coords
        self.connectivity = connectivity

    def __getitem__(self, idx):
        return self.coords[self.connectivity[idx]]

    def __len__(self):
        return self.connectivity.shape[0]


class PiecewiseLinearShapeNN2D(nn.Module):
    """P1-triangle interpolant with trainable nodal values and coordinates
    (reference models.py:241-376).  Calling the class with `grid_x=` / `grid_y=` selects the
    structured tensor-product model the reference defines first under the same name
    (models.py:93-212; exported here as StructuredShapeNN2D)."""

    def __new__(cls, *args, **kwargs):
        if cls is PiecewiseLinearShapeNN2D and ("grid_x" in kwargs or "grid_y" in kwargs):
            from .models_grid import StructuredShapeNN2D
            return StructuredShapeNN2D(*args, **kwargs)
        return super().__new__(cls)

    def __init__(self, node_coords, connectivity, boundary_mask=None, dirichlet_mask=None, u_fixed=None,
                 neumann_edges=None, jinv_transpose=False):
        super().__init__()
        # correct-math switch (keyword only in spirit; default off = the reference's J^-1 in dN/dx): with True
        # every kernel of this model (forward, fused energy, gradients) uses dN/dx = J^-T dN/dxi
        self.jinv_transpose = bool(jinv_transpose)
        self.scale = 1e-5
        self.dim_u = 2
        self.register_buffer("initial_node_coords", node_coords.clone())
        self.Nnodes = node_coords.shape[0]
        self.register_buffer("connectivity", connectivity.long().clone())
        self.Nelems = connectivity.shape[0]
        if boundary_mask is None:
            boundary_mask = torch.zeros(self.Nnodes, dtype=torch.bool)
        self.register_buffer("boundary_mask", boundary_mask.clone())
        free_mask = ~boundary_mask
        self.node_coords_free = nn.Parameter(node_coords[free_mask])
        self.register_buffer("node_coords_fixed", node_coords[boundary_mask])
        self.register_buffer("free_mask", free_mask)
        if dirichlet_mask is None:
            dirichlet_mask = torch.zeros(self.Nnodes, dtype=torch.bool)
        self.register_buffer("dirichlet_mask", dirichlet_mask.clone())
        u_free_mask = ~dirichlet_mask
        self.register_buffer("u_free_mask", u_free_mask)
        # always float32 at construction, like the reference (Q7); .double() converts it
        self.u_free = nn.Parameter(self.scale * torch.randn(int(u_free_mask.sum().item()), self.dim_u))
        if u_fixed is not None:
            self.register_buffer("u_fixed", torch.tensor(u_fixed))
        if neumann_edges is not None:
            self.register_buffer("neumann_edges", neumann_edges)
            self.N_edges = neumann_edges.shape[0]
        self._plans = {}
        self._ufix_cache = None
        self.tile_nodes = 0          # 0 = library default for the dtype
        self.priority_nodes = None   # multi-GPU: local ids of the nodes shared with other ranks (their tiles run first)

    # -- reference properties ---------------------------------------------------------------
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def coords(self):
        self._check_ready()
        return _AssembleFn.apply(self.node_coords_free, self.node_coords_fixed, self, 0)

    @property
    def u_full(self):
        self._check_ready()
        if self.u_fixed is not None:      # AttributeError when constructed with u_fixed=None, like the reference (Q8)
            pass
        return _AssembleFn.apply(self.u_free, self._fixed_pair()[1], self, 1)

    @property
    def domain_elements(self):
        return ConnectivityWrapper(self.coords, self.connectivity)

    @property
    def nm_edges(self):
        return NeumannEdgesWrapper(self.coords, self.neumann_edges)

    # -- plumbing -----------------------------------------------------------------------------
    def _check_ready(self):
        p = self.node_coords_free
        if p.device.type != "cuda":
            raise _lib.HidennError("PiecewiseLinearShapeNN2D: parameters are on %s; the B200 path has no CPU fallback "
                                   "(move the model with .to('cuda'))" % p.device)
        if self.u_free.dtype != p.dtype:
            # the reference fails here too ("Index put requires the source and destination dtypes match", Q7)
            raise RuntimeError("node_coords_free is %s but u_free is %s: call model.double() / model.float()"
                               % (p.dtype, self.u_free.dtype))
        _lib.suffix(p.dtype)

    def _plan(self) -> TriPlan:
        self._check_ready()
        p = self.node_coords_free
        key = (p.device.index if p.device.index is not None else torch.cuda.current_device(), p.dtype, self.tile_nodes)
        plan = self._plans.get(key)
        if plan is None:
            edges = self.neumann_edges if hasattr(self, "neumann_edges") else None
            plan = TriPlan(self.connectivity, self.Nnodes, self.initial_node_coords.double(), self.boundary_mask,
                           self.dirichlet_mask, edges, tile_nodes=self.tile_nodes,
                           real_bytes=8 if p.dtype == torch.float64 else 4, device=torch.device("cuda", key[0]),
                           first_nodes=self.priority_nodes, options=1 if self.jinv_transpose else 0)
            self._plans[key] = plan
        return plan

    def _fixed_pair(self):
        """(node_coords_fixed, u_fixed broadcast to [N_dirichlet,2]) in the parameter dtype."""
        p = self.node_coords_free
        xb = self.node_coords_fixed
        if xb.dtype != p.dtype:
            xb = xb.to(p.dtype)
        uf = self.u_fixed           # AttributeError if absent (reference behaviour)
        key = (uf._version, uf.data_ptr(), p.dtype, p.device)
        if self._ufix_cache is None or self._ufix_cache[0] != key:
            n_fix = self.Nnodes - self.u_free.shape[0]
            ub = uf.to(device=p.device, dtype=p.dtype).expand(n_fix, 2).contiguous() if n_fix > 0 else \
                torch.zeros(0, 2, device=p.device, dtype=p.dtype)
            self._ufix_cache = (key, ub)
        return xb.contiguous(), self._ufix_cache[1]

    # -- forward (reference models.py:316-376) ---------------------------------------------------
    def forward(self, x_eval, elem_id, edge=False):
        self._check_ready()
        if not edge:
            return _TriEvalFn.apply(self.node_coords_free, self.u_free, self, x_eval, elem_id)
        # Edge branch: O(#edge rows) work; differentiable through the assembled arrays
        x_i, x_ip1 = self.nm_edges[elem_id]
        xi = x_eval[:, 0:1]
        N = torch.cat([1.0 - xi, xi], dim=1)
        u_nodes = self.u_full[self.neumann_edges[elem_id]]
        u_h = torch.sum(N.unsqueeze(2) * u_nodes, dim=1)
        ds = torch.norm(x_ip1 - x_i, dim=1)
        return u_h, ds

    def edge_forward_nograd(self, x_eval, edge_id):
        """Kernel version of the edge branch (no autograd): u_h [M,2], ds [M]."""
        plan = self._plan()
        dt, dev = self.dtype, self.device
        M = edge_id.shape[0]
        xi = x_eval.reshape(-1).to(dt).contiguous()
        eid = _lib.check_ids(edge_id.to(torch.int64), self.N_edges, "edge_forward_nograd(x_eval, edge_id)").contiguous()
        u_h = torch.empty(M, 2, device=dev, dtype=dt)
        ds = torch.empty(M, device=dev, dtype=dt)
        xb, ub = self._fixed_pair()
        with torch.cuda.device(dev):
            _lib.check(_lib.fn("hidenn_tri_edge_fwd", dt)(
                plan.handle, _lib.ptr(self.node_coords_free.detach()), _lib.ptr(xb), _lib.ptr(self.u_free.detach()), _lib.ptr(ub),
                _lib.ptr(xi), _lib.ptr(eid), c_i64(M), _lib.ptr(u_h), _lib.ptr(ds), _lib.stream_ptr()))
        return u_h, ds


from .models_grid import PiecewiseLinearShapeNN, StructuredShapeNN2D  # noqa: E402,F401
