"""1D and structured-2D models -- drop-in mirrors of /root/reference/src/models.py:6-212 on the
grid kernels (include/hidenn_b200_grid.h), plus the fused 1D bar energy of
/root/reference/examples/example3.py:27-70.

Constructor signatures, Parameter / buffer names and registration order follow the reference
so `state_dict()` round-trips.  `forward` stays differentiable w.r.t. `x_eval` to second order
(the reference's example3 calls `autograd.grad(u, xq, create_graph=True)`).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

c_i64 = C.c_int64


def _require_cuda(t, what):
    if t.device.type != "cuda":
        raise _lib.HidennError(f"{what}: tensors are on {t.device}; the B200 path has no CPU fallback")
    _lib.suffix(t.dtype)


_scratch_cache = {}


def _scratch_1d(n, like):
    """Scan / reduction scratch (hidenn_1d_scratch_size), cached per size, device, dtype and stream."""
    key = (n, like.device, like.dtype, torch.cuda.current_stream(like.device).cuda_stream)
    t = _scratch_cache.get(key)
    if t is None:
        L = _lib.lib()
        L.hidenn_1d_scratch_size.restype = C.c_int64
        t = torch.empty(int(L.hidenn_1d_scratch_size(c_i64(n))), device=like.device, dtype=like.dtype)
        if len(_scratch_cache) > 64:
            _scratch_cache.clear()
        _scratch_cache[key] = t
    return t


# ------------------------------------------------------------------------------------------------
# r-adaptive grid: softplus -> clamp -> cumsum -> normalise (models.py:45-53)
# ------------------------------------------------------------------------------------------------
class _GridFn(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, p, x0, xN):
        _require_cuda(p, "grid")
        p = p.contiguous()
        n = p.shape[0]
        dt = p.dtype
        x0c, xNc = x0.to(dt).contiguous(), xN.to(dt).contiguous()
        grid = torch.empty(n + 1, device=p.device, dtype=dt)
        cum = torch.empty(n, device=p.device, dtype=dt)
        sc = _scratch_1d(n, p)
        _lib.check(_lib.fn("hidenn_1d_grid_fwd", dt)(_lib.ptr(p), c_i64(n), _lib.ptr(x0c), _lib.ptr(xNc), _lib.ptr(grid),
                                                     _lib.ptr(cum), _lib.ptr(sc), _lib.stream_ptr()))
        ctx.save_for_backward(p, cum, x0c, xNc)
        return grid

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, dgrid):
        p, cum, x0c, xNc = ctx.saved_tensors
        n = p.shape[0]
        dp = torch.empty_like(p)
        sc = _scratch_1d(n, p)
        _lib.check(_lib.fn("hidenn_1d_grid_bwd", p.dtype)(_lib.ptr(dgrid.contiguous()), _lib.ptr(p), _lib.ptr(cum), _lib.ptr(x0c),
                                                          _lib.ptr(xNc), c_i64(n), _lib.ptr(dp), _lib.ptr(sc), _lib.stream_ptr()))
        return dp, None, None


def _fold_1d(rows, elem, N):
    """Deterministic fold of per-row (du_e, du_e+1, dg_e, dg_e+1) to nodes: stable sort by element."""
    dt, dev = rows.dtype, rows.device
    e64 = elem.to(torch.int64)
    sorted_e, order = torch.sort(e64, stable=True)
    seg = torch.searchsorted(sorted_e, torch.arange(N, device=dev, dtype=torch.int64))     # N-1 elements -> N bounds
    tmp = torch.empty(max(N - 1, 1), 4, device=dev, dtype=dt)
    du = torch.empty(N, device=dev, dtype=dt)
    dg = torch.empty(N, device=dev, dtype=dt)
    _lib.check(_lib.fn("hidenn_1d_fold_rows", dt)(_lib.ptr(rows), _lib.ptr(order), _lib.ptr(seg), c_i64(N), _lib.ptr(tmp),
                                                  _lib.ptr(du), _lib.ptr(dg), _lib.stream_ptr()))
    return du, dg


class _Interp1DFn(torch.autograd.Function):
    """u(x) of models.py:70-90.  backward is itself a differentiable Function so that
    d u / d x (the element slope) can be differentiated again w.r.t. grid and u (example3.py:56)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, grid, u_full, x):
        _require_cuda(grid, "forward")
        dt = grid.dtype
        xs = x.to(dt).contiguous()
        g, uf = grid.contiguous(), u_full.to(dt).contiguous()
        M = xs.numel()
        u = torch.empty(xs.shape, device=g.device, dtype=dt)
        elem = torch.empty(M, device=g.device, dtype=torch.int32)
        _lib.check(_lib.fn("hidenn_1d_interp_fwd", dt)(_lib.ptr(g), c_i64(g.shape[0]), _lib.ptr(uf), _lib.ptr(xs), c_i64(M),
                                                       _lib.ptr(u), _lib.ptr(elem), _lib.ptr(None), _lib.stream_ptr()))
        ctx.save_for_backward(grid, u_full, x, elem)
        return u

    @staticmethod
    @_lib.on_device
    def backward(ctx, r):
        grid, u_full, x, elem = ctx.saved_tensors
        dgrid, du, dx = _Interp1DBwdFn.apply(grid, u_full, x, elem, r)
        return dgrid, du, dx


class _Interp1DBwdFn(torch.autograd.Function):
    """(d grid, d u_full, d x) = VJP of the interpolation for cotangent r; its own backward handles the
    cotangent of d x (= r * slope), which is what the double-backward of example3.py needs."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, grid, u_full, x, elem, r):
        dt = grid.dtype
        g, uf, xs, rr = grid.contiguous(), u_full.to(dt).contiguous(), x.to(dt).contiguous(), r.to(dt).contiguous()
        M = xs.numel()
        rows = torch.empty(M, 4, device=g.device, dtype=dt)
        dx = torch.empty(xs.shape, device=g.device, dtype=dt)
        _lib.check(_lib.fn("hidenn_1d_interp_bwd", dt)(_lib.ptr(g), c_i64(g.shape[0]), _lib.ptr(uf), _lib.ptr(xs), _lib.ptr(elem),
                                                       _lib.ptr(rr), _lib.ptr(None), c_i64(M), _lib.ptr(rows), _lib.ptr(dx),
                                                       _lib.stream_ptr()))
        du, dg = _fold_1d(rows, elem, g.shape[0])
        ctx.save_for_backward(grid, u_full, x, elem, r)
        return dg, du.to(u_full.dtype), dx.to(x.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, g_dgrid, g_du, g_dx):
        grid, u_full, x, elem, r = ctx.saved_tensors
        if g_dgrid is not None or g_du is not None:
            # Only the cotangent of d x is supported at second order (that is all the reference's losses use).
            if (g_dgrid is not None and bool(g_dgrid.abs().max() > 0)) or (g_du is not None and bool(g_du.abs().max() > 0)):
                raise NotImplementedError("second-order derivatives through d grid / d u are not implemented")
        if g_dx is None:
            return None, None, None, None, None
        dt = grid.dtype
        g, uf, xs = grid.contiguous(), u_full.to(dt).contiguous(), x.to(dt).contiguous()
        M = xs.numel()
        # d x = r * slope(grid, u):  cotangent on the slope is g_dx * r; cotangent on r is g_dx * slope
        rs = (g_dx.to(dt) * r.to(dt)).contiguous()
        rows = torch.empty(M, 4, device=g.device, dtype=dt)
        _lib.check(_lib.fn("hidenn_1d_interp_bwd", dt)(_lib.ptr(g), c_i64(g.shape[0]), _lib.ptr(uf), _lib.ptr(xs), _lib.ptr(elem),
                                                       _lib.ptr(None), _lib.ptr(rs), c_i64(M), _lib.ptr(rows), _lib.ptr(None),
                                                       _lib.stream_ptr()))
        du, dg = _fold_1d(rows, elem, g.shape[0])
        slope = torch.empty(xs.shape, device=g.device, dtype=dt)
        utmp = torch.empty(xs.shape, device=g.device, dtype=dt)
        _lib.check(_lib.fn("hidenn_1d_interp_fwd", dt)(_lib.ptr(g), c_i64(g.shape[0]), _lib.ptr(uf), _lib.ptr(xs), c_i64(M),
                                                       _lib.ptr(utmp), _lib.ptr(None), _lib.ptr(slope), _lib.stream_ptr()))
        return dg, du.to(u_full.dtype), None, None, (g_dx.to(dt) * slope).to(r.dtype)


class PiecewiseLinearShapeNN(nn.Module):
    """1D P1 interpolant with optional r-adaptive nodes (reference models.py:6-90)."""

    def __init__(self, node_coords, r_adapt=False, u0=None, uN=None):
        super().__init__()
        self.N = len(node_coords)
        self.r_adapt = r_adapt
        self.register_buffer("x0", node_coords[0:1])
        self.register_buffer("xN", node_coords[-1:])
        if self.r_adapt and self.N > 2:
            self.x_increments = nn.Parameter(node_coords[1:] - node_coords[:-1])
        else:
            self.register_buffer("x_inner", node_coords[1:-1])
        if u0 is not None:
            self.register_buffer("u0_fixed", torch.tensor([u0], dtype=torch.float32))
        else:
            self.u0_fixed = None
        if uN is not None:
            self.register_buffer("uN_fixed", torch.tensor([uN], dtype=torch.float32))
        else:
            self.uN_fixed = None
        if (self.u0_fixed is not None) and (self.uN_fixed is not None):
            self.u = nn.Parameter(torch.zeros(self.N - 2))
        elif (self.u0_fixed is not None) ^ (self.uN_fixed is not None):
            self.u = nn.Parameter(torch.zeros(self.N - 1))
        else:
            self.u = nn.Parameter(torch.zeros(self.N))
        self.epsilon = 1e-10

    @property
    def grid(self):
        if self.r_adapt and self.N > 2:
            return _GridFn.apply(self.x_increments, self.x0, self.xN)
        return torch.cat([self.x0, self.x_inner, self.xN], dim=0)

    @property
    def u_full(self):
        if self.u0_fixed is not None and self.uN_fixed is not None:
            return torch.cat([self.u0_fixed, self.u.view(-1), self.uN_fixed])
        elif self.u0_fixed is not None:
            return torch.cat([self.u0_fixed, self.u.view(-1)])
        elif self.uN_fixed is not None:
            return torch.cat([self.u, self.uN_fixed])
        return self.u.view(-1)

    def lookup(self, x_eval):
        """Element index of every point, bit-exact with clamp(searchsorted(grid,x)-1, 0, N-2) (models.py:73-74)."""
        grid = self.grid.detach().contiguous()
        _require_cuda(grid, "lookup")
        xs = x_eval.to(grid.dtype).contiguous()
        out = torch.empty(xs.numel(), device=grid.device, dtype=torch.int32)
        _lib.check(_lib.fn("hidenn_1d_lookup", grid.dtype)(_lib.ptr(grid), c_i64(grid.shape[0]), _lib.ptr(xs), c_i64(xs.numel()),
                                                           _lib.ptr(out), _lib.stream_ptr()))
        return out.reshape(xs.shape)

    def forward(self, x_eval):
        return _Interp1DFn.apply(self.grid, self.u_full, x_eval)


# ------------------------------------------------------------------------------------------------
# fused bar energy (examples/example3.py:27-70)
# ------------------------------------------------------------------------------------------------
class _BarEnergyFn(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, grid, u_full, xi, wi, E, b_table, state):
        _require_cuda(grid, "bar_energy")
        dt = grid.dtype
        g, uf = grid.contiguous(), u_full.to(dt).contiguous()
        N = g.shape[0]
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        loss = torch.empty(1, device=g.device, dtype=dt)
        du = torch.empty(N, device=g.device, dtype=dt) if need else None
        dg = torch.empty(N, device=g.device, dtype=dt) if need else None
        flag = torch.empty(1, device=g.device, dtype=torch.int32)
        sc = _scratch_1d(N, g)
        Ec = C.c_double(float(E)) if dt == torch.float64 else C.c_float(float(E))
        _lib.check(_lib.fn("hidenn_1d_bar_energy", dt)(_lib.ptr(g), c_i64(N), _lib.ptr(uf), _lib.ptr(xi), _lib.ptr(wi),
                                                       C.c_int(int(xi.numel())), Ec, _lib.ptr(b_table), C.c_int(1 if need else 0),
                                                       _lib.ptr(loss), _lib.ptr(du), _lib.ptr(dg), _lib.ptr(flag), _lib.ptr(sc),
                                                       _lib.stream_ptr()))
        state.note_flag(flag)
        ctx.save_for_backward(dg, du)
        ctx.udtype = u_full.dtype
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, go):
        dg, du = ctx.saved_tensors
        if dg is None:
            return None, None, None, None, None, None, None
        g1 = go.reshape(1).to(dg.dtype).contiguous()
        for t in (dg, du):       # grad_output applied on device; every block exits at once when it is 1
            _lib.check(_lib.fn("hidenn_scale_inplace", t.dtype)(_lib.ptr(t), c_i64(t.numel()), _lib.ptr(g1), _lib.stream_ptr()))
        return dg, du.to(ctx.udtype), None, None, None, None, None


class _BarStepFn(torch.autograd.Function):
    """Fused r-adaptive bar step (hidenn_1d_bar_step_*): grid, energy and both gradients in three launches; backward only
    applies grad_output (device side, exits at once when it is 1)."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, p, u, x0, xN, u0, uN, xi, wi, E, b_table, state):
        _require_cuda(p, "bar_energy")
        dt, dev = p.dtype, p.device
        pc, uc = p.contiguous(), u.to(dt).contiguous().reshape(-1)
        n = pc.shape[0]
        x0c, xNc = x0.to(dt).contiguous(), xN.to(dt).contiguous()
        u0c = None if u0 is None else u0.to(device=dev, dtype=dt).contiguous()
        uNc = None if uN is None else uN.to(device=dev, dtype=dt).contiguous()
        loss = torch.empty(1, device=dev, dtype=dt)
        dp, du, gam = torch.empty_like(pc), torch.empty_like(uc), torch.empty_like(pc)
        flag = torch.empty(1, device=dev, dtype=torch.int32)
        key = ("bar_step", n, dev, dt, torch.cuda.current_stream(dev).cuda_stream)
        sc = _scratch_cache.get(key)
        if sc is None:
            sc = torch.zeros(int(_lib.lib().hidenn_1d_bar_step_scratch(c_i64(n))), device=dev, dtype=dt)     # tickets start at zero
            _scratch_cache[key] = sc
        Ec = C.c_double(float(E)) if dt == torch.float64 else C.c_float(float(E))
        _lib.check(_lib.fn("hidenn_1d_bar_step", dt)(
            _lib.ptr(pc), c_i64(n), _lib.ptr(x0c), _lib.ptr(xNc), _lib.ptr(uc), _lib.ptr(u0c), _lib.ptr(uNc), _lib.ptr(xi), _lib.ptr(wi),
            C.c_int(int(xi.numel())), Ec, _lib.ptr(b_table), _lib.ptr(loss), _lib.ptr(dp), _lib.ptr(du), _lib.ptr(gam), _lib.ptr(flag),
            _lib.ptr(sc), _lib.stream_ptr()))
        state.note_flag(flag)
        ctx.save_for_backward(dp, du)
        ctx.ushape, ctx.udtype = u.shape, u.dtype
        ctx.used = False
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, go):
        if ctx.used:
            raise RuntimeError("bar_energy_loss: backward called twice on the same loss (the gradient buffers are scaled in place); "
                               "evaluate the loss again instead")
        ctx.used = True
        dp, du = ctx.saved_tensors
        g1 = go.reshape(1).to(dp.dtype).contiguous()
        _lib.check(_lib.fn("hidenn_scale_inplace2", dp.dtype)(_lib.ptr(dp), c_i64(dp.numel()), _lib.ptr(du), c_i64(du.numel()),
                                                              _lib.ptr(g1), _lib.stream_ptr()))
        return dp, du.reshape(ctx.ushape).to(ctx.udtype), None, None, None, None, None, None, None, None, None


class _FlagState:
    """Deferred check of the kernel's degenerate-lookup flag (no host sync on the hot path)."""

    def __init__(self):
        self.pending = []
        self.free = []
        self.captured = []

    def note_flag(self, flag):
        if torch.cuda.is_current_stream_capturing():
            # graph.GraphedStep: the copy becomes a node of the graph; check() reads the pinned word after replays
            host = torch.empty(1, dtype=torch.int32).pin_memory()
            host.zero_()
            host.copy_(flag, non_blocking=True)
            self.captured.append(host)
            return
        if self.free:
            host, ev = self.free.pop()
        else:
            host, ev = torch.empty(1, dtype=torch.int32).pin_memory(), torch.cuda.Event()
        host.copy_(flag, non_blocking=True)
        ev.record()
        self.pending.append((host, ev))
        if len(self.pending) > 4:
            self.check(block=False)

    def check(self, block=True):
        if self.captured:
            if block:
                torch.cuda.synchronize()
            if any(int(h[0]) != 0 for h in self.captured):
                raise RuntimeError("bar_energy_loss: a Gauss point fell outside its own element (degenerate grid); "
                                   "the fused path is invalid here -- use energy_loss_generic")
        keep = []
        for host, ev in self.pending:
            if block:
                ev.synchronize()
            if ev.query():
                bad = int(host[0]) != 0
                self.free.append((host, ev))
                if bad:
                    self.pending = []
                    raise RuntimeError("bar_energy_loss: a Gauss point fell outside its own element (degenerate grid); "
                                       "the fused path is invalid here -- use energy_loss_generic")
            else:
                keep.append((host, ev))
        self.pending = keep


_bar_state = _FlagState()


def bar_energy_loss(model, xi, wi, b_force, E, L=10.0, b_builtin=False):
    """Fused drop-in for `energy_loss(model, xi, wi, b_force, E, L)` of examples/example3.py:27-70.
    `b_builtin=True` evaluates that example's own b_force (example3.py:16-24) inside the kernel instead of
    calling the Python callable on the [Ne, ng] quadrature points."""
    fused_step = b_builtin and getattr(model, "r_adapt", False) and model.N > 2 and torch.is_grad_enabled() \
        and model.x_increments.requires_grad and model.u.requires_grad
    if fused_step:
        # the whole step in three launches (no grid / u_full tensors are materialised)
        p = model.x_increments
        dt = p.dtype
        xi_c, wi_c = xi.to(device=p.device, dtype=dt).contiguous(), wi.to(device=p.device, dtype=dt).contiguous()
        return _BarStepFn.apply(p, model.u, model.x0, model.xN, model.u0_fixed, model.uN_fixed, xi_c, wi_c, E, None, _bar_state)
    grid = model.grid
    dt = grid.dtype
    xi_c, wi_c = xi.to(device=grid.device, dtype=dt).contiguous(), wi.to(device=grid.device, dtype=dt).contiguous()
    b_table = None
    if not b_builtin:
        with torch.no_grad():
            g = grid.detach()
            x_i, x_ip1 = g[:-1].unsqueeze(1), g[1:].unsqueeze(1)
            xq = 0.5 * (x_ip1 - x_i) * xi_c + 0.5 * (x_ip1 + x_i)
            b_table = b_force(xq).to(dt).contiguous()
    return _BarEnergyFn.apply(grid, model.u_full, xi_c, wi_c, E, b_table, _bar_state)


def energy_loss_generic(model, xi, wi, b_force, E, L=10.0):
    """The reference's energy_loss verbatim in structure (examples/example3.py:27-70), running on the generic
    differentiable forward (double backward through autograd.grad)."""
    with torch.no_grad():
        grid = model.grid
        x_i, x_ip1 = grid[:-1].unsqueeze(1), grid[1:].unsqueeze(1)
        xq = 0.5 * (x_ip1 - x_i) * xi + 0.5 * (x_ip1 + x_i)
        wq = 0.5 * (x_ip1 - x_i) * wi
    xq.requires_grad_(True)
    u = model(xq)
    du_dx = torch.autograd.grad(u, xq, grad_outputs=torch.ones_like(u), create_graph=True)[0]
    return torch.sum(wq * (0.5 * E * du_dx ** 2 - b_force(xq) * u))


def example3_b_force(x):
    """examples/example3.py:16-24."""
    N1 = 4 * torch.pi ** 2 * (x - 2.5) ** 2 - 2 * torch.pi
    D1 = torch.exp(torch.pi * (x - 2.5) ** 2)
    N2 = 8 * torch.pi ** 2 * (x - 7.5) ** 2 - 4 * torch.pi
    D2 = torch.exp(torch.pi * (x - 7.5) ** 2)
    return -N1 / D1 - N2 / D2


# ------------------------------------------------------------------------------------------------
# structured Q1 model (models.py:93-212)
# ------------------------------------------------------------------------------------------------
def _q1_backward(gx, gy, uf, xs, ix, iy, r):
    """VJP of the structured interpolation for cotangent r [M]: sort-free deterministic binning (integer histogram ->
    prefix -> scatter -> per-cell ascending sample id) + fused per-cell fold + node / grid-line folds."""
    dt, dev = gx.dtype, gx.device
    M, Nx, Ny = xs.shape[0], gx.shape[0], gy.shape[0]
    s = _lib.stream_ptr()
    L = _lib.lib()
    ncell = (Nx - 1) * (Ny - 1)
    count = torch.zeros(ncell, device=dev, dtype=torch.int32)
    _lib.check(L.hidenn_q1_bin_count(_lib.ptr(ix), _lib.ptr(iy), c_i64(M), c_i64(Ny), _lib.ptr(count), s))
    seg = torch.zeros(ncell + 1, device=dev, dtype=torch.int32)
    torch.cumsum(count, 0, dtype=torch.int32, out=seg[1:])
    count.zero_()
    order = torch.empty(max(M, 1), device=dev, dtype=torch.int32)
    _lib.check(L.hidenn_q1_bin_scatter(_lib.ptr(ix), _lib.ptr(iy), c_i64(M), c_i64(Nx), c_i64(Ny), _lib.ptr(seg),
                                       _lib.ptr(count), _lib.ptr(order), s))
    tmp = torch.empty(ncell, 8, device=dev, dtype=dt)
    du = torch.empty(Nx, Ny, device=dev, dtype=dt)
    dgx = torch.empty(Nx, device=dev, dtype=dt)
    dgy = torch.empty(Ny, device=dev, dtype=dt)
    _lib.check(_lib.fn("hidenn_q1_bwd_fused", dt)(_lib.ptr(gx), c_i64(Nx), _lib.ptr(gy), c_i64(Ny), _lib.ptr(uf), _lib.ptr(xs),
                                                  _lib.ptr(r), c_i64(M), _lib.ptr(seg), _lib.ptr(order),
                                                  _lib.ptr(tmp), _lib.ptr(du), _lib.ptr(dgx), _lib.ptr(dgy), s))
    return dgx, dgy, du


class _Q1L2Fn(torch.autograd.Function):
    """mean((model(x) - target)^2) in one forward pass (lookups, interpolation, residual weights, loss) and the fused
    structured backward on the residual weights."""

    @staticmethod
    @_lib.on_device
    def forward(ctx, gx, gy, u_full, x, target):
        _require_cuda(gx, "l2_projection_loss")
        dt, dev = gx.dtype, gx.device
        gxc, gyc, uf = gx.contiguous(), gy.to(dt).contiguous(), u_full.to(dt).contiguous()
        xs, tg = x.to(dt).contiguous(), target.to(dt).contiguous().reshape(-1)
        M = xs.shape[0]
        if tg.shape[0] != M:
            raise ValueError("l2_projection_loss: target must have one value per sample")
        r = torch.empty(M, device=dev, dtype=dt)
        ix = torch.empty(M, device=dev, dtype=torch.int32)
        iy = torch.empty(M, device=dev, dtype=torch.int32)
        partial = torch.empty(int(_lib.lib().hidenn_q1_l2_partials()), device=dev, dtype=dt)
        loss = torch.empty(1, device=dev, dtype=dt)
        _lib.check(_lib.fn("hidenn_q1_l2_fwd", dt)(_lib.ptr(gxc), c_i64(gxc.shape[0]), _lib.ptr(gyc), c_i64(gyc.shape[0]), _lib.ptr(uf),
                                                   _lib.ptr(xs), _lib.ptr(tg), c_i64(M), _lib.ptr(r), _lib.ptr(ix), _lib.ptr(iy),
                                                   _lib.ptr(partial), _lib.ptr(loss), _lib.stream_ptr()))
        ctx.save_for_backward(gxc, gyc, uf, xs, ix, iy, r)
        ctx.udtype = u_full.dtype
        ctx.used = False
        return loss[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, go):
        if ctx.used:
            raise RuntimeError("l2_projection_loss: backward called twice on the same loss (the residual buffer is scaled in "
                               "place); evaluate the loss again instead")
        ctx.used = True
        gx, gy, uf, xs, ix, iy, r = ctx.saved_tensors
        dt = gx.dtype
        scale = go.reshape(1).to(dt).contiguous()
        _lib.check(_lib.fn("hidenn_scale_inplace", dt)(_lib.ptr(r), c_i64(r.numel()), _lib.ptr(scale), _lib.stream_ptr()))   # exits at once if 1
        with _lib.nvtx("hidenn.q1_backward"):
            dgx, dgy, du = _q1_backward(gx, gy, uf, xs, ix, iy, r)
        return dgx, dgy, du.to(ctx.udtype), None, None


def l2_projection_loss(model, x, u_true):
    """Fused drop-in for `((model(x) - u_true) ** 2).mean()` of examples/example2.py:45-46 on the structured model
    (one forward pass instead of interpolation + three elementwise passes; same deterministic backward)."""
    gx, gy = model.grid
    return _Q1L2Fn.apply(gx, gy, model.u_full, x, u_true)


class _Q1InterpFn(torch.autograd.Function):
    @staticmethod
    @_lib.on_device
    def forward(ctx, gx, gy, u_full, x):
        _require_cuda(gx, "forward")
        dt = gx.dtype
        gxc, gyc, uf, xs = gx.contiguous(), gy.to(dt).contiguous(), u_full.to(dt).contiguous(), x.to(dt).contiguous()
        M = xs.shape[0]
        u = torch.empty(M, device=gxc.device, dtype=dt)
        ix = torch.empty(M, device=gxc.device, dtype=torch.int32)
        iy = torch.empty(M, device=gxc.device, dtype=torch.int32)
        _lib.check(_lib.fn("hidenn_q1_interp_fwd", dt)(_lib.ptr(gxc), c_i64(gxc.shape[0]), _lib.ptr(gyc), c_i64(gyc.shape[0]),
                                                       _lib.ptr(uf), _lib.ptr(xs), c_i64(M), _lib.ptr(u), _lib.ptr(ix), _lib.ptr(iy),
                                                       _lib.stream_ptr()))
        ctx.save_for_backward(gxc, gyc, uf, xs, ix, iy)
        ctx.udtype = u_full.dtype
        return u

    @staticmethod
    @torch.autograd.function.once_differentiable
    @_lib.on_device
    def backward(ctx, r):
        gx, gy, uf, xs, ix, iy = ctx.saved_tensors
        dt, dev = gx.dtype, gx.device
        M = xs.shape[0]
        Nx, Ny = gx.shape[0], gy.shape[0]
        dgx, dgy, du = _q1_backward(gx, gy, uf, xs, ix, iy, r.to(dt).contiguous())
        return dgx, dgy, du.to(ctx.udtype), None


class StructuredShapeNN2D(nn.Module):
    """Tensor-product Q1 interpolant on two r-adaptive 1D grids: the *first* class named
    PiecewiseLinearShapeNN2D in the reference (models.py:93-212), which its second definition shadows."""

    def __init__(self, grid_x, grid_y, boundary_mask_x=None, boundary_mask_y=None, r_adapt=False, u_fixed=None):
        super().__init__()
        self.Nx = grid_x.numel()
        self.Ny = grid_y.numel()
        self.r_adapt = r_adapt
        self.register_buffer("initial_x_grid", grid_x.clone())
        self.register_buffer("initial_y_grid", grid_y.clone())
        self.register_buffer("x0", grid_x.flatten()[0:1])
        self.register_buffer("xN", grid_x.flatten()[-1:])
        self.register_buffer("y0", grid_y.flatten()[0:1])
        self.register_buffer("yN", grid_y.flatten()[-1:])
        if self.r_adapt and max(self.Nx, self.Ny) > 2:
            self.increments_x = nn.Parameter(grid_x[1:] - grid_x[:-1])
            self.increments_y = nn.Parameter(grid_y[1:] - grid_y[:-1])
        else:
            self.register_buffer("x_grid_inner", grid_x[1:-1])
            self.register_buffer("y_grid_inner", grid_y[1:-1])
        if boundary_mask_x is None:
            boundary_mask_x = torch.zeros(self.Nx, dtype=torch.bool)
            boundary_mask_x[0] = boundary_mask_x[-1] = True
        if boundary_mask_y is None:
            boundary_mask_y = torch.zeros(self.Ny, dtype=torch.bool)
            boundary_mask_y[0] = boundary_mask_y[-1] = True
        self.register_buffer("boundary_mask_x", boundary_mask_x)
        self.register_buffer("boundary_mask_y", boundary_mask_y)
        self.register_buffer("node_mask", self.boundary_mask_x[:, None] | self.boundary_mask_y[None, :])
        if u_fixed is not None:
            self.register_buffer("u_fixed", torch.tensor([u_fixed], dtype=torch.float32))
        else:
            self.u_fixed = None
        self.u = nn.Parameter(torch.randn(self.Nx, self.Ny))
        self.epsilon = 1e-10

    @property
    def grid(self):
        if self.r_adapt and max(self.Nx, self.Ny) > 2:
            x_full = _GridFn.apply(self.increments_x, self.x0, self.xN)
            y_full = _GridFn.apply(self.increments_y, self.y0, self.yN)
        else:
            x_full = torch.cat([self.x0, self.x_grid_inner, self.xN], dim=0)
            y_full = torch.cat([self.y0, self.y_grid_inner, self.yN], dim=0)
        x_full = torch.where(self.boundary_mask_x, self.initial_x_grid, x_full)
        y_full = torch.where(self.boundary_mask_y, self.initial_y_grid, y_full)
        return x_full, y_full

    @property
    def u_full(self):
        if self.u_fixed is not None:
            return torch.where(self.node_mask, self.u_fixed, self.u)
        return self.u

    def forward(self, x_eval):
        gx, gy = self.grid
        return _Q1InterpFn.apply(gx, gy, self.u_full, x_eval)
