"""ORACLE (test infrastructure, NOT product code) -- torch-autograd CPU port.

A functional restatement of the reference's *computation graph* for the hot path, using
the same ATen operations in the same order (masked index_put assembly, advanced-index
gathers, batched linalg.det / linalg.inv per Gauss point, einsum, torch.sum, autograd
backward).  It exists for two reasons:

  * it is what bench.py times as the CPU baseline / `--impl reference` arm on the GPU box,
    where /root/reference does not exist (cpu_baseline.kind = "port");
  * it is a second, independently written oracle next to oracle/closed_form.py.

Pinned against the unmodified reference by tests/golden (see tests/test_oracle_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _assemble(free_vals, fixed_vals, free_mask, fixed_mask, n):
    # models.py:292-305: zeros, then two masked index_put
    full = torch.zeros(n, 2, device=free_vals.device, dtype=free_vals.dtype)
    full[free_mask] = free_vals
    if fixed_vals is not None:
        full[fixed_mask] = fixed_vals
    return full


class TriPort:
    """State of the triangle model (models.py:241-282) held as plain tensors."""

    def __init__(self, node_coords, connectivity, boundary_mask, dirichlet_mask, u_fixed, neumann_edges,
                 u_free=None):
        self.Nn = node_coords.shape[0]
        self.conn = connectivity.long()
        self.bmask = boundary_mask.clone()
        self.fmask = ~boundary_mask
        self.dmask = dirichlet_mask.clone()
        self.umask = ~dirichlet_mask
        self.x_free = node_coords[self.fmask].clone().requires_grad_(True)
        self.x_fixed = node_coords[self.bmask].clone()
        if u_free is None:
            u_free = 1e-5 * torch.randn(int(self.umask.sum()), 2)
        self.u_free = u_free.to(node_coords.dtype).clone().requires_grad_(True)
        self.u_fixed = None if u_fixed is None else torch.as_tensor(u_fixed, dtype=node_coords.dtype)
        self.edges = neumann_edges.long()

    def coords(self):
        return _assemble(self.x_free, self.x_fixed, self.fmask, self.bmask, self.Nn)

    def u_full(self):
        return _assemble(self.u_free, self.u_fixed, self.umask, self.dmask, self.Nn)


def tri_points(m: TriPort, x_ref, elem_id, jinv_transpose=False):
    """models.py:316-357.  jinv_transpose=True: the correct-math variant dN/dx = J^-T dN/dxi (the package's default-off
    switch), NOT what the reference computes."""
    nodes = m.conn[elem_id]
    v = m.coords()[nodes]
    xi, eta = x_ref[:, 0:1], x_ref[:, 1:2]
    N = torch.cat([xi, eta, 1.0 - xi - eta], dim=1)
    un = m.u_full()[nodes]
    u_h = torch.sum(N.unsqueeze(2) * un, dim=1)
    Jm = torch.stack([v[:, 0, :] - v[:, 2, :], v[:, 1, :] - v[:, 2, :]], dim=2)
    det = torch.linalg.det(Jm)
    Jinv = torch.linalg.inv(Jm)
    R = torch.tensor([[1., 0., -1.], [0., 1., -1.]], dtype=v.dtype)
    dN = torch.einsum("mji,jk->mik", Jinv, R) if jinv_transpose else torch.einsum("mij,jk->mik", Jinv, R)
    G = torch.einsum("mai,mja->mij", un, dN)
    return u_h, det, G


def tri_edge_points(m: TriPort, xi, edge_id):
    """models.py:359-376."""
    co = m.coords()
    e = m.edges[edge_id]
    x0, x1 = co[e[:, 0]], co[e[:, 1]]
    N = torch.cat([1.0 - xi[:, 0:1], xi[:, 0:1]], dim=1)
    un = m.u_full()[e]
    return torch.sum(N.unsqueeze(2) * un, dim=1), torch.norm(x1 - x0, dim=1)


def tri_energy(m: TriPort, C, xg, wg, xi1, w1, b_force=None, t_force=None, jinv_transpose=False):
    """loss.py:55-116."""
    Ne, ng = m.conn.shape[0], xg.shape[0]
    x_eval = xg.unsqueeze(0).expand(Ne, ng, 2).reshape(-1, 2)
    elem_id = torch.arange(Ne).unsqueeze(1).repeat(1, ng).reshape(-1)
    wflat = wg.unsqueeze(0).repeat(Ne, 1).reshape(-1)
    u_h, det, G = tri_points(m, x_eval, elem_id, jinv_transpose)
    eps = torch.stack([G[:, 0, 0], G[:, 1, 1], 2 * (0.5 * (G[:, 0, 1] + G[:, 1, 0]))], dim=1)
    sig = eps @ C.T
    psi = 0.5 * torch.sum(eps * sig, dim=1)
    b = b_force(x_eval) if b_force is not None else torch.zeros_like(x_eval)
    qw = wflat * det.abs()
    dom = torch.sum(qw * psi) - torch.sum(qw * torch.sum(b * u_h, dim=1))
    # edge term
    Ned, ng1 = m.edges.shape[0], xi1.shape[0]
    co = m.coords()
    x0, x1 = co[m.edges[:, 0]], co[m.edges[:, 1]]
    xq = (1.0 - xi1[None, :, None]) * x0[:, None, :] + xi1[None, :, None] * x1[:, None, :]
    xq = xq.reshape(-1, 2)
    wq = w1[None, :].expand(Ned, ng1).reshape(-1)
    xe = xi1[None, :].expand(Ned, ng1).reshape(-1, 1)
    eid = torch.repeat_interleave(torch.arange(Ned), repeats=ng1)
    ue, ds = tri_edge_points(m, xe, eid)
    if t_force is not None:
        t = t_force(xq)
    else:
        t = torch.stack([torch.full((xq.shape[0],), 100e3 / 1.0, dtype=xq.dtype),
                         torch.zeros(xq.shape[0], dtype=xq.dtype)], dim=1)
    edge = torch.sum((ue * t).sum(dim=1) * (wq * ds))
    return dom - edge


# ------------------------------------------------------------------ 1D (models.py:6-90)

def grid_1d(p, x0, xN):
    inc = torch.clamp(F.softplus(p), min=1e-6)
    cum = torch.cumsum(inc, dim=0)
    return torch.cat([x0, x0 + (xN - x0) * cum / cum[-1]], dim=0)


def interp_1d(grid, u_full, x):
    N = grid.shape[0]
    e = (torch.searchsorted(grid, x) - 1).clamp(0, N - 2)
    xi_, xip = grid[e], grid[e + 1]
    N1 = (xip - x) / (xip - xi_).clamp(1e-10)
    N2 = (x - xi_) / (xip - xi_).clamp(1e-10)
    return u_full[e] * N1 + u_full[e + 1] * N2


def example3_b_force(x):
    N1 = 4 * torch.pi ** 2 * (x - 2.5) ** 2 - 2 * torch.pi
    D1 = torch.exp(torch.pi * (x - 2.5) ** 2)
    N2 = 8 * torch.pi ** 2 * (x - 7.5) ** 2 - 4 * torch.pi
    D2 = torch.exp(torch.pi * (x - 7.5) ** 2)
    return -N1 / D1 - N2 / D2


def bar_energy(grid_fn, u_full_fn, xi, wi, E, b_force=example3_b_force):
    """examples/example3.py:27-70 (double-backward through autograd.grad)."""
    with torch.no_grad():
        g = grid_fn()
        gi, gp = g[:-1].unsqueeze(1), g[1:].unsqueeze(1)
        xq = 0.5 * (gp - gi) * xi + 0.5 * (gp + gi)
        wq = 0.5 * (gp - gi) * wi
    xq.requires_grad_(True)
    u = interp_1d(grid_fn(), u_full_fn(), xq)
    du = torch.autograd.grad(u, xq, grad_outputs=torch.ones_like(u), create_graph=True)[0]
    return torch.sum(wq * (0.5 * E * du ** 2 - b_force(xq) * u))


# ------------------------------------------------------------------ structured Q1 (models.py:93-212)

def q1_interp(gx, gy, u_full, x):
    Nx, Ny = gx.shape[0], gy.shape[0]
    ix = (torch.searchsorted(gx, x[:, 0].contiguous()) - 1).clamp(0, Nx - 2)
    iy = (torch.searchsorted(gy, x[:, 1].contiguous()) - 1).clamp(0, Ny - 2)
    xi_, xip, yi_, yip = gx[ix], gx[ix + 1], gy[iy], gy[iy + 1]
    N1x = (xip - x[:, 0]) / (xip - xi_).clamp(1e-10)
    N2x = (x[:, 0] - xi_) / (xip - xi_).clamp(1e-10)
    N1y = (yip - x[:, 1]) / (yip - yi_).clamp(1e-10)
    N2y = (x[:, 1] - yi_) / (yip - yi_).clamp(1e-10)
    return (N1x * N1y * u_full[ix, iy] + N2x * N1y * u_full[ix + 1, iy]
            + N1x * N2y * u_full[ix, iy + 1] + N2x * N2y * u_full[ix + 1, iy + 1])
