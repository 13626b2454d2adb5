"""ORACLE (test infrastructure, NOT product code) -- closed-form CPU restatement.

numpy restatement of the reference's quadrature hot path, forward AND backward,
written from the formulas of SURVEY.md Appendix A.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package.

Pinned by tests/test_oracle_golden.py against fixtures in tests/golden/*.npz that were
produced by running the UNMODIFIED reference (/root/reference/src/models.py, loss.py,
utils.py, examples/example3.py) in the build container (tests/golden/make_golden.py).

Every function cites the reference lines it restates.  The reference's quirks
 are kept on purpose: J^-1 (not J^-T), order-4/6 weights summing
to 0.25, raw [-1,1] Gauss points on edges, body force at reference coordinates, |det J|.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# quadrature tables  (/root/reference/src/utils.py:4-81)
# --------------------------------------------------------------------------------------

def interval_gauss_points(order=1, dtype=np.float64, unit_interval=False):
    """utils.py:4-11 -- raw Gauss-Legendre on [-1,1] (Q3: *not* mapped to [0,1]).
    unit_interval=True: the correct-math variant (points and weights mapped to [0,1]), not the reference's."""
    xi, wi = np.polynomial.legendre.leggauss(order)
    if unit_interval:
        xi, wi = 0.5 * (xi + 1.0), 0.5 * wi
    return xi.astype(dtype), wi.astype(dtype)


def triangle_gauss_points(order=1, dtype=np.float64, fix_weights=False):
    """utils.py:13-81 -- orders 1,3,4,6,7; orders 4 and 6 carry the extra 0.5 (Q2).
    fix_weights=True: the correct-math variant (orders 4 and 6 sum to the triangle area 0.5), not the reference's."""
    if order == 1:
        rs = [[1 / 3, 1 / 3]]
        w = [0.5]
    elif order == 3:
        a = 1 / 6
        rs = [[a, a], [4 * a, a], [a, 4 * a]]
        w = [1 / 6, 1 / 6, 1 / 6]
    elif order == 4:
        rs = [[1 / 3, 1 / 3], [0.6, 0.2], [0.2, 0.6], [0.2, 0.2]]
        w = [-27 / 96, 25 / 96, 25 / 96, 25 / 96]
    elif order == 6:
        a, b = 0.445948490915965, 0.091576213509771
        w1, w2 = 0.111690794839005, 0.054975871827661
        rs = [[a, a], [1 - 2 * a, a], [a, 1 - 2 * a], [b, b], [1 - 2 * b, b], [b, 1 - 2 * b]]
        w = [w1, w1, w1, w2, w2, w2]
    elif order == 7:
        rs = [[1 / 3, 1 / 3], [0.0597158717, 0.4701420641], [0.4701420641, 0.0597158717],
              [0.4701420641, 0.4701420641], [0.7974269853, 0.1012865073],
              [0.1012865073, 0.7974269853], [0.1012865073, 0.1012865073]]
        w = [0.225, 0.1323941527, 0.1323941527, 0.1323941527, 0.1259391805, 0.1259391805, 0.1259391805]
    else:
        raise NotImplementedError("Supported orders: 1, 3, 4, 6, 7")
    # the reference builds the table in `dtype` and multiplies by python 0.5 in that dtype
    rs = np.asarray(rs, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    if order in (4, 6, 7) and not (fix_weights and order in (4, 6)):
        w = (np.asarray(0.5, dtype=dtype) * w).astype(dtype)
    return rs, w


def plane_stress_C(E=10e9, nu=0.3, dtype=np.float64):
    """loss.py:29-32."""
    factor = E / (1 - nu ** 2)
    base = np.asarray([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, (1.0 - nu) / 2.0]], dtype=dtype)
    return (base * np.asarray(factor, dtype=dtype)).astype(dtype)


# --------------------------------------------------------------------------------------
# triangle model: full arrays from parameters  (models.py:292-305)
# --------------------------------------------------------------------------------------

def assemble_full(free_vals, fixed_vals, free_mask):
    """coords / u_full: rows of free_mask <- free_vals (ascending node order), others <- fixed."""
    n = free_mask.shape[0]
    out = np.zeros((n, 2), dtype=free_vals.dtype)
    out[free_mask] = free_vals
    fm = ~free_mask
    if fixed_vals is not None:
        out[fm] = fixed_vals
    return out


def tri_forward_points(coords, U, conn, x_ref, elem_id, jinv_transpose=False):
    """models.py:316-357 -- (u_h, detJ signed, grad_u) at reference points of given elements.
    jinv_transpose=True: correct-math variant grad_u = dU . J^-1 (i.e. dN/dx = J^-T dN/dxi), not the reference's."""
    n = conn[elem_id]
    v = coords[n]                       # [M,3,2]
    u = U[n]
    xi, eta = x_ref[:, 0], x_ref[:, 1]
    N = np.stack([xi, eta, 1.0 - xi - eta], axis=1)
    u_h = (N[:, :, None] * u).sum(axis=1)
    a = v[:, 0, 0] - v[:, 2, 0]
    b = v[:, 1, 0] - v[:, 2, 0]
    c = v[:, 0, 1] - v[:, 2, 1]
    d = v[:, 1, 1] - v[:, 2, 1]
    det = a * d - b * c
    inv = 1.0 / det
    J00, J01, J10, J11 = d * inv, -b * inv, -c * inv, a * inv
    du0 = u[:, 0] - u[:, 2]
    du1 = u[:, 1] - u[:, 2]
    G = np.empty((n.shape[0], 2, 2), dtype=coords.dtype)
    if jinv_transpose:
        G[:, :, 0] = du0 * J00[:, None] + du1 * J10[:, None]
        G[:, :, 1] = du0 * J01[:, None] + du1 * J11[:, None]
    else:
        G[:, :, 0] = du0 * J00[:, None] + du1 * J01[:, None]
        G[:, :, 1] = du0 * J10[:, None] + du1 * J11[:, None]
    return u_h, det, G


def tri_edge_forward(coords, U, edges, xi, edge_id):
    """models.py:359-376 -- (u_h, ds) on Neumann edges, N=[1-xi, xi]."""
    e = edges[edge_id]
    x0, x1 = coords[e[:, 0]], coords[e[:, 1]]
    u0, u1 = U[e[:, 0]], U[e[:, 1]]
    u_h = (1.0 - xi)[:, None] * u0 + xi[:, None] * u1
    d = x1 - x0
    return u_h, np.sqrt((d * d).sum(axis=1))


def _fold(idx, vals, n):
    out = np.empty((n, 2), dtype=vals.dtype)
    out[:, 0] = np.bincount(idx, weights=vals[:, 0], minlength=n)
    out[:, 1] = np.bincount(idx, weights=vals[:, 1], minlength=n)
    return out


def tri_energy_full(coords, U, conn, C, xg, wg, bg=None, edges=None, xi1=None, w1=None,
                    t_q=None, dt_dx=None, want_grad=True, jinv_transpose=False):
    """Total potential  E_dom - E_edge  and its gradient w.r.t. the FULL coords / U arrays.

    Restates loss.py:55-116 over models.py:316-376 with SURVEY Appendix A.1.
      bg    [ng,2]    b_force evaluated at the *reference* Gauss points (loss.py:80, Q4); None = 0
      t_q   [Ned,ng1,2] traction at the physical edge points (loss.py:106); None = (1e5,0) (loss.py:47-51)
      dt_dx [Ned,ng1,2,2] optional d t_i / d x_j at those points (only if t_force depends on x)
    Accumulation is float64 regardless of input dtype only where numpy's bincount forces it;
    callers compare with tolerances stated in the tests.
    """
    dt = coords.dtype
    n0, n1, n2 = conn[:, 0], conn[:, 1], conn[:, 2]
    v0, v1, v2 = coords[n0], coords[n1], coords[n2]
    U0, U1, U2 = U[n0], U[n1], U[n2]
    a = v0[:, 0] - v2[:, 0]
    b = v1[:, 0] - v2[:, 0]
    c = v0[:, 1] - v2[:, 1]
    d = v1[:, 1] - v2[:, 1]
    det = a * d - b * c
    A = np.abs(det)
    s = np.sign(det)
    inv = 1.0 / det
    J00, J01, J10, J11 = d * inv, -b * inv, -c * inv, a * inv       # Jinv
    du0, du1 = U0 - U2, U1 - U2                                       # [Ne,2] (component i)
    if jinv_transpose:
        # correct-math variant (not the reference's): G = dU . Jinv, i.e. G[i][j] = sum_m dU[i][m] Jinv[m][j]
        G00 = du0[:, 0] * J00 + du1[:, 0] * J10
        G01 = du0[:, 0] * J01 + du1[:, 0] * J11
        G10 = du0[:, 1] * J00 + du1[:, 1] * J10
        G11 = du0[:, 1] * J01 + du1[:, 1] * J11
    else:
        # G[i][j] = sum_m dU[i][m] Jinv[j][m]
        G00 = du0[:, 0] * J00 + du1[:, 0] * J01
        G01 = du0[:, 0] * J10 + du1[:, 0] * J11
        G10 = du0[:, 1] * J00 + du1[:, 1] * J01
        G11 = du0[:, 1] * J10 + du1[:, 1] * J11
    e0, e1, e2 = G00, G11, G01 + G10
    s0 = C[0, 0] * e0 + C[0, 1] * e1 + C[0, 2] * e2
    s1 = C[1, 0] * e0 + C[1, 1] * e1 + C[1, 2] * e2
    s2 = C[2, 0] * e0 + C[2, 1] * e1 + C[2, 2] * e2
    psi = 0.5 * (e0 * s0 + e1 * s1 + e2 * s2)
    W = wg.sum(dtype=dt)
    Ng = np.stack([xg[:, 0], xg[:, 1], 1.0 - xg[:, 0] - xg[:, 1]], axis=1)      # [ng,3]
    if bg is None:
        Fb = np.zeros((3, 2), dtype=dt)
    else:
        Fb = np.einsum("g,gk,gi->ki", wg, Ng, bg).astype(dt)
    bw = (U0 * Fb[0]).sum(1) + (U1 * Fb[1]).sum(1) + (U2 * Fb[2]).sum(1)
    dens = W * psi - bw
    E_dom = (A * dens).sum()

    Nn = coords.shape[0]
    dU = dX = None
    if want_grad:
        # when C is not symmetric d psi/d eps = 0.5 (C + C^T) eps; the reference's C is symmetric
        Cs = 0.5 * (C + C.T)
        q0 = Cs[0, 0] * e0 + Cs[0, 1] * e1 + Cs[0, 2] * e2
        q1 = Cs[1, 0] * e0 + Cs[1, 1] * e1 + Cs[1, 2] * e2
        q2 = Cs[2, 0] * e0 + Cs[2, 1] * e1 + Cs[2, 2] * e2
        # P = d psi / d G = [[q0,q2],[q2,q1]] ;  M = d psi / d dU = P . Jinv  (P . Jinv^T in the correct-math variant)
        if jinv_transpose:
            M00 = q0 * J00 + q2 * J01
            M01 = q0 * J10 + q2 * J11
            M10 = q2 * J00 + q1 * J01
            M11 = q2 * J10 + q1 * J11
        else:
            M00 = q0 * J00 + q2 * J10
            M01 = q0 * J01 + q2 * J11
            M10 = q2 * J00 + q1 * J10
            M11 = q2 * J01 + q1 * J11
        AW = A * W
        gU0 = np.stack([AW * M00, AW * M10], 1) - A[:, None] * Fb[0]
        gU1 = np.stack([AW * M01, AW * M11], 1) - A[:, None] * Fb[1]
        gU2 = -np.stack([AW * (M00 + M01), AW * (M10 + M11)], 1) - A[:, None] * Fb[2]
        if jinv_transpose:
            # d psi / d J = -G^T M   (from dG = -G dJ Jinv)
            K00 = -(G00 * M00 + G10 * M10)
            K01 = -(G00 * M01 + G10 * M11)
            K10 = -(G01 * M00 + G11 * M10)
            K11 = -(G01 * M01 + G11 * M11)
        else:
            # d psi / d J = -M^T G
            K00 = -(M00 * G00 + M10 * G10)
            K01 = -(M00 * G01 + M10 * G11)
            K10 = -(M01 * G00 + M11 * G10)
            K11 = -(M01 * G01 + M11 * G11)
        sd = s * dens
        # d|det|/dJ = s * [[d,-c],[-b,a]]
        D00 = sd * d + AW * K00
        D01 = -sd * c + AW * K01
        D10 = -sd * b + AW * K10
        D11 = sd * a + AW * K11
        gX0 = np.stack([D00, D10], 1)       # column 0 of dE/dJ -> v0
        gX1 = np.stack([D01, D11], 1)       # column 1 -> v1
        gX2 = -(gX0 + gX1)
        idx = np.concatenate([n0, n1, n2])
        dU = _fold(idx, np.concatenate([gU0, gU1, gU2]), Nn).astype(dt)
        dX = _fold(idx, np.concatenate([gX0, gX1, gX2]), Nn).astype(dt)

    E_edge = np.asarray(0.0, dtype=dt)
    if edges is not None and edges.shape[0] > 0:
        i0, i1 = edges[:, 0], edges[:, 1]
        x0, x1 = coords[i0], coords[i1]
        u0, u1 = U[i0], U[i1]
        dvec = x1 - x0
        ds = np.sqrt((dvec * dvec).sum(1))
        xi = xi1[None, :, None]
        uq = (1.0 - xi) * u0[:, None, :] + xi * u1[:, None, :]           # [Ned,ng1,2]
        if t_q is None:
            t_q = np.zeros(uq.shape, dtype=dt)
            t_q[..., 0] = 100e3 / 1.0
        ut = (uq * t_q).sum(2)                                            # [Ned,ng1]
        S = (ut * w1[None, :]).sum(1)                                     # [Ned]
        E_edge = (S * ds).sum()
        if want_grad:
            wt = t_q * w1[None, :, None]
            g0 = -(ds[:, None]) * (wt * (1.0 - xi)).sum(1)
            g1 = -(ds[:, None]) * (wt * xi).sum(1)
            dU += _fold(np.concatenate([i0, i1]), np.concatenate([g0, g1]), Nn).astype(dt)
            dirn = dvec / ds[:, None]
            gx0 = S[:, None] * dirn
            gx1 = -S[:, None] * dirn
            if dt_dx is not None:
                # g_q[j] = ds w_q sum_i u_q[i] dt_i/dx_j
                gq = ds[:, None, None] * w1[None, :, None] * np.einsum("eqi,eqij->eqj", uq, dt_dx)
                gx0 = gx0 - ((1.0 - xi) * gq).sum(1)
                gx1 = gx1 - (xi * gq).sum(1)
            dX += _fold(np.concatenate([i0, i1]), np.concatenate([gx0, gx1]), Nn).astype(dt)
    loss = E_dom - E_edge
    return loss, dX, dU


# --------------------------------------------------------------------------------------
# 1D model  (models.py:6-90) and bar energy (examples/example3.py:27-70)
# --------------------------------------------------------------------------------------

def softplus(p):
    """torch.nn.functional.softplus, beta=1, threshold=20."""
    return np.where(p > 20.0, p, np.log1p(np.exp(np.minimum(p, 20.0))))


def grid_1d(p, x0, xN):
    """models.py:45-53 (r-adaptive branch).  Returns grid [N] and the pieces the chain rule needs."""
    sp = softplus(p)
    inc = np.maximum(sp, np.asarray(1e-6, dtype=p.dtype))
    cum = np.cumsum(inc, dtype=p.dtype)
    S = cum[-1]
    L = xN - x0
    inner = x0 + L * cum / S
    g = np.concatenate([np.atleast_1d(x0).astype(p.dtype), inner])
    return g, (sp, cum, S, L)


def grid_1d_backward(dg, p, aux):
    """SURVEY A.2 chain:  dL/dg[1:] -> dL/dp."""
    sp, cum, S, L = aux
    gam = dg[1:]
    dcum = L * gam / S
    dcum = dcum.copy()
    dcum[-1] -= L * (gam * cum).sum() / (S * S)
    dinc = np.cumsum(dcum[::-1], dtype=p.dtype)[::-1]
    sig = 1.0 / (1.0 + np.exp(-p))
    dsp = np.where(p > 20.0, 1.0, sig)
    return dinc * (sp >= 1e-6) * dsp


def lookup_1d(grid, x, N):
    """models.py:73-74: clamp(searchsorted_left(grid,x)-1, 0, N-2)  (Q14) -- bit-exact integer work."""
    return np.clip(np.searchsorted(grid, x, side="left") - 1, 0, N - 2)


def interp_1d(grid, u_full, x):
    """models.py:70-90."""
    e = lookup_1d(grid, x, grid.shape[0])
    xi_, xip = grid[e], grid[e + 1]
    h = np.maximum(xip - xi_, np.asarray(1e-10, dtype=grid.dtype))
    N1 = (xip - x) / h
    N2 = (x - xi_) / h
    return u_full[e] * N1 + u_full[e + 1] * N2, e


def interp_1d_backward(grid, u_full, x, r):
    """VJP of interp_1d for upstream r: returns (d grid, d u_full, d x)."""
    N = grid.shape[0]
    e = lookup_1d(grid, x, N)
    ge, gp = grid[e], grid[e + 1]
    ue, up = u_full[e], u_full[e + 1]
    hraw = gp - ge
    h = np.maximum(hraw, np.asarray(1e-10, dtype=grid.dtype))
    act = (hraw >= 1e-10).astype(grid.dtype)       # clamp passes gradient only where not clamped
    N1 = (gp - x) / h
    N2 = (x - ge) / h
    num = ue * (gp - x) + up * (x - ge)
    du = np.bincount(e.ravel(), weights=(r * N1).ravel(), minlength=N) + \
        np.bincount((e + 1).ravel(), weights=(r * N2).ravel(), minlength=N)
    # d/dg_e: direct (-u_{e+1}/h) and through h (num/h^2 * act);  d/dg_{e+1}: u_e/h - num/h^2*act
    dge = r * (-up / h + act * num / (h * h))
    dgp = r * (ue / h - act * num / (h * h))
    dg = np.bincount(e.ravel(), weights=dge.ravel(), minlength=N) + \
        np.bincount((e + 1).ravel(), weights=dgp.ravel(), minlength=N)
    dx = r * (up - ue) / h
    return dg.astype(grid.dtype), du.astype(grid.dtype), dx


def example3_b_force(x):
    """examples/example3.py:16-24."""
    N1 = 4 * np.pi ** 2 * (x - 2.5) ** 2 - 2 * np.pi
    D1 = np.exp(np.pi * (x - 2.5) ** 2)
    N2 = 8 * np.pi ** 2 * (x - 7.5) ** 2 - 4 * np.pi
    D2 = np.exp(np.pi * (x - 7.5) ** 2)
    return -N1 / D1 - N2 / D2


def bar_energy(grid, u_full, xi, wi, E, b_force=example3_b_force, want_grad=True):
    """examples/example3.py:27-70: sum wq (0.5 E u'^2 - b(xq) u); xq,wq detached from the grid (Q15).

    Returns loss, d grid, d u_full (full arrays)."""
    N = grid.shape[0]
    gi, gp1 = grid[:-1, None], grid[1:, None]
    xq = 0.5 * (gp1 - gi) * xi[None, :] + 0.5 * (gp1 + gi)
    wq = 0.5 * (gp1 - gi) * wi[None, :]
    e = lookup_1d(grid, xq, N)
    ge, gp = grid[e], grid[e + 1]
    ue, up = u_full[e], u_full[e + 1]
    hraw = gp - ge
    h = np.maximum(hraw, np.asarray(1e-10, dtype=grid.dtype))
    act = (hraw >= 1e-10).astype(grid.dtype)
    N1 = (gp - xq) / h
    N2 = (xq - ge) / h
    u = ue * N1 + up * N2
    du = (up - ue) / h
    b = b_force(xq)
    loss = (wq * (0.5 * E * du * du - b * u)).sum()
    if not want_grad:
        return loss, None, None
    r_u = -wq * b
    r_s = wq * E * du                       # upstream on the slope du
    num = ue * (gp - xq) + up * (xq - ge)
    due = r_u * N1 - r_s / h
    dup = r_u * N2 + r_s / h
    # slope = (up-ue)/h -> d slope/d g_e = +act*(up-ue)/h^2 ; d/d g_{e+1} = -act*(up-ue)/h^2
    dge = r_u * (-up / h + act * num / (h * h)) + r_s * act * (up - ue) / (h * h)
    dgp = r_u * (ue / h - act * num / (h * h)) - r_s * act * (up - ue) / (h * h)
    dU = np.bincount(e.ravel(), weights=due.ravel(), minlength=N) + \
        np.bincount((e + 1).ravel(), weights=dup.ravel(), minlength=N)
    dG = np.bincount(e.ravel(), weights=dge.ravel(), minlength=N) + \
        np.bincount((e + 1).ravel(), weights=dgp.ravel(), minlength=N)
    return loss, dG.astype(grid.dtype), dU.astype(grid.dtype)


# --------------------------------------------------------------------------------------
# structured Q1 model (models.py:93-212) and the L2 loss (examples/example2.py:45-46)
# --------------------------------------------------------------------------------------

def q1_interp(gx, gy, u_full, x):
    """models.py:180-212."""
    Nx, Ny = gx.shape[0], gy.shape[0]
    ix = lookup_1d(gx, x[:, 0], Nx)
    iy = lookup_1d(gy, x[:, 1], Ny)
    eps = np.asarray(1e-10, dtype=gx.dtype)
    hx = np.maximum(gx[ix + 1] - gx[ix], eps)
    hy = np.maximum(gy[iy + 1] - gy[iy], eps)
    N1x = (gx[ix + 1] - x[:, 0]) / hx
    N2x = (x[:, 0] - gx[ix]) / hx
    N1y = (gy[iy + 1] - x[:, 1]) / hy
    N2y = (x[:, 1] - gy[iy]) / hy
    u00, u10 = u_full[ix, iy], u_full[ix + 1, iy]
    u01, u11 = u_full[ix, iy + 1], u_full[ix + 1, iy + 1]
    uh = N1x * N1y * u00 + N2x * N1y * u10 + N1x * N2y * u01 + N2x * N2y * u11
    return uh, ix, iy


def q1_interp_backward(gx, gy, u_full, x, r):
    """VJP of q1_interp: (d gx, d gy, d u_full)."""
    Nx, Ny = gx.shape[0], gy.shape[0]
    ix = lookup_1d(gx, x[:, 0], Nx)
    iy = lookup_1d(gy, x[:, 1], Ny)
    eps = np.asarray(1e-10, dtype=gx.dtype)
    hxr = gx[ix + 1] - gx[ix]
    hyr = gy[iy + 1] - gy[iy]
    hx, hy = np.maximum(hxr, eps), np.maximum(hyr, eps)
    ax, ay = (hxr >= 1e-10).astype(gx.dtype), (hyr >= 1e-10).astype(gx.dtype)
    N1x = (gx[ix + 1] - x[:, 0]) / hx
    N2x = (x[:, 0] - gx[ix]) / hx
    N1y = (gy[iy + 1] - x[:, 1]) / hy
    N2y = (x[:, 1] - gy[iy]) / hy
    u00, u10 = u_full[ix, iy], u_full[ix + 1, iy]
    u01, u11 = u_full[ix, iy + 1], u_full[ix + 1, iy + 1]
    dU = np.zeros(Nx * Ny, dtype=np.float64)
    for (di, dj, wgt) in ((0, 0, N1x * N1y), (1, 0, N2x * N1y), (0, 1, N1x * N2y), (1, 1, N2x * N2y)):
        dU += np.bincount((ix + di) * Ny + (iy + dj), weights=r * wgt, minlength=Nx * Ny)
    # x direction: u = A N1x + B N2x
    Aa = N1y * u00 + N2y * u01
    Bb = N1y * u10 + N2y * u11
    numx = Aa * (gx[ix + 1] - x[:, 0]) + Bb * (x[:, 0] - gx[ix])
    dgx = np.bincount(ix, weights=r * (-Bb / hx + ax * numx / (hx * hx)), minlength=Nx) + \
        np.bincount(ix + 1, weights=r * (Aa / hx - ax * numx / (hx * hx)), minlength=Nx)
    Cc = N1x * u00 + N2x * u10
    Dd = N1x * u01 + N2x * u11
    numy = Cc * (gy[iy + 1] - x[:, 1]) + Dd * (x[:, 1] - gy[iy])
    dgy = np.bincount(iy, weights=r * (-Dd / hy + ay * numy / (hy * hy)), minlength=Ny) + \
        np.bincount(iy + 1, weights=r * (Cc / hy - ay * numy / (hy * hy)), minlength=Ny)
    return dgx.astype(gx.dtype), dgy.astype(gx.dtype), dU.reshape(Nx, Ny).astype(gx.dtype)
