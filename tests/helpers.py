"""Shared test helpers: golden loading and oracle evaluation on the same inputs."""
import os

import numpy as np

from oracle import closed_form as cf
import forces

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRI_CASES = ["tri_f64_jitter", "tri_f64_inverted", "tri_f32_jitter", "tri_f64_default", "tri_f64_order3"]


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def relmax(a, b):
    """max-norm relative error  ||a-b||_inf / ||b||_inf  (SURVEY §8(c) parity metric)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def tri_oracle(g, tag, dtype=None, want_grad=True):
    """Closed-form oracle on a golden triangle case -> loss, gx_free, gu_free (parameter layout)."""
    dt = g["node_coords_free"].dtype if dtype is None else dtype
    fmask = ~g["boundary_mask"]
    umask = ~g["dirichlet_mask"]
    coords = cf.assemble_full(g["node_coords_free"].astype(dt), g["node_coords_fixed"].astype(dt), fmask)
    ufix = np.broadcast_to(np.asarray(g["u_fixed"], dtype=dt), (int((~umask).sum()), 2))
    U = cf.assemble_full(g["u_free"].astype(dt), ufix, umask)
    xg, wg = cf.triangle_gauss_points(int(g["gauss_order"]), dt)
    xi1, w1 = cf.interval_gauss_points(int(g["gauss_order_1d"]), dt)
    C = cf.plane_stress_C(10e9, 0.3, dt)
    edges = g["neumann_edges"]
    bg = tq = dtdx = None
    if tag == "forces":
        bg = forces.b_force_np(xg).astype(dt)
        x0, x1 = coords[edges[:, 0]], coords[edges[:, 1]]
        xq = (1.0 - xi1[None, :, None]) * x0[:, None, :] + xi1[None, :, None] * x1[:, None, :]
        tq = forces.t_force_np(xq).astype(dt)
        dtdx = forces.dt_dx_np(xq).astype(dt)
    loss, dX, dU = cf.tri_energy_full(coords, U, g["connectivity"], C, xg, wg, bg, edges, xi1, w1, tq, dtdx,
                                      want_grad=want_grad)
    if not want_grad:
        return loss, None, None
    return loss, dX[fmask], dU[umask]
