"""CPU: pin oracle/ (closed form + torch port) to the fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import torch_port as tp
from helpers import gold, relmax, tri_oracle, TRI_CASES
import forces


def tol_of(dt):
    return 1e-12 if dt == np.float64 else 2e-5


def test_quadrature_tables_bit_exact():
    g = gold("quadrature")
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        for o in (1, 3, 4, 6, 7):
            rs, w = cf.triangle_gauss_points(o, dt)
            assert np.array_equal(rs, g[f"tri{o}_rs_{tag}"]) and np.array_equal(w, g[f"tri{o}_w_{tag}"]), (o, tag)
        for o in (1, 2, 3, 4, 5):
            x, w = cf.interval_gauss_points(o, dt)
            assert np.array_equal(x, g[f"int{o}_x_{tag}"]) and np.array_equal(w, g[f"int{o}_w_{tag}"])
        assert relmax(cf.plane_stress_C(10e9, 0.3, dt), g[f"C_{tag}"]) < 1e-7 if dt == np.float32 else 1e-15
    # the reference's quirk: order 4 / 6 weights sum to 0.25
    assert abs(g["tri4_w_f64"].sum() - 0.25) < 1e-15 and abs(g["tri6_w_f64"].sum() - 0.25) < 1e-12
    assert abs(g["int2_w_f64"].sum() - 2.0) < 1e-15          # raw [-1,1] rule (Q3)


@pytest.mark.parametrize("case", TRI_CASES)
@pytest.mark.parametrize("tag", ["default", "forces"])
def test_tri_closed_form_vs_reference(case, tag):
    g = gold(case)
    dt = g["node_coords_free"].dtype
    loss, gx, gu = tri_oracle(g, tag)
    tol = tol_of(dt)
    assert abs(float(loss) - float(g[f"loss_{tag}"])) <= tol * abs(float(g[f"loss_{tag}"]))
    assert relmax(gx, g[f"gx_{tag}"]) < tol
    assert relmax(gu, g[f"gu_{tag}"]) < tol


@pytest.mark.parametrize("case", TRI_CASES)
def test_tri_generic_forward_vs_reference(case):
    g = gold(case)
    dt = g["node_coords_free"].dtype
    fmask, umask = ~g["boundary_mask"], ~g["dirichlet_mask"]
    coords = cf.assemble_full(g["node_coords_free"], g["node_coords_fixed"], fmask)
    U = cf.assemble_full(g["u_free"], np.zeros((int((~umask).sum()), 2), dt), umask)
    u_h, det, G = cf.tri_forward_points(coords, U, g["connectivity"], g["pt_x"], g["pt_e"])
    tol = tol_of(dt)
    assert relmax(u_h, g["pt_u"]) < tol and relmax(det, g["pt_det"]) < tol and relmax(G, g["pt_G"]) < 5 * tol
    ue, ds = cf.tri_edge_forward(coords, U, g["neumann_edges"], g["ed_x"][:, 0], g["ed_e"])
    assert relmax(ue, g["ed_u"]) < tol and relmax(ds, g["ed_ds"]) < tol


@pytest.mark.parametrize("case", ["tri_f64_jitter", "tri_f64_inverted", "tri_f32_jitter"])
@pytest.mark.parametrize("tag", ["default", "forces"])
def test_tri_torch_port_vs_reference(case, tag):
    g = gold(case)
    dt = torch.float64 if g["node_coords_free"].dtype == np.float64 else torch.float32
    T = torch.tensor
    fmask = ~g["boundary_mask"]
    coords = cf.assemble_full(g["node_coords_free"], g["node_coords_fixed"], fmask)
    m = tp.TriPort(T(coords), T(g["connectivity"]), T(g["boundary_mask"]), T(g["dirichlet_mask"]), float(g["u_fixed"]),
                   T(g["neumann_edges"]), u_free=T(g["u_free"]))
    xg, wg = cf.triangle_gauss_points(int(g["gauss_order"]), g["node_coords_free"].dtype)
    xi1, w1 = cf.interval_gauss_points(int(g["gauss_order_1d"]), g["node_coords_free"].dtype)
    C = T(cf.plane_stress_C(10e9, 0.3, g["node_coords_free"].dtype))
    bf, tf = (forces.b_force_test, forces.t_force_test) if tag == "forces" else (None, None)
    loss = tp.tri_energy(m, C, T(xg), T(wg), T(xi1), T(w1), bf, tf)
    loss.backward()
    tol = 1e-12 if dt == torch.float64 else 2e-5
    assert abs(loss.item() - float(g[f"loss_{tag}"])) <= tol * abs(float(g[f"loss_{tag}"]))
    assert relmax(m.x_free.grad.numpy(), g[f"gx_{tag}"]) < tol
    assert relmax(m.u_free.grad.numpy(), g[f"gu_{tag}"]) < tol


def test_lookup_bit_exact():
    g = gold("one_d")
    idx = cf.lookup_1d(g["lk_grid"], g["lk_x"], 4)
    assert np.array_equal(idx, g["lk_idx"])            # Q14: [-1,0,.5,1,2.999,3,4,...] -> [0,0,0,0,2,2,2,...]
    u, e = cf.interp_1d(g["lk_grid"], np.array([1.0, -2.0, 0.5, 3.0], np.float32), g["lk_x"])
    assert np.array_equal(e, g["lk_idx"]) and relmax(u, g["lk_u"]) < 1e-6


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("mode", ["f", "r"])
def test_1d_l2_closed_form(tag, mode):
    g = gold("one_d")
    dt = np.float64 if tag == "f64" else np.float32
    k = f"ex1_{tag}_{mode}"
    xt = np.linspace(0, 1, 1000).astype(dt) if dt == np.float64 else torch.linspace(0, 1, 1000).numpy()
    ut = torch.sin(2 * torch.pi * torch.tensor(xt)).numpy()
    u = g[k + "_u"]
    if mode == "r":
        grid, aux = cf.grid_1d(g[k + "_p"], dt(0.0), dt(1.0))
    else:
        grid = g[k + "_grid"]
    tol = 1e-12 if dt == np.float64 else 3e-5
    assert relmax(grid, g[k + "_grid"]) < (1e-14 if dt == np.float64 else 1e-6)
    grid = g[k + "_grid"]
    pred, e = cf.interp_1d(grid, u, xt)
    assert relmax(pred, g[k + "_pred"]) < tol
    r = 2.0 * (pred - ut) / pred.size
    dg, du, _ = cf.interp_1d_backward(grid, u, xt, r)
    assert relmax(du, g[k + "_gu"]) < tol
    if mode == "r":
        dp = cf.grid_1d_backward(dg, g[k + "_p"], aux)
        assert relmax(dp, g[k + "_gp"]) < (1e-11 if dt == np.float64 else 2e-4)


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("npts", [89, 300])
def test_bar_energy_closed_form(tag, npts):
    g = gold("one_d")
    dt = np.float64 if tag == "f64" else np.float32
    k = f"ex3_{tag}_{npts}"
    p, u = g[k + "_p"], g[k + "_u"]
    grid, aux = cf.grid_1d(p, dt(0.0), dt(10.0))
    assert relmax(grid, g[k + "_grid"]) < (1e-14 if dt == np.float64 else 1e-6)
    ufull = np.concatenate([[dt(0)], u, [dt(0)]]).astype(dt)
    xi, wi = cf.interval_gauss_points(int(g[k + "_ng"]), dt)
    loss, dG, dU = cf.bar_energy(grid, ufull, xi, wi, dt(175.0))
    tol = 1e-11 if dt == np.float64 else 1e-4
    assert abs(float(loss) - float(g[k + "_loss"])) <= tol * abs(float(g[k + "_loss"]))
    assert relmax(dU[1:-1], g[k + "_gu"]) < tol
    dp = cf.grid_1d_backward(dG, p, aux)
    assert relmax(dp, g[k + "_gp"]) < (1e-10 if dt == np.float64 else 2e-3)


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("ftag", ["free", "fix"])
def test_structured_closed_form(tag, ftag):
    g = gold("structured")
    dt = np.float64 if tag == "f64" else np.float32
    k = f"q1_{tag}_{ftag}"
    gx, ax = cf.grid_1d(g[k + "_px"], dt(0.0), dt(1.0))
    gy, ay = cf.grid_1d(g[k + "_py"], dt(0.0), dt(1.0))
    # boundary coordinates re-imposed from the initial grid (models.py:165-166)
    gx[0], gx[-1], gy[0], gy[-1] = 0.0, 1.0, 0.0, 1.0
    assert relmax(gx, g[k + "_gx"]) < (1e-14 if dt == np.float64 else 1e-6)
    gx, gy = g[k + "_gx"], g[k + "_gy"]
    u = g[k + "_u"].copy()
    mask = np.zeros(u.shape, bool)
    mask[0, :] = mask[-1, :] = mask[:, 0] = mask[:, -1] = True
    if ftag == "fix":
        u = np.where(mask, dt(g[k + "_ufix"]), u)
    x, ut = g[k + "_x"], g[k + "_ut"]
    pred, ix, iy = cf.q1_interp(gx, gy, u, x)
    assert np.array_equal(ix, g[k + "_ix"]) and np.array_equal(iy, g[k + "_iy"])      # lookup bit-exact
    tol = 1e-12 if dt == np.float64 else 3e-5
    assert relmax(pred, g[k + "_pred"]) < tol
    r = (2.0 * (pred - ut) / pred.size).astype(dt)
    dgx, dgy, dU = cf.q1_interp_backward(gx, gy, u, x, r)
    if ftag == "fix":
        dU = np.where(mask, 0.0, dU)
    assert relmax(dU, g[k + "_gu"]) < tol
    dgx[0] = dgx[-1] = 0.0
    dgy[0] = dgy[-1] = 0.0
    dpx = cf.grid_1d_backward(dgx, g[k + "_px"], ax)
    dpy = cf.grid_1d_backward(dgy, g[k + "_py"], ay)
    gtol = 1e-10 if dt == np.float64 else 2e-3
    assert relmax(dpx, g[k + "_gpx"]) < gtol and relmax(dpy, g[k + "_gpy"]) < gtol


def test_correct_math_variants_of_the_oracle():
    """The default-off correct-math switches (J^-T, proper order-4/6 weights, [0,1] edge rule) have their own oracle
    variants: the closed form agrees with torch autograd over the same op sequence, a linear field on skew triangles
    returns its exact gradient (SURVEY Q1's counter-example is fixed), and the quadrature tables integrate 1 exactly."""
    import torch
    from oracle import torch_port as tp
    from hidenn_fem_b200 import meshgen
    m = meshgen.plate_mesh(13, 9, jitter=0.3, diag="random", seed=1, ordering="random")
    rng = np.random.default_rng(0)
    U = rng.standard_normal(m.node_coords.shape) * 1e-3
    C = cf.plane_stress_C()
    for order in (4, 6):
        assert abs(cf.triangle_gauss_points(order, fix_weights=True)[1].sum() - 0.5) < 1e-12
        assert abs(cf.triangle_gauss_points(order)[1].sum() - 0.25) < 1e-12                 # the reference's tables are untouched
    xg, wg = cf.triangle_gauss_points(4, fix_weights=True)
    xi1, w1 = cf.interval_gauss_points(2, unit_interval=True)
    assert abs(w1.sum() - 1.0) < 1e-15 and (xi1 > 0).all() and (xi1 < 1).all()
    lo, dX, dU = cf.tri_energy_full(m.node_coords, U, m.connectivity, C, xg, wg, None, m.neumann_edges, xi1, w1, jinv_transpose=True)
    T = torch.tensor
    none = np.zeros(m.node_coords.shape[0], bool)
    port = tp.TriPort(T(m.node_coords), T(m.connectivity), T(none), T(none), 0.0, T(m.neumann_edges), u_free=T(U))
    l = tp.tri_energy(port, T(C), T(xg), T(wg), T(xi1), T(w1), jinv_transpose=True)
    l.backward()
    assert abs(float(l.detach()) - lo) <= 1e-13 * abs(lo)
    assert relmax(dX, port.x_free.grad.numpy()) < 1e-13 and relmax(dU, port.u_free.grad.numpy()) < 1e-13
    A = np.array([[3.0, 5.0], [7.0, 11.0]])
    Ne = m.connectivity.shape[0]
    xr = np.full((Ne, 2), 1.0 / 3.0)
    _, _, G = cf.tri_forward_points(m.node_coords, m.node_coords @ A.T, m.connectivity, xr, np.arange(Ne), jinv_transpose=True)
    assert np.abs(G - A).max() < 1e-11
    _, _, Gq = cf.tri_forward_points(m.node_coords, m.node_coords @ A.T, m.connectivity, xr, np.arange(Ne))
    assert np.abs(Gq - A).max() > 0.1                                                      # the reference's J^-1 does not
