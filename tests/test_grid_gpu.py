"""GPU parity of the 1D and structured-2D paths vs the reference's golden fixtures and the oracle.
FP64 1e-10 / FP32 1e-5 on values; lookup indices bit-exact."""
import numpy as np
import pytest
import torch

from helpers import gold, relmax
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
T = lambda a, **k: torch.tensor(a, device="cuda", **k)
f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)
TOL = {torch.float64: 1e-10, torch.float32: 1e-5}          # BASELINE.json north_star


def chain_check(p_param, x0, xN, d_grid, tol):
    """softplus -> clamp -> cumsum -> normalise chain (models.py:45-53), CUDA VJP for the cotangent `d_grid` against the
    FP64 evaluation of the same parameter and cotangent bits."""
    from hidenn_fem_b200 import models_grid as mg
    pl = p_param.detach().clone().requires_grad_(True)
    mg._GridFn.apply(pl, x0, xN).backward(d_grid)
    p64 = f64(p_param)
    _, aux = cf.grid_1d(p64, np.float64(f64(x0)[0]), np.float64(f64(xN)[0]))
    want = cf.grid_1d_backward(f64(d_grid), p64, aux)
    assert relmax(f64(pl.grad), want) < tol, relmax(f64(pl.grad), want)
    return want


def e2e_vs_golden(got, golden, exact, tol):
    """End to end the increment gradients are ill-conditioned in the input rounding (a 1-ulp change of a grid coordinate
    moves 1/h by N ulp), in the reference's own run as much as here: the pieces are held to the contract tolerance on
    identical bits (callers), and the composed result must be as close to the FP64 evaluation of the same parameters
    as the reference's own golden value is (factor 4), or within the contract tolerance of the golden."""
    e_got, e_gold = relmax(got, exact), relmax(golden, exact)
    assert relmax(got, golden) < tol or e_got <= 4.0 * e_gold + tol, (relmax(got, golden), e_got, e_gold)


def make_1d(g, k, dt, r_adapt, u0=None, uN=None, npts=100, L=1.0):
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN
    xg = torch.linspace(0, L, npts, dtype=dt)
    model = PiecewiseLinearShapeNN(xg, r_adapt=r_adapt, u0=u0, uN=uN)
    if dt == torch.float64:
        model = model.double()
    model = model.cuda()
    with torch.no_grad():
        model.u.copy_(T(g[k + "_u"]))
        if r_adapt:
            model.x_increments.copy_(T(g[k + "_p"]))
    return model


def test_lookup_bit_exact_gpu():
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN
    g = gold("one_d")
    m = PiecewiseLinearShapeNN(torch.tensor(g["lk_grid"])).cuda()
    with torch.no_grad():
        m.u.copy_(T([1.0, -2.0, 0.5, 3.0]))
    x = T(g["lk_x"])
    assert np.array_equal(m.lookup(x).cpu().numpy(), g["lk_idx"])
    assert relmax(m(x).detach().cpu().numpy(), g["lk_u"]) < 1e-6
    # large random check against numpy searchsorted on a non-uniform grid, incl. points exactly on nodes
    rng = np.random.default_rng(0)
    grid = np.sort(rng.random(5001))
    xs = np.concatenate([rng.random(20000) * 1.2 - 0.1, grid[::7]])
    m2 = PiecewiseLinearShapeNN(torch.tensor(grid)).cuda()
    got = m2.lookup(T(xs)).cpu().numpy()
    assert np.array_equal(got, cf.lookup_1d(grid, xs, grid.size))


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("mode", ["f", "r"])
def test_example1_l2_step_and_adam(tag, mode):
    g = gold("one_d")
    dt = torch.float64 if tag == "f64" else torch.float32
    k = f"ex1_{tag}_{mode}"
    model = make_1d(g, k, dt, mode == "r")
    from hidenn_fem_b200 import models_grid as mg
    tol = TOL[dt]
    assert relmax(model.grid.detach().cpu().numpy(), g[k + "_grid"]) < (1e-13 if dt == torch.float64 else 1e-6)
    xt = torch.linspace(0, 1, 1000, dtype=dt).cuda()
    ut = torch.sin(2 * torch.pi * xt)
    pred = model(xt)
    loss = ((pred - ut) ** 2).mean()
    loss.backward()
    # (1) every kernel against an FP64 evaluation of the SAME input bits, at the contract tolerance
    grid_bits = model.grid.detach()
    gl, ul = grid_bits.clone().requires_grad_(True), model.u_full.detach().clone().requires_grad_(True)
    pr = mg._Interp1DFn.apply(gl, ul, xt)
    lp = ((pr - ut) ** 2).mean()
    lp.backward()
    assert torch.equal(pr, pred)
    u64, x64, ut64 = f64(ul), f64(xt), f64(ut)
    po, _ = cf.interp_1d(f64(grid_bits), u64, x64)
    lo = ((po - ut64) ** 2).mean()
    dgo, duo, _ = cf.interp_1d_backward(f64(grid_bits), u64, x64, 2.0 * (po - ut64) / x64.size)
    assert relmax(f64(pr), po) < tol and abs(lp.item() - lo) <= tol * abs(lo)
    assert relmax(f64(ul.grad), duo) < tol
    if mode == "r":
        assert relmax(f64(gl.grad), dgo) < tol
        chain_check(model.x_increments, model.x0, model.xN, gl.grad, tol)
    # (2) composed result against the reference's golden run
    if mode == "f" or dt == torch.float64:
        assert relmax(pred.detach().cpu().numpy(), g[k + "_pred"]) < tol
        assert abs(loss.item() - float(g[k + "_loss"])) <= tol * abs(float(g[k + "_loss"]))
        assert relmax(model.u.grad.cpu().numpy(), g[k + "_gu"]) < tol
        if mode == "r":
            assert relmax(model.x_increments.grad.cpu().numpy(), g[k + "_gp"]) < tol
    else:
        p64 = f64(model.x_increments)
        grid64, aux = cf.grid_1d(p64, np.float64(f64(model.x0)[0]), np.float64(f64(model.xN)[0]))
        pe, _ = cf.interp_1d(grid64, u64, x64)
        dge, due, _ = cf.interp_1d_backward(grid64, u64, x64, 2.0 * (pe - ut64) / x64.size)
        e2e_vs_golden(f64(pred), g[k + "_pred"], pe, tol)
        e2e_vs_golden(f64(model.u.grad), g[k + "_gu"], due, tol)
        e2e_vs_golden(f64(model.x_increments.grad), g[k + "_gp"], cf.grid_1d_backward(dge, p64, aux), tol)
    # unchanged Adam loop of examples/example1.py:31-40
    model.zero_grad()
    opt = torch.optim.Adam(model.parameters(), lr=0.005)
    tr = []
    for _ in range(25):
        opt.zero_grad()
        l = ((model(xt) - ut) ** 2).mean()
        l.backward()
        opt.step()
        tr.append(l.item())
    assert np.allclose(tr, g[k + "_adam"], rtol=1e-7 if dt == torch.float64 else 2e-3)


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("npts", [89, 300])
@pytest.mark.parametrize("path", ["fused_builtin", "fused_callable", "generic"])
def test_example3_bar_energy(tag, npts, path):
    from hidenn_fem_b200 import models_grid as mg
    from hidenn_fem_b200.utils import interval_gauss_points
    g = gold("one_d")
    dt = torch.float64 if tag == "f64" else torch.float32
    k = f"ex3_{tag}_{npts}"
    model = make_1d(g, k, dt, True, u0=0.0, uN=0.0, npts=npts, L=10.0)
    xi, wi = interval_gauss_points(int(g[k + "_ng"]), device="cuda", dtype=dt)
    if path == "fused_builtin":
        loss = mg.bar_energy_loss(model, xi, wi, None, 175.0, b_builtin=True)
    elif path == "fused_callable":
        loss = mg.bar_energy_loss(model, xi, wi, mg.example3_b_force, 175.0)
    else:
        loss = mg.energy_loss_generic(model, xi, wi, mg.example3_b_force, E=175.0)
    loss.backward()
    tol = TOL[dt]
    mg._bar_state.check(block=True)
    # (1) the energy kernel on the SAME grid bits and the grid chain on the SAME cotangent bits vs FP64, contract tolerance
    grid_bits = model.grid.detach()
    gl, ul = grid_bits.clone().requires_grad_(True), model.u_full.detach().clone().requires_grad_(True)
    if path == "fused_builtin":
        lp = mg._BarEnergyFn.apply(gl, ul, xi, wi, 175.0, None, mg._bar_state)
    elif path == "fused_callable":
        with torch.no_grad():
            xq = 0.5 * (grid_bits[1:, None] - grid_bits[:-1, None]) * xi + 0.5 * (grid_bits[1:, None] + grid_bits[:-1, None])
            bt = mg.example3_b_force(xq).contiguous()
        lp = mg._BarEnergyFn.apply(gl, ul, xi, wi, 175.0, bt, mg._bar_state)
    else:
        lp = None
    if lp is not None:
        lp.backward()
        xin, win = cf.interval_gauss_points(int(g[k + "_ng"]))
        lo, dGo, dUo = cf.bar_energy(f64(grid_bits), f64(ul), f64(xi), f64(wi), 175.0)
        assert abs(lp.item() - lo) <= tol * abs(lo), (lp.item(), lo)
        assert relmax(f64(ul.grad), dUo) < tol and relmax(f64(gl.grad), dGo) < tol
        chain_check(model.x_increments, model.x0, model.xN, gl.grad, tol)
    # (2) composed result against the reference's golden run
    if dt == torch.float64:
        assert abs(loss.item() - float(g[k + "_loss"])) <= tol * abs(float(g[k + "_loss"]))
        assert relmax(model.u.grad.cpu().numpy(), g[k + "_gu"]) < tol
        assert relmax(model.x_increments.grad.cpu().numpy(), g[k + "_gp"]) < tol
    else:
        p64 = f64(model.x_increments)
        grid64, aux = cf.grid_1d(p64, np.float64(0.0), np.float64(f64(model.xN)[0]))
        le, dGe, dUe = cf.bar_energy(grid64, f64(model.u_full), f64(xi), f64(wi), 175.0)
        e2e_vs_golden(np.asarray([loss.item()]), np.asarray([float(g[k + "_loss"])]), np.asarray([le]), tol)
        e2e_vs_golden(f64(model.u.grad), g[k + "_gu"], dUe[1:-1], tol)
        e2e_vs_golden(f64(model.x_increments.grad), g[k + "_gp"], cf.grid_1d_backward(dGe, p64, aux), tol)
    if path == "generic" or dt == torch.float32:
        return
    # unchanged Adam loop of examples/example3.py:89-96 on the fused loss
    model.zero_grad()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    tr = []
    for _ in range(15):
        opt.zero_grad()
        l = mg.bar_energy_loss(model, xi, wi, mg.example3_b_force, 175.0, b_builtin=(path == "fused_builtin"))
        l.backward()
        opt.step()
        tr.append(l.item())
    assert np.allclose(tr, g[k + "_adam"], rtol=1e-8)


def test_bar_energy_1m_vs_oracle():
    """Config C2: 1M elements FP64; oracle = closed form on the same grid."""
    from hidenn_fem_b200 import models_grid as mg
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN
    from hidenn_fem_b200.utils import interval_gauss_points
    N = 1_000_001
    xg = torch.linspace(0, 10.0, N, dtype=torch.float64)
    model = PiecewiseLinearShapeNN(xg, r_adapt=True, u0=0.0, uN=0.0).double().cuda()
    gen = torch.Generator().manual_seed(0)
    with torch.no_grad():
        model.u.copy_((1e-2 * torch.randn(N - 2, generator=gen, dtype=torch.float64)).cuda())
        model.x_increments.add_((0.1 * 1e-5 * torch.randn(N - 1, generator=gen, dtype=torch.float64)).cuda())
    xi, wi = interval_gauss_points(2, device="cuda", dtype=torch.float64)
    loss = mg.bar_energy_loss(model, xi, wi, None, 175.0, b_builtin=True)
    loss.backward()
    mg._bar_state.check(block=True)
    p = model.x_increments.detach().cpu().numpy()
    u = model.u.detach().cpu().numpy()
    grid, aux = cf.grid_1d(p, np.float64(0.0), np.float64(10.0))
    gpu_grid = model.grid.detach().cpu().numpy()
    assert relmax(gpu_grid, grid) < 1e-13          # scan vs numpy cumsum: a few ulp of the coordinates
    # With h = 1e-5 on [0,10] a 1-ulp change of a coordinate moves h by 1e-10 relative, so the energy kernel is
    # checked on the SAME grid bits the GPU produced (the oracle's own sensitivity to its cumsum order is as large).
    ufull = np.concatenate([[0.0], u, [0.0]])
    xin, win = cf.interval_gauss_points(2)
    lo, dG, dU = cf.bar_energy(gpu_grid, ufull, xin, win, 175.0)
    assert abs(loss.item() - lo) <= 1e-10 * abs(lo)
    assert relmax(model.u.grad.cpu().numpy(), dU[1:-1]) < 1e-10
    # grid chain on the cotangent bits the GPU produced
    gl = model.grid.detach().clone().requires_grad_(True)
    mg._BarEnergyFn.apply(gl, model.u_full.detach(), xi, wi, 175.0, None, mg._bar_state).backward()
    assert relmax(f64(gl.grad), dG) < 1e-10
    chain_check(model.x_increments, model.x0, model.xN, gl.grad, 1e-10)


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("ftag", ["free", "fix"])
def test_structured_q1(tag, ftag):
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN2D, StructuredShapeNN2D
    g = gold("structured")
    dt = torch.float64 if tag == "f64" else torch.float32
    k = f"q1_{tag}_{ftag}"
    Nx, Ny = 25, 19
    gx, gy = torch.linspace(0, 1, Nx, dtype=dt), torch.linspace(0, 1, Ny, dtype=dt)
    ufix = float(g[k + "_ufix"]) if ftag == "fix" else None
    # the reference name with grid_x/grid_y keywords selects the structured class (models.py:93 vs :241)
    model = PiecewiseLinearShapeNN2D(grid_x=gx, grid_y=gy, boundary_mask_x=None, boundary_mask_y=None, r_adapt=True, u_fixed=ufix)
    assert isinstance(model, StructuredShapeNN2D)
    if dt == torch.float64:
        model = model.double()
    model = model.cuda()
    with torch.no_grad():
        model.increments_x.copy_(T(g[k + "_px"]))
        model.increments_y.copy_(T(g[k + "_py"]))
        model.u.copy_(T(g[k + "_u"]))
    gxx, gyy = model.grid
    assert relmax(gxx.detach().cpu().numpy(), g[k + "_gx"]) < (1e-13 if dt == torch.float64 else 1e-6)
    x, ut = T(g[k + "_x"]), T(g[k + "_ut"])
    pred = model(x)
    loss = ((pred - ut) ** 2).mean()
    loss.backward()
    tol = TOL[dt]
    assert relmax(pred.detach().cpu().numpy(), g[k + "_pred"]) < tol
    assert abs(loss.item() - float(g[k + "_loss"])) <= tol * abs(float(g[k + "_loss"]))
    assert relmax(model.u.grad.cpu().numpy(), g[k + "_gu"]) < tol
    # grid gradients: (1) the interpolation VJP on the SAME grid bits and the chain on the SAME cotangent bits vs FP64
    from hidenn_fem_b200 import models_grid as mg
    gxb, gyb = (a.detach() for a in model.grid)
    gxl, gyl = gxb.clone().requires_grad_(True), gyb.clone().requires_grad_(True)
    pr = mg._Q1InterpFn.apply(gxl, gyl, model.u_full.detach(), x)
    ((pr - ut) ** 2).mean().backward()
    x64, ut64, u64 = f64(x), f64(ut), f64(model.u_full)
    po, _, _ = cf.q1_interp(f64(gxb), f64(gyb), u64, x64)
    dgx, dgy, _ = cf.q1_interp_backward(f64(gxb), f64(gyb), u64, x64, 2.0 * (po - ut64) / x64.shape[0])
    assert relmax(f64(gxl.grad), dgx) < tol and relmax(f64(gyl.grad), dgy) < tol
    mx, my = ~model.boundary_mask_x, ~model.boundary_mask_y            # torch.where(mask, initial, grid): models.py:165-166
    wx = chain_check(model.increments_x, model.x0, model.xN, gxl.grad * mx, tol)
    wy = chain_check(model.increments_y, model.y0, model.yN, gyl.grad * my, tol)
    # (2) composed result against the reference's golden run
    if dt == torch.float64:
        assert relmax(model.increments_x.grad.cpu().numpy(), g[k + "_gpx"]) < tol
        assert relmax(model.increments_y.grad.cpu().numpy(), g[k + "_gpy"]) < tol
    else:
        px, py = f64(model.increments_x), f64(model.increments_y)
        gx64, ax = cf.grid_1d(px, np.float64(f64(model.x0)[0]), np.float64(f64(model.xN)[0]))
        gy64, ay = cf.grid_1d(py, np.float64(f64(model.y0)[0]), np.float64(f64(model.yN)[0]))
        gx64[f64(model.boundary_mask_x) > 0] = f64(model.initial_x_grid)[f64(model.boundary_mask_x) > 0]
        gy64[f64(model.boundary_mask_y) > 0] = f64(model.initial_y_grid)[f64(model.boundary_mask_y) > 0]
        pe, _, _ = cf.q1_interp(gx64, gy64, u64, x64)
        ex, ey, _ = cf.q1_interp_backward(gx64, gy64, u64, x64, 2.0 * (pe - ut64) / x64.shape[0])
        e2e_vs_golden(f64(model.increments_x.grad), g[k + "_gpx"], cf.grid_1d_backward(ex * f64(mx), px, ax), tol)
        e2e_vs_golden(f64(model.increments_y.grad), g[k + "_gpy"], cf.grid_1d_backward(ey * f64(my), py, ay), tol)
    # Adam loop of examples/example2.py:37-48 with the fixed sample set of the fixture
    model.zero_grad()
    opt = torch.optim.Adam(model.parameters(), lr=0.005)
    tr = []
    for _ in range(15):
        opt.zero_grad()
        l = ((model(x) - ut) ** 2).mean()
        l.backward()
        opt.step()
        tr.append(l.item())
    assert np.allclose(tr, g[k + "_adam"], rtol=1e-7 if dt == torch.float64 else 2e-3)


def test_structured_lookup_and_large_vs_oracle():
    from hidenn_fem_b200.models import StructuredShapeNN2D
    rng = np.random.default_rng(3)
    Nx, Ny, M = 257, 193, 200_000
    gx, gy = torch.linspace(0, 1, Nx, dtype=torch.float64), torch.linspace(0, 1, Ny, dtype=torch.float64)
    model = StructuredShapeNN2D(gx, gy, r_adapt=True).double().cuda()
    with torch.no_grad():
        model.increments_x.mul_(T(1.0 + 0.3 * rng.standard_normal(Nx - 1)))
        model.increments_y.mul_(T(1.0 + 0.3 * rng.standard_normal(Ny - 1)))
    x = rng.random((M, 2)) * 1.1 - 0.05
    ut = np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1])
    pred = model(T(x))
    loss = ((pred - T(ut)) ** 2).mean()
    loss.backward()
    gxx, gyy = (a.detach().cpu().numpy() for a in model.grid)
    u = model.u.detach().cpu().numpy()
    po, ix, iy = cf.q1_interp(gxx, gyy, u, x)
    assert relmax(pred.detach().cpu().numpy(), po) < 1e-12
    r = 2.0 * (po - ut) / M
    dgx, dgy, dU = cf.q1_interp_backward(gxx, gyy, u, x, r)
    assert relmax(model.u.grad.cpu().numpy(), dU) < 1e-10
    # grid gradients (the r-adaptive part): interpolation VJP on the same grid bits, then the chain on the same cotangent bits
    from hidenn_fem_b200 import models_grid as mg
    gxl, gyl = T(gxx).requires_grad_(True), T(gyy).requires_grad_(True)
    pr = mg._Q1InterpFn.apply(gxl, gyl, model.u_full.detach(), T(x))
    ((pr - T(ut)) ** 2).mean().backward()
    assert relmax(f64(gxl.grad), dgx) < 1e-10 and relmax(f64(gyl.grad), dgy) < 1e-10
    wx = chain_check(model.increments_x, model.x0, model.xN, gxl.grad * (~model.boundary_mask_x), 1e-10)
    wy = chain_check(model.increments_y, model.y0, model.yN, gyl.grad * (~model.boundary_mask_y), 1e-10)
    assert relmax(f64(model.increments_x.grad), wx) < 1e-10 and relmax(f64(model.increments_y.grad), wy) < 1e-10
    # determinism: a second evaluation is bit-identical
    g1 = model.u.grad.clone()
    model.zero_grad()
    ((model(T(x)) - T(ut)) ** 2).mean().backward()
    assert torch.equal(g1, model.u.grad)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_structured_fused_backward_matches_row_fold(dt):
    """Sort-free binning + fused cell fold vs materialised rows + stable sort + fold: the permutation is bit-exact
    (ascending sample id inside a cell), the sums agree to rounding (the fused kernel contracts a += r*(...) into FMAs)
    and are bit-identical run to run although the scatter uses integer atomics.  Includes cells with thousands of samples."""
    import ctypes as C
    from hidenn_fem_b200 import _lib
    rng = np.random.default_rng(11)
    Nx, Ny, M = 65, 49, 120_000
    gx = T(np.sort(rng.random(Nx)), dtype=dt)
    gy = T(np.sort(rng.random(Ny)), dtype=dt)
    uf = T(rng.standard_normal((Nx, Ny)), dtype=dt)
    xs = rng.random((M, 2))
    xs[:5000] = 0.5 + 1e-4 * rng.standard_normal((5000, 2))          # a few cells hold thousands of samples
    x = T(xs, dtype=dt)
    r = T(rng.standard_normal(M), dtype=dt)
    L, s, i64 = _lib.lib(), _lib.stream_ptr(), C.c_int64
    u = torch.empty(M, device="cuda", dtype=dt)
    ix = torch.empty(M, device="cuda", dtype=torch.int32)
    iy = torch.empty_like(ix)
    _lib.check(_lib.fn("hidenn_q1_interp_fwd", dt)(_lib.ptr(gx), i64(Nx), _lib.ptr(gy), i64(Ny), _lib.ptr(uf), _lib.ptr(x), i64(M),
                                                   _lib.ptr(u), _lib.ptr(ix), _lib.ptr(iy), s))
    ncell = (Nx - 1) * (Ny - 1)
    # (a) rows + stable sort
    rows = torch.empty(M, 8, device="cuda", dtype=dt)
    _lib.check(_lib.fn("hidenn_q1_interp_bwd", dt)(_lib.ptr(gx), i64(Nx), _lib.ptr(gy), i64(Ny), _lib.ptr(uf), _lib.ptr(x),
                                                   _lib.ptr(ix), _lib.ptr(iy), _lib.ptr(r), i64(M), _lib.ptr(rows), s))
    cell = ix.long() * (Ny - 1) + iy.long()
    _, order64 = torch.sort(cell, stable=True)
    seg64 = torch.zeros(ncell + 1, device="cuda", dtype=torch.int64)
    seg64[1:] = torch.cumsum(torch.bincount(cell, minlength=ncell), 0)
    outs = []
    for _ in range(2):
        outs.append((torch.empty(ncell, 8, device="cuda", dtype=dt), torch.empty(Nx, Ny, device="cuda", dtype=dt),
                     torch.empty(Nx, device="cuda", dtype=dt), torch.empty(Ny, device="cuda", dtype=dt)))
    t, du, dgx, dgy = outs[0]
    _lib.check(_lib.fn("hidenn_q1_fold_rows", dt)(_lib.ptr(rows), _lib.ptr(order64), _lib.ptr(seg64), i64(Nx), i64(Ny), _lib.ptr(t),
                                                  _lib.ptr(du), _lib.ptr(dgx), _lib.ptr(dgy), s))
    # (b) histogram + scatter + per-cell canonical order + fused fold
    cnt = torch.zeros(ncell, device="cuda", dtype=torch.int32)
    _lib.check(L.hidenn_q1_bin_count(_lib.ptr(ix), _lib.ptr(iy), i64(M), i64(Ny), _lib.ptr(cnt), s))
    assert torch.equal(cnt.long(), torch.bincount(cell, minlength=ncell))
    seg = torch.zeros(ncell + 1, device="cuda", dtype=torch.int32)
    seg[1:] = torch.cumsum(cnt, 0).int()
    cnt.zero_()
    order = torch.empty(M, device="cuda", dtype=torch.int32)
    _lib.check(L.hidenn_q1_bin_scatter(_lib.ptr(ix), _lib.ptr(iy), i64(M), i64(Nx), i64(Ny), _lib.ptr(seg), _lib.ptr(cnt),
                                       _lib.ptr(order), s))
    assert torch.equal(order.long(), order64)                        # index work: bit-exact with the stable sort
    t2, du2, dgx2, dgy2 = outs[1]
    _lib.check(_lib.fn("hidenn_q1_bwd_fused", dt)(_lib.ptr(gx), i64(Nx), _lib.ptr(gy), i64(Ny), _lib.ptr(uf), _lib.ptr(x), _lib.ptr(r),
                                                  i64(M), _lib.ptr(seg), _lib.ptr(order), _lib.ptr(t2), _lib.ptr(du2), _lib.ptr(dgx2),
                                                  _lib.ptr(dgy2), s))
    tol = 1e-12 if dt == torch.float64 else 2e-5
    for a, b in ((du, du2), (dgx, dgx2), (dgy, dgy2)):
        assert relmax(b.cpu().numpy(), a.cpu().numpy()) < tol
    # run-to-run: redo the (atomic) scatter and the fold
    cnt.zero_()
    order_b = torch.empty_like(order)
    _lib.check(L.hidenn_q1_bin_scatter(_lib.ptr(ix), _lib.ptr(iy), i64(M), i64(Nx), i64(Ny), _lib.ptr(seg), _lib.ptr(cnt),
                                       _lib.ptr(order_b), s))
    t3, du3, dgx3, dgy3 = (torch.empty_like(v) for v in outs[1])
    _lib.check(_lib.fn("hidenn_q1_bwd_fused", dt)(_lib.ptr(gx), i64(Nx), _lib.ptr(gy), i64(Ny), _lib.ptr(uf), _lib.ptr(x), _lib.ptr(r),
                                                  i64(M), _lib.ptr(seg), _lib.ptr(order_b), _lib.ptr(t3), _lib.ptr(du3), _lib.ptr(dgx3),
                                                  _lib.ptr(dgy3), s))
    torch.cuda.synchronize()
    assert torch.equal(order, order_b)
    assert torch.equal(du2, du3) and torch.equal(dgx2, dgx3) and torch.equal(dgy2, dgy3)


def test_structured_forward_smem_and_global_lookup_agree():
    """The shared-memory-staged forward (grid lines fit in smem, M >= 65536) and the global-memory one give the same bits."""
    from hidenn_fem_b200.models import StructuredShapeNN2D
    rng = np.random.default_rng(5)
    Nx, Ny = 1025, 513
    model = StructuredShapeNN2D(torch.linspace(0, 2, Nx, dtype=torch.float64), torch.linspace(0, 1, Ny, dtype=torch.float64),
                                r_adapt=True).double().cuda()
    with torch.no_grad():
        model.u.copy_(T(rng.standard_normal((Nx, Ny))) if model.u.dim() == 2 else T(rng.standard_normal(model.u.shape)))
    x = T(rng.random((100_000, 2)) * np.array([2.2, 1.1]) - 0.05)
    with torch.no_grad():
        big = model(x)                       # smem path
        small = torch.cat([model(x[i:i + 50_000 // 2]) for i in range(0, 100_000, 25_000)])   # < 65536 -> global path
    assert torch.equal(big, small)
    gxx, gyy = (a.detach().cpu().numpy() for a in model.grid)
    po, _, _ = cf.q1_interp(gxx, gyy, model.u.detach().cpu().numpy().reshape(Nx, Ny), x.cpu().numpy())
    assert relmax(big.cpu().numpy(), po) < 1e-12


def test_structured_full_size_properties():
    """BASELINE config C3 at full size (4097 x 4097 nodes, 2^26 samples, FP64) through oracle-free properties:
    a bilinear field is reproduced, u -> 2u scales prediction and gradients exactly, the nodal gradients of an MSE
    loss add up to the sum of the residual weights (partition of unity), and the step is bit-reproducible."""
    from hidenn_fem_b200.models import StructuredShapeNN2D
    Nx = Ny = 4097
    M = 1 << 26
    g1 = torch.linspace(0, 1, Nx, dtype=torch.float64)
    model = StructuredShapeNN2D(g1, g1.clone(), r_adapt=True).double().cuda()
    gen = torch.Generator(device="cuda").manual_seed(0)
    with torch.no_grad():
        model.increments_x.add_(0.3 * (torch.rand(Nx - 1, device="cuda", dtype=torch.float64, generator=gen) - 0.5) * model.increments_x.abs())
        model.increments_y.add_(0.3 * (torch.rand(Ny - 1, device="cuda", dtype=torch.float64, generator=gen) - 0.5) * model.increments_y.abs())
    x = torch.rand(M, 2, device="cuda", dtype=torch.float64, generator=gen)
    gx, gy = (a.detach() for a in model.grid)
    bil = lambda X, Y: 0.3 + 1.7 * X - 0.9 * Y + 2.3 * X * Y
    with torch.no_grad():
        model.u.copy_(bil(gx[:, None], gy[None, :]))
        pred = model(x)
    assert (pred - bil(x[:, 0], x[:, 1])).abs().max().item() < 1e-12            # bilinear reproduction

    target = torch.sin(6.0 * x[:, 0]) * torch.cos(4.0 * x[:, 1])

    def step():
        model.zero_grad(set_to_none=True)
        p = model(x)
        loss = ((p - target) ** 2).mean()
        loss.backward()
        return p.detach(), loss.item(), model.u.grad.clone(), model.increments_x.grad.clone()

    p1, l1, du1, dix1 = step()
    p2, l2, du2, dix2 = step()
    assert l1 == l2 and torch.equal(du1, du2) and torch.equal(dix1, dix2)          # deterministic (integer atomics only)
    r = 2.0 * (p1 - target) / M
    assert abs(du1.sum().item() - r.sum().item()) <= 1e-10 * r.abs().sum().item()  # partition of unity
    with torch.no_grad():
        model.u.mul_(2.0)
        p3 = model(x)
    assert torch.equal(p3, 2.0 * p1)                                               # exact power-of-two scaling


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_fused_l2_projection_loss_matches_generic_expression(dt):
    """models_grid.l2_projection_loss == ((model(x) - u_true) ** 2).mean() (examples/example2.py:45-46): loss and the
    gradients w.r.t. nodal values and both increment vectors, plus bit-reproducibility and grad_output scaling."""
    from hidenn_fem_b200.models import StructuredShapeNN2D
    from hidenn_fem_b200.models_grid import l2_projection_loss
    rng = np.random.default_rng(2)
    Nx, Ny, M = 129, 97, 150_000
    mk = lambda: StructuredShapeNN2D(torch.linspace(0, 1, Nx, dtype=dt), torch.linspace(0, 2, Ny, dtype=dt), r_adapt=True).to(dt).cuda()
    a, b = mk(), mk()
    with torch.no_grad():
        for m in (a, b):
            m.u.copy_(T(rng.standard_normal((Nx, Ny)), dtype=dt) if m is a else a.u)
            m.increments_x.copy_(a.increments_x * (1 if m is a else 1))
        pert = T(1.0 + 0.2 * rng.standard_normal(Nx - 1), dtype=dt)
        a.increments_x.mul_(pert); b.increments_x.mul_(pert)
    x = T(rng.random((M, 2)) * np.array([1.0, 2.0]), dtype=dt)
    ut = torch.sin(5 * x[:, 0]) * torch.cos(3 * x[:, 1])
    la = 3.0 * l2_projection_loss(a, x, ut)
    la.backward()
    lb = 3.0 * ((b(x) - ut) ** 2).mean()
    lb.backward()
    tol = 1e-11 if dt == torch.float64 else 2e-4
    assert abs(la.item() - lb.item()) <= tol * abs(lb.item())
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert relmax(pa.grad.cpu().numpy(), pb.grad.cpu().numpy()) < tol
    g1 = [p.grad.clone() for p in a.parameters()]
    a.zero_grad()
    (3.0 * l2_projection_loss(a, x, ut)).backward()
    assert all(torch.equal(u, p.grad) for u, p in zip(g1, a.parameters()))


def test_graphed_step_on_the_bar_energy():
    """graph.GraphedStep around the fused 1D bar energy: same bits as the eager step, follows in-place updates."""
    from hidenn_fem_b200.graph import GraphedStep
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN
    from hidenn_fem_b200 import models_grid as mg
    from hidenn_fem_b200.utils import interval_gauss_points
    N = 20_001
    mk = lambda: PiecewiseLinearShapeNN(torch.linspace(0, 10.0, N, dtype=torch.float64), r_adapt=True, u0=0.0, uN=0.0).double().cuda()
    a, b = mk(), mk()
    xi, wi = interval_gauss_points(2, device="cuda", dtype=torch.float64)
    with torch.no_grad():
        a.u.normal_(0, 1e-2)
        b.u.copy_(a.u)
    step = GraphedStep(b, lambda: mg.bar_energy_loss(b, xi, wi, None, 175.0, b_builtin=True))
    for _ in range(3):
        a.zero_grad()
        la = mg.bar_energy_loss(a, xi, wi, None, 175.0, b_builtin=True)
        la.backward()
        lb = step()
        assert la.item() == lb.item()
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert torch.equal(pa.grad, pb.grad)
        with torch.no_grad():
            for pa, pb in zip(a.parameters(), b.parameters()):
                pa.sub_(1e-4 * pa.grad)
                pb.sub_(1e-4 * pb.grad)
    mg._bar_state.check(block=True)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("M", [5_000, 200_000])
def test_structured_lookup_indices_bit_exact(dt, M):
    """ix / iy of the structured forward == clamp(searchsorted(grid, x) - 1, 0, N-2) (src/models.py:183-186), for the
    shared-memory (large M) and global (small M) variants, on grids far from uniform (geometric grading: the uniform
    guess is wrong by hundreds of lines, so the bisection fallback runs), with repeated grid values, points exactly
    on nodes, outside the domain and at +-inf."""
    import ctypes as C
    from hidenn_fem_b200 import _lib
    rng = np.random.default_rng(7)
    Nx, Ny = 1500, 700
    gx = np.cumsum(1.012 ** np.arange(Nx))
    gx = (gx - gx[0]) / (gx[-1] - gx[0])
    gy = np.sort(rng.random(Ny)) * 3.0 - 1.0
    gy[100:104] = gy[100]                                      # repeated nodes
    gy[0], gy[-1] = -1.0, 2.0
    gxt, gyt = T(gx, dtype=dt), T(gy, dtype=dt)
    gx, gy = gxt.cpu().numpy(), gyt.cpu().numpy()             # the values the kernel sees
    x = np.stack([rng.random(M) * 1.2 - 0.1, rng.random(M) * 3.4 - 1.2], 1).astype(gx.dtype)
    k = M // 4
    x[:k, 0] = gx[rng.integers(0, Nx, k)]                      # exactly on nodes
    x[:k, 1] = gy[rng.integers(0, Ny, k)]
    x[k:k + 4] = np.array([[-np.inf, np.inf], [np.inf, -np.inf], [gx[0], gy[-1]], [gx[-1], gy[0]]], dtype=gx.dtype)
    xt = T(x, dtype=dt)
    uf = T(rng.standard_normal((Nx, Ny)), dtype=dt)
    u = torch.empty(M, device="cuda", dtype=dt)
    ix = torch.empty(M, device="cuda", dtype=torch.int32)
    iy = torch.empty_like(ix)
    _lib.check(_lib.fn("hidenn_q1_interp_fwd", dt)(_lib.ptr(gxt), C.c_int64(Nx), _lib.ptr(gyt), C.c_int64(Ny), _lib.ptr(uf), _lib.ptr(xt),
                                                   C.c_int64(M), _lib.ptr(u), _lib.ptr(ix), _lib.ptr(iy), _lib.stream_ptr()))
    ex = torch.clamp(torch.searchsorted(gxt, xt[:, 0].contiguous()) - 1, 0, Nx - 2)
    ey = torch.clamp(torch.searchsorted(gyt, xt[:, 1].contiguous()) - 1, 0, Ny - 2)
    assert torch.equal(ix.long(), ex) and torch.equal(iy.long(), ey)
