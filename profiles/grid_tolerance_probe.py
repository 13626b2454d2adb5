"""Where do the grid-path gradient errors come from?  For every 1D / structured golden case: error of the CUDA
gradient w.r.t. the increments against (a) the reference's own golden values, (b) an FP64 oracle evaluation of the
same (rounded) parameters, (c) piecewise: d loss/d grid on the SAME grid bits and the softplus/cumsum chain on the
SAME cotangent bits, each against an FP64 evaluation.  Run on the GPU box: python profiles/grid_tolerance_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
from helpers import gold, relmax
from oracle import closed_form as cf
from hidenn_fem_b200 import models_grid as mg
from hidenn_fem_b200.models import PiecewiseLinearShapeNN, StructuredShapeNN2D
T = lambda a, **k: torch.tensor(a, device="cuda", **k)
f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)

g = gold("one_d")
for tag in ("f32", "f64"):
    dt = torch.float64 if tag == "f64" else torch.float32
    k = f"ex1_{tag}_r"
    xg = torch.linspace(0, 1, 100, dtype=dt)
    model = PiecewiseLinearShapeNN(xg, r_adapt=True)
    model = (model.double() if tag == "f64" else model).cuda()
    with torch.no_grad():
        model.u.copy_(T(g[k + "_u"])); model.x_increments.copy_(T(g[k + "_p"]))
    xt = torch.linspace(0, 1, 1000, dtype=dt).cuda()
    ut = torch.sin(2 * torch.pi * xt)
    ((model(xt) - ut) ** 2).mean().backward()
    gp = f64(model.x_increments.grad)
    # (b) FP64 oracle on the same parameters
    p64, u64, x64, ut64 = f64(model.x_increments), f64(model.u_full), f64(xt), f64(ut)
    grid64, aux = cf.grid_1d(p64, np.float64(0.0), np.float64(f64(model.xN)[0]))
    pred, _ = cf.interp_1d(grid64, u64, x64)
    r = 2.0 * (pred - ut64) / x64.size
    dg, du, _ = cf.interp_1d_backward(grid64, u64, x64, r)
    dp_b = cf.grid_1d_backward(dg, p64, aux)
    # (c) pieces on identical bits
    grid_bits = model.grid.detach()
    gl = grid_bits.clone().requires_grad_(True)
    pr = mg._Interp1DFn.apply(gl, model.u_full.detach(), xt)
    ((pr - ut) ** 2).mean().backward()
    dG_gpu = gl.grad
    predc, _ = cf.interp_1d(f64(grid_bits), u64, x64)
    rc = 2.0 * (predc - ut64) / x64.size
    dgc, duc, _ = cf.interp_1d_backward(f64(grid_bits), u64, x64, rc)
    pl = model.x_increments.detach().clone().requires_grad_(True)
    gg = mg._GridFn.apply(pl, model.x0, model.xN)
    gg.backward(dG_gpu)
    dp_chain = cf.grid_1d_backward(f64(dG_gpu), p64, aux)
    print(f"{k}: gp vs golden {relmax(gp, g[k+'_gp']):.2e} | vs fp64(same params) {relmax(gp, dp_b):.2e} | golden vs fp64 {relmax(g[k+'_gp'], dp_b):.2e}"
          f" | dG same grid bits {relmax(f64(dG_gpu), dgc):.2e} | chain same cotangent {relmax(f64(pl.grad), dp_chain):.2e}"
          f" | pred same bits {relmax(f64(pr), predc):.2e} | gu {relmax(f64(model.u.grad), du):.2e}")
