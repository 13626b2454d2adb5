"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; CPU only)

The reference is imported from /root/reference (never copied).  Its structured-2D class is
shadowed by the triangle class of the same name (src/models.py:93 vs :241) and
examples/example3.py cannot be imported (ImportError at :5, module-level training), so
those two pieces are obtained by AST extraction + exec of the reference's own source text
at run time.  Outputs are small .npz fixtures that travel to the GPU box.
"""
import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("HIDENN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import src.models as ref_models          # noqa: E402
import src.loss as ref_loss              # noqa: E402
import src.utils as ref_utils            # noqa: E402

from hidenn_fem_b200 import meshgen      # noqa: E402


def _extract(path, kind, name, which=0):
    src = open(path).read()
    tree = ast.parse(src)
    nodes = [n for n in tree.body if isinstance(n, kind) and n.name == name]
    seg = ast.get_source_segment(src, nodes[which])
    ns = {"torch": torch, "nn": nn, "F": F}
    exec(compile(seg, path, "exec"), ns)
    return ns[name]


StructuredRef = _extract(os.path.join(REF, "src/models.py"), ast.ClassDef, "PiecewiseLinearShapeNN2D", 0)
ex3_b_force = _extract(os.path.join(REF, "examples/example3.py"), ast.FunctionDef, "b_force")
ex3_energy = _extract(os.path.join(REF, "examples/example3.py"), ast.FunctionDef, "energy_loss")

T = torch.tensor


def save(name, **kw):
    out = {}
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


# ----------------------------------------------------------------------------- quadrature
def gold_quadrature():
    d = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        for o in (1, 3, 4, 6, 7):
            rs, w = ref_utils.triangle_gauss_points(o, device=torch.device("cpu"), dtype=dt)
            d[f"tri{o}_rs_{tag}"], d[f"tri{o}_w_{tag}"] = rs, w
        for o in (1, 2, 3, 4, 5):
            x, w = ref_utils.interval_gauss_points(o, device=torch.device("cpu"), dtype=dt)
            d[f"int{o}_x_{tag}"], d[f"int{o}_w_{tag}"] = x, w
        L = ref_loss.EnergyLoss2D(E=10e9, nu=0.3, device=torch.device("cpu"), dtype=dt)
        d[f"C_{tag}"] = L.C
    save("quadrature", **d)


# ----------------------------------------------------------------------------- triangle energy
sys.path.insert(0, HERE)
from forces import b_force_test, t_force_test   # noqa: E402


def tri_case(name, nx, ny, dtype, ordering, jitter, invert, u_scale, free_neumann, gauss_order=4, g1=2, seed=0,
             u_fixed=0.0):
    m = meshgen.plate_mesh(nx, ny, jitter=jitter, diag="random", seed=seed, ordering=ordering)
    conn = meshgen.invert_some_elements(m.connectivity, invert, seed) if invert > 0 else m.connectivity
    bmask = m.boundary_mask.copy()
    if free_neumann:      # let the Neumann-side nodes move so the edge term has coordinate gradients
        bmask &= ~m.neumann_mask
    coords = T(m.node_coords, dtype=dtype)
    model = ref_models.PiecewiseLinearShapeNN2D(coords, T(conn), boundary_mask=T(bmask),
                                                dirichlet_mask=T(m.dirichlet_mask), u_fixed=u_fixed,
                                                neumann_edges=T(m.neumann_edges))
    g = torch.Generator().manual_seed(seed + 11)
    with torch.no_grad():
        model.u_free.copy_(u_scale * torch.randn(model.u_free.shape, generator=g, dtype=torch.float64).to(torch.float32))
    if dtype == torch.float64:
        model = model.double()
    loss_fn = ref_loss.EnergyLoss2D(E=10e9, nu=0.3, gauss_order=gauss_order, gauss_order_1d=g1,
                                    device=torch.device("cpu"), dtype=dtype)
    out = dict(node_coords=m.node_coords, connectivity=conn, boundary_mask=bmask, dirichlet_mask=m.dirichlet_mask,
               neumann_edges=m.neumann_edges, u_fixed=np.asarray(u_fixed), gauss_order=gauss_order, gauss_order_1d=g1,
               node_coords_free=model.node_coords_free, u_free=model.u_free, node_coords_fixed=model.node_coords_fixed)
    for tag, bf, tf in (("default", None, None), ("forces", b_force_test, t_force_test)):
        model.zero_grad()
        loss = loss_fn(model, bf, tf)
        loss.backward()
        out[f"loss_{tag}"] = loss
        out[f"gx_{tag}"] = model.node_coords_free.grad.clone()
        out[f"gu_{tag}"] = model.u_free.grad.clone()
        model.zero_grad()
        dom = loss_fn.domain_energy(model, bf)
        edge = loss_fn.edge_energy(model, tf)
        out[f"domain_{tag}"], out[f"edge_{tag}"] = dom, edge
    # generic forward at random reference points / random elements (a6) and on edges (a7)
    g2 = torch.Generator().manual_seed(seed + 5)
    M = 257
    eid = torch.randint(0, model.Nelems, (M,), generator=g2)
    xr = torch.rand(M, 2, generator=g2, dtype=torch.float64).to(dtype) * 0.5
    u_h, det, G = model(xr, eid)
    out.update(pt_x=xr, pt_e=eid, pt_u=u_h, pt_det=det, pt_G=G)
    # VJP of the generic forward with fixed cotangents
    cu = torch.randn(M, 2, generator=g2, dtype=torch.float64).to(dtype)
    cd = torch.randn(M, generator=g2, dtype=torch.float64).to(dtype)
    cG = torch.randn(M, 2, 2, generator=g2, dtype=torch.float64).to(dtype)
    model.zero_grad()
    ((u_h * cu).sum() + (det * cd).sum() + (G * cG).sum()).backward()
    out.update(pt_cu=cu, pt_cd=cd, pt_cG=cG, pt_gx=model.node_coords_free.grad.clone(), pt_gu=model.u_free.grad.clone())
    Me = 33
    ee = torch.randint(0, model.N_edges, (Me,), generator=g2)
    xe = torch.rand(Me, 1, generator=g2, dtype=torch.float64).to(dtype)
    ue, ds = model(xe, ee, edge=True)
    out.update(ed_x=xe, ed_e=ee, ed_u=ue, ed_ds=ds)
    save(name, **out)


def tri_trajectory(name, dtype):
    """3 LBFGS outer steps + 10 Adam steps with the unchanged loops of examples/example4.py:54-80."""
    m = meshgen.plate_mesh(17, 9, jitter=0.2, diag="alt", seed=3, ordering="natural")
    mk = lambda: ref_models.PiecewiseLinearShapeNN2D(T(m.node_coords, dtype=dtype), T(m.connectivity),
                                                    boundary_mask=T(m.boundary_mask), dirichlet_mask=T(m.dirichlet_mask),
                                                    u_fixed=0.0, neumann_edges=T(m.neumann_edges))
    torch.manual_seed(7)
    model = mk()
    if dtype == torch.float64:
        model = model.double()
    u0 = model.u_free.detach().clone()
    loss_fn = ref_loss.EnergyLoss2D(E=10e9, nu=0.3, device=torch.device("cpu"), dtype=dtype)
    opt = torch.optim.LBFGS(model.parameters())
    lb = []
    for _ in range(3):
        def closure():
            opt.zero_grad()
            l = loss_fn(model)
            l.backward()
            return l
        lb.append(opt.step(closure).item())
    lb_final = loss_fn(model).item()
    u_lb, x_lb = model.u_free.detach().clone(), model.node_coords_free.detach().clone()
    # Adam from the same start (example4.py:54-65 learning rates)
    model2 = mk()
    if dtype == torch.float64:
        model2 = model2.double()
    with torch.no_grad():
        model2.u_free.copy_(u0)
    opt2 = torch.optim.Adam([{"params": model2.u_free, "lr": 1e-4}, {"params": model2.node_coords_free, "lr": 1e-5}], lr=1e-4)
    ad = []
    for _ in range(10):
        opt2.zero_grad()
        l = loss_fn(model2)
        l.backward()
        opt2.step()
        ad.append(l.item())
    save(name, node_coords=m.node_coords, connectivity=m.connectivity, boundary_mask=m.boundary_mask,
         dirichlet_mask=m.dirichlet_mask, neumann_edges=m.neumann_edges, u_free0=u0,
         lbfgs_losses=np.asarray(lb), lbfgs_final=lb_final, lbfgs_u=u_lb, lbfgs_x=x_lb,
         adam_losses=np.asarray(ad), adam_u=model2.u_free, adam_x=model2.node_coords_free)


# ----------------------------------------------------------------------------- 1D
def gold_1d():
    d = {}
    # lookup semantics (Q14)
    grid = torch.tensor([0.0, 1.0, 2.0, 3.0])
    x = torch.tensor([-1.0, 0.0, 0.5, 1.0, 2.999, 3.0, 4.0, 2.0, 1.0000001])
    d["lk_grid"], d["lk_x"] = grid, x
    d["lk_idx"] = (torch.searchsorted(grid, x) - 1).clamp(0, 2)
    mm = ref_models.PiecewiseLinearShapeNN(grid)
    with torch.no_grad():
        mm.u.copy_(torch.tensor([1.0, -2.0, 0.5, 3.0]))
    d["lk_u"] = mm(x)
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        for r_adapt in (False, True):
            # example1.py:25-32
            xg = torch.linspace(0, 1, 100, dtype=dt)
            xt = torch.linspace(0, 1, 1000, dtype=dt)
            ut = torch.sin(2 * torch.pi * xt)
            torch.manual_seed(0)
            model = ref_models.PiecewiseLinearShapeNN(xg, r_adapt=r_adapt)
            with torch.no_grad():
                model.u.copy_(0.3 * torch.randn(model.u.shape, dtype=torch.float64).to(dt))
                if r_adapt:
                    model.x_increments.add_(0.2 * model.x_increments * torch.randn(model.x_increments.shape, dtype=torch.float64).to(dt))
            if dt == torch.float64:
                model = model.double()
            k = f"ex1_{tag}_{'r' if r_adapt else 'f'}"
            d[k + "_u"] = model.u.detach().clone()
            if r_adapt:
                d[k + "_p"] = model.x_increments.detach().clone()
            d[k + "_grid"] = model.grid
            pred = model(xt)
            loss = ((pred - ut) ** 2).mean()
            loss.backward()
            d[k + "_pred"], d[k + "_loss"], d[k + "_gu"] = pred, loss, model.u.grad.clone()
            if r_adapt:
                d[k + "_gp"] = model.x_increments.grad.clone()
            # short Adam trajectory with the unchanged loop (example1.py:31-40)
            opt = torch.optim.Adam(model.parameters(), lr=0.005)
            tr = []
            for _ in range(25):
                opt.zero_grad()
                l = ((model(xt) - ut) ** 2).mean()
                l.backward()
                opt.step()
                tr.append(l.item())
            d[k + "_adam"] = np.asarray(tr)
            d[k + "_adam_u"] = model.u.detach().clone()
    # example3.py:74-96 bar energy, FP64 and FP32, fixed ends, r-adaptive
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        for npts, ng in ((89, 2), (300, 3)):
            xg = torch.linspace(0, 10.0, npts, dtype=dt)
            model = ref_models.PiecewiseLinearShapeNN(xg, r_adapt=True, u0=0.0, uN=0.0)
            g = torch.Generator().manual_seed(npts)
            with torch.no_grad():
                model.u.copy_(1e-2 * torch.randn(model.u.shape, generator=g, dtype=torch.float64).to(dt))
                model.x_increments.add_(0.1 * model.x_increments * torch.randn(model.x_increments.shape, generator=g, dtype=torch.float64).to(dt))
            if dt == torch.float64:
                model = model.double()
            xi, wi = ref_utils.interval_gauss_points(ng, dtype=dt)
            k = f"ex3_{tag}_{npts}"
            d[k + "_u"], d[k + "_p"] = model.u.detach().clone(), model.x_increments.detach().clone()
            d[k + "_ng"] = ng
            loss = ex3_energy(model, xi, wi, ex3_b_force, E=175.0)
            loss.backward()
            d[k + "_loss"], d[k + "_gu"], d[k + "_gp"] = loss, model.u.grad.clone(), model.x_increments.grad.clone()
            d[k + "_grid"] = model.grid
            opt = torch.optim.Adam(model.parameters(), lr=1e-4)
            tr = []
            for _ in range(15):
                opt.zero_grad()
                l = ex3_energy(model, xi, wi, ex3_b_force, E=175.0)
                l.backward()
                opt.step()
                tr.append(l.item())
            d[k + "_adam"] = np.asarray(tr)
    save("one_d", **d)


# ----------------------------------------------------------------------------- structured Q1
def gold_structured():
    d = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        for ufix, ftag in ((None, "free"), (0.25, "fix")):
            Nx, Ny = 25, 19
            gx = torch.linspace(0, 1, Nx, dtype=dt)
            gy = torch.linspace(0, 1, Ny, dtype=dt)
            torch.manual_seed(1)
            model = StructuredRef(grid_x=gx, grid_y=gy, r_adapt=True, u_fixed=ufix)
            g = torch.Generator().manual_seed(4)
            with torch.no_grad():
                model.increments_x.add_(0.2 * model.increments_x * torch.randn(Nx - 1, generator=g, dtype=torch.float64).to(dt))
                model.increments_y.add_(0.2 * model.increments_y * torch.randn(Ny - 1, generator=g, dtype=torch.float64).to(dt))
            if dt == torch.float64:
                model = model.double()
            M = 2000
            x = torch.rand(M, 2, generator=g, dtype=torch.float64).to(dt)
            x[:7] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.5, 0.5], [-0.1, 0.3], [0.3, 1.2], [1.0, 0.0], [0.25, 0.75]], dtype=dt)
            ut = torch.sin(2 * torch.pi * x[:, 0]) * torch.cos(2 * torch.pi * x[:, 1])
            k = f"q1_{tag}_{ftag}"
            d[k + "_px"], d[k + "_py"], d[k + "_u"] = model.increments_x.detach().clone(), model.increments_y.detach().clone(), model.u.detach().clone()
            gxx, gyy = model.grid
            d[k + "_gx"], d[k + "_gy"] = gxx, gyy
            d[k + "_x"], d[k + "_ut"] = x, ut
            pred = model(x)
            loss = ((pred - ut) ** 2).mean()
            loss.backward()
            d[k + "_pred"], d[k + "_loss"] = pred, loss
            d[k + "_gu"], d[k + "_gpx"], d[k + "_gpy"] = model.u.grad.clone(), model.increments_x.grad.clone(), model.increments_y.grad.clone()
            d[k + "_ix"] = (torch.searchsorted(gxx.detach(), x[:, 0].contiguous()) - 1).clamp(0, Nx - 2)
            d[k + "_iy"] = (torch.searchsorted(gyy.detach(), x[:, 1].contiguous()) - 1).clamp(0, Ny - 2)
            if ufix is not None:
                d[k + "_ufix"] = ufix
            opt = torch.optim.Adam(model.parameters(), lr=0.005)
            tr = []
            for _ in range(15):
                opt.zero_grad()
                l = ((model(x) - ut) ** 2).mean()
                l.backward()
                opt.step()
                tr.append(l.item())
            d[k + "_adam"] = np.asarray(tr)
    save("structured", **d)


if __name__ == "__main__":
    torch.set_num_threads(1)       # deterministic reduction order for the fixtures
    gold_quadrature()
    tri_case("tri_f64_jitter", 13, 8, torch.float64, "random", 0.25, 0.0, 1.0e3, True)
    tri_case("tri_f64_inverted", 11, 7, torch.float64, "morton", 0.3, 0.5, 1.0e3, True, gauss_order=7, g1=3, seed=1)
    tri_case("tri_f32_jitter", 13, 8, torch.float32, "random", 0.25, 0.0, 1.0e3, True)
    tri_case("tri_f64_default", 21, 11, torch.float64, "natural", 0.0, 0.0, 1.0, False, seed=2)
    tri_case("tri_f64_order3", 9, 6, torch.float64, "natural", 0.1, 0.0, 1.0e3, True, gauss_order=3, g1=1, seed=4,
             u_fixed=0.0)
    tri_trajectory("tri_traj_f64", torch.float64)
    tri_trajectory("tri_traj_f32", torch.float32)
    gold_1d()
    gold_structured()
