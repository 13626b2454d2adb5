"""2D plate under traction -- the loop of /root/reference/examples/example4.py:68-80 (LBFGS) and :54-65 (Adam),
unchanged, on the B200 drop-in classes.  The gmsh mesh is replaced by the seeded synthetic plate generator.

    python -m examples.example4 [--nx 401 --ny 201] [--adam]
"""
import argparse

import numpy as np
import torch

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200 import meshgen
from hidenn_fem_b200.models import PiecewiseLinearShapeNN2D
from hidenn_fem_b200.loss import EnergyLoss2D
from hidenn_fem_b200.utils import test_gradients

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=201)
ap.add_argument("--ny", type=int, default=101)
ap.add_argument("--adam", action="store_true")
ap.add_argument("--double", action="store_true")
args = ap.parse_args()

device = torch.device("cuda")
dtype = torch.float64 if args.double else torch.float32
length, height = 2.0, 1.0
m = meshgen.plate_mesh(args.nx, args.ny, length, height, jitter=0.2, diag="random", ordering="random")
T = torch.tensor
node_coords, connectivity = T(m.node_coords, dtype=torch.float32), T(m.connectivity)
geom_boundary_mask, bc_mask, mn_mask, neumann_edges = T(m.boundary_mask), T(m.dirichlet_mask), T(m.neumann_mask), T(m.neumann_edges)
print("Nodes:", node_coords.shape)
print("Connectivity:", connectivity.shape)
print("Geometric boundary nodes:", geom_boundary_mask.sum().item())
print("Dirichlet BC nodes:", bc_mask.sum().item())
print("Neumann MN nodes:", mn_mask.sum().item())
print("Neumann edges:", neumann_edges.shape)

model = PiecewiseLinearShapeNN2D(node_coords, connectivity, boundary_mask=geom_boundary_mask, dirichlet_mask=bc_mask,
                                 u_fixed=0.0, neumann_edges=neumann_edges)
if args.double:
    model = model.double()
model = model.to(device)
loss_fn = EnergyLoss2D(E=10e9, nu=0.3, length=length, height=height, device=device, dtype=dtype)
test_gradients(model, loss_fn)
model.zero_grad()

if args.adam:
    optimizer = torch.optim.Adam([{"params": model.u_free, "lr": 1e-4}, {"params": model.node_coords_free, "lr": 1e-5}], lr=1e-4)
    for epoch in range(2000):
        optimizer.zero_grad()
        loss = loss_fn(model)
        loss.backward()
        optimizer.step()
        if epoch % 200 == 0:
            print(f"Epoch {epoch}: Loss = {loss.item():.6e}")
else:
    optimizer = torch.optim.LBFGS(model.parameters())
    for epoch in range(30):
        def closure():
            optimizer.zero_grad()
            loss = loss_fn(model)
            loss.backward()
            return loss
        loss = optimizer.step(closure)
        if epoch % 5 == 0:
            print(f"Epoch {epoch:04d}: Loss = {loss.item():.6e}")

print("Training finished.")
u_vals = model.u_full.cpu().detach().numpy()
print("Nodal values u", u_vals.shape)
print("Nodal values u_x:", np.mean(u_vals[:, 0]), np.min(u_vals[:, 0]), np.max(u_vals[:, 0]))
print("Nodal values u_y:", np.mean(u_vals[:, 1]), np.min(u_vals[:, 1]), np.max(u_vals[:, 1]))
# von Mises at centroids: the second caller of forward(x_ref, elem_id) in the reference (src/plots.py:183-201)
n_elem = model.Nelems
x_eval = torch.tensor([[1 / 3, 1 / 3]], dtype=model.dtype, device=model.device).expand(n_elem, 2)
_, _, grad_u = model(x_eval, torch.arange(n_elem, device=model.device))
g = grad_u.detach().cpu().numpy()
exx, eyy, exy = g[:, 0, 0], g[:, 1, 1], 0.5 * (g[:, 0, 1] + g[:, 1, 0])
E_, nu_ = 10e9, 0.3
sxx, syy, sxy = E_ / (1 - nu_ ** 2) * (exx + nu_ * eyy), E_ / (1 - nu_ ** 2) * (eyy + nu_ * exx), E_ / (1 + nu_) * exy
print("max von Mises stress [Pa]:", float(np.sqrt(sxx ** 2 - sxx * syy + syy ** 2 + 3 * sxy ** 2).max()))
