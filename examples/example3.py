"""1D bar under body force -- the loop of /root/reference/examples/example3.py:74-98 on the drop-in class.
`--generic` uses the reference's own energy_loss formulation (double backward through autograd.grad);
the default is the fused kernel (bar_energy_loss)."""
import sys

import torch
import torch.optim as optim

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200.models import PiecewiseLinearShapeNN
from hidenn_fem_b200.models_grid import bar_energy_loss, energy_loss_generic, example3_b_force as b_force
from hidenn_fem_b200.utils import interval_gauss_points

L, E, u0, uN = 10.0, 175.0, 0.0, 0.0
grid_pts, n_gauss, r_adapt = 89, 2, True
device = torch.device("cuda")
x_grid = torch.linspace(0, L, grid_pts).to(device)
xi, wi = interval_gauss_points(n_gauss, device=device)
energy_loss = energy_loss_generic if "--generic" in sys.argv else bar_energy_loss

model = PiecewiseLinearShapeNN(x_grid, r_adapt=r_adapt, u0=u0, uN=uN).to(device)
optimizer = optim.Adam(model.parameters(), lr=1e-4)
for epoch in range(4000):
    optimizer.zero_grad()
    loss = energy_loss(model, xi, wi, b_force, E=E)
    loss.backward()
    optimizer.step()
    if epoch % 500 == 0:
        print(f"Epoch {epoch}: loss={loss.item():.6f}")
