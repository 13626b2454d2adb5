"""2D L2 projection on a structured grid -- the loop of /root/reference/examples/example2.py:13-50 on the drop-in
class (the reference script itself raises TypeError because its structured class is shadowed)."""
import torch
import torch.optim as optim

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200.models import PiecewiseLinearShapeNN2D

device = torch.device("cuda")
Nx, Ny = 25, 25
grid_x = torch.linspace(0, 1, Nx, device=device)
grid_y = torch.linspace(0, 1, Ny, device=device)
nx_train, ny_train, M = 100, 100, 1000
XX, YY = torch.meshgrid(torch.linspace(0, 1, nx_train, device=device), torch.linspace(0, 1, ny_train, device=device), indexing="ij")
x_train = torch.stack([XX.flatten(), YY.flatten()], dim=1)
u_true = torch.sin(2 * torch.pi * x_train[:, 0]) * torch.cos(2 * torch.pi * x_train[:, 1])

model = PiecewiseLinearShapeNN2D(grid_x=grid_x, grid_y=grid_y, boundary_mask_x=None, boundary_mask_y=None, r_adapt=True).to(device)
optimizer = optim.Adam(model.parameters(), lr=0.005)
for epoch in range(5000):
    optimizer.zero_grad()
    indices = torch.randint(0, x_train.shape[0], (M,), device=device)
    pred = model(x_train[indices])
    loss = ((pred - u_true[indices]) ** 2).mean()
    loss.backward()
    optimizer.step()
    if epoch % 500 == 0:
        print(f"Epoch {epoch}: loss={loss.item():.6f}")
