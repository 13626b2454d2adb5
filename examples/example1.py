"""1D L2 projection -- the loop of /root/reference/examples/example1.py:25-42, unchanged, on the drop-in class."""
import torch
import torch.optim as optim

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200.models import PiecewiseLinearShapeNN

device = torch.device("cuda")
x_grid = torch.linspace(0, 1, 100).to(device)
x_train = torch.linspace(0, 1, 1000).to(device)
u_true = torch.sin(2 * torch.pi * x_train)

model = PiecewiseLinearShapeNN(x_grid, r_adapt=True).to(device)
optimizer = optim.Adam(model.parameters(), lr=0.005)
for epoch in range(500):
    optimizer.zero_grad()
    pred = model(x_train)
    loss = ((pred - u_true) ** 2).mean()
    loss.backward()
    optimizer.step()
    if epoch % 100 == 0:
        print(f"Epoch {epoch}: loss={loss.item():.6f}")
