"""2D L2 projection of sin(2 pi x) cos(2 pi y) onto the structured (tensor-product) r-adaptive model on the B200 path.

Workload of the reference's second example (/root/reference/examples/example2.py:13-50: 25 x 25 nodes, a 100 x 100
sample lattice, 1000 random samples per step, Adam with lr 5e-3, FP32).  The reference script itself stops with a
TypeError because its structured class is shadowed by the triangle class of the same name; here the
structured model is selected by the `grid_x` / `grid_y` keywords.  `models_grid.l2_projection_loss(net, x, u)` is the
fused form of the loss expression used below."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200.models import PiecewiseLinearShapeNN2D


def target(p):
    return torch.sin(2.0 * math.pi * p[:, 0]) * torch.cos(2.0 * math.pi * p[:, 1])


def sample_lattice(n, dev):
    line = torch.linspace(0.0, 1.0, n, device=dev)
    a, b = torch.meshgrid(line, line, indexing="ij")
    return torch.stack([a.reshape(-1), b.reshape(-1)], dim=1)


def fit(nodes_per_side=25, lattice=100, batch=1000, steps=5000, lr=5e-3, report_every=500, device="cuda"):
    dev = torch.device(device)
    axis = torch.linspace(0.0, 1.0, nodes_per_side, device=dev)
    points = sample_lattice(lattice, dev)
    wanted = target(points)
    net = PiecewiseLinearShapeNN2D(grid_x=axis, grid_y=axis.clone(), r_adapt=True).to(dev)
    adam = torch.optim.Adam(net.parameters(), lr=lr)
    for it in range(steps):
        adam.zero_grad()
        pick = torch.randint(0, points.shape[0], (batch,), device=dev)
        mse = torch.mean(torch.square(net(points[pick]) - wanted[pick]))
        mse.backward()
        adam.step()
        if it % report_every == 0:
            print(f"Epoch {it}: loss={mse.item():.6f}")
    with torch.no_grad():
        full = torch.mean(torch.square(net(points) - wanted)).item()
    print(f"mean squared error on the full {lattice} x {lattice} lattice: {full:.3e}")
    return net


if __name__ == "__main__":
    fit()
