"""1D L2 projection of sin(2 pi x) onto an r-adaptive piecewise-linear model on the B200 path.

Workload of the reference's first example (/root/reference/examples/example1.py:25-42: 100 nodes, 1000 samples, Adam
with lr 5e-3 for 500 steps, FP32); only the model class comes from this package.  Ends with the maximum nodal error
and the range of element sizes the r-adaptation produced."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200.models import PiecewiseLinearShapeNN


def target(x):
    return torch.sin(2.0 * math.pi * x)


def fit(n_nodes=100, n_samples=1000, steps=500, lr=5e-3, report_every=100, device="cuda"):
    dev = torch.device(device)
    nodes = torch.linspace(0.0, 1.0, n_nodes, device=dev)
    samples = torch.linspace(0.0, 1.0, n_samples, device=dev)
    wanted = target(samples)
    net = PiecewiseLinearShapeNN(nodes, r_adapt=True).to(dev)
    adam = torch.optim.Adam(net.parameters(), lr=lr)
    for it in range(steps):
        adam.zero_grad()
        mse = torch.mean(torch.square(net(samples) - wanted))
        mse.backward()
        adam.step()
        if it % report_every == 0:
            print(f"Epoch {it}: loss={mse.item():.6f}")
    with torch.no_grad():
        worst = (net(samples) - wanted).abs().max().item()
        sizes = net.grid.diff()
    print(f"max |u_h - u| on the samples: {worst:.3e}; element sizes {sizes.min().item():.5f} .. {sizes.max().item():.5f}")
    return net


if __name__ == "__main__":
    fit()
