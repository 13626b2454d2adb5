"""One C2 step (1D bar, 1 M elements, FP64, fused bar step) a few times -- for launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/c2.csv python profiles/c2_step.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hidenn_fem_b200.models import PiecewiseLinearShapeNN
from hidenn_fem_b200 import models_grid as mg
from hidenn_fem_b200.utils import interval_gauss_points
N = 1_000_001
dev = torch.device("cuda:0")
m1 = PiecewiseLinearShapeNN(torch.linspace(0, 10.0, N, dtype=torch.float64), r_adapt=True, u0=0.0, uN=0.0).double().to(dev)
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    m1.u.copy_((1e-2 * torch.randn(N - 2, generator=g, dtype=torch.float64)).to(dev))
xi, wi = interval_gauss_points(2, device=dev, dtype=torch.float64)
for _ in range(4):
    m1.zero_grad(set_to_none=True)
    mg.bar_energy_loss(m1, xi, wi, None, 175.0, b_builtin=True).backward()
torch.cuda.synchronize()
mg._bar_state.check(block=True)
print("ok")
