"""torch-profiler breakdown of the fused C3 step (models_grid.l2_projection_loss; 4097^2 nodes, 2^26 meshgrid samples, FP64)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hidenn_fem_b200.models import StructuredShapeNN2D
from hidenn_fem_b200.models_grid import l2_projection_loss
dev = torch.device("cuda:0")
Ng, Ms = 4097, 8192
gx = torch.linspace(0, 1, Ng, dtype=torch.float64)
m2 = StructuredShapeNN2D(gx, gx.clone(), r_adapt=True).double().to(dev)
xs = torch.linspace(0, 1, Ms, dtype=torch.float64, device=dev)
XX, YY = torch.meshgrid(xs, xs, indexing="ij")
x_train = torch.stack([XX.flatten(), YY.flatten()], dim=1); del XX, YY
u_true = torch.sin(2 * torch.pi * x_train[:, 0]) * torch.cos(2 * torch.pi * x_train[:, 1])
def step():
    m2.zero_grad(set_to_none=True); l2_projection_loss(m2, x_train, u_true).backward()
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
