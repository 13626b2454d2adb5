"""Plate with holes on N GPUs -- the LBFGS loop of /root/reference/examples/example4.py:68-80 on a partitioned mesh.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/example4_distributed.py

Every rank builds its column strip of the same global plate (bit-identical to that part of the global mesh), the halo
plan finds the nodes shared with the neighbouring strips, `DistributedEnergyLoss2D` completes their gradients with one
packed all-reduce per evaluation, and `ShardedLBFGS` runs the reference's optimiser with global inner products."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from anywhere
from hidenn_fem_b200 import dist as hd, meshgen
from hidenn_fem_b200.models import PiecewiseLinearShapeNN2D
from hidenn_fem_b200.optim import ShardedLBFGS

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
dtype = torch.float64

nx, ny = meshgen.plate_dims_for_elements(200_000)
mesh = hd.strip_mesh(nx, ny, rank, world, jitter=0.2, diag="random", seed=0, ordering="morton")
T = torch.tensor
model = PiecewiseLinearShapeNN2D(T(mesh.node_coords, dtype=dtype), T(mesh.connectivity), T(mesh.boundary_mask),
                                 T(mesh.dirichlet_mask), 0.0, T(mesh.neumann_edges)).double().to(device)
with torch.no_grad():
    model.u_free.zero_()                       # same start on every copy of a shared node
halo = hd.setup_strip_halo(mesh, mesh.boundary_mask, mesh.dirichlet_mask, device, dtype)
loss_fn = hd.DistributedEnergyLoss2D(E=10e9, nu=0.3, length=2.0, height=1.0, device=device, dtype=dtype, halo=halo)

model.node_coords_free.requires_grad_(False)   # solve for the displacements on the fixed mesh (example4.py:92-103 style)
optimizer = ShardedLBFGS([model.u_free], weights=[halo.row_weights[1]])
for epoch in range(6):
    def closure():
        optimizer.zero_grad()
        loss = loss_fn(model)
        loss.backward()
        return loss
    loss = optimizer.step(closure)
    if rank == 0:
        print(f"Epoch {epoch:04d}: Loss = {loss.item():.6e}")

umax = model.u_free.detach().abs().max().reshape(1)
dist.all_reduce(umax, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"{world} ranks, {halo.S} shared nodes, max |u| = {umax.item():.4e}")
dist.destroy_process_group()
