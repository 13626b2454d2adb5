"""C3 (4097^2 nodes, 2^26 samples, FP64) step times: reference expression vs fused l2_projection_loss."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
out = bench.bench_grid_paths(torch.device("cuda:0"), 20, 5, 6546.2, full_c3=True)
for k, v in out.items():
    print(k, "%.3f ms/step" % v["ms_per_step"], "%.3e evals/s" % v["evals_per_s"])
