"""Launch the dominant tile kernel of the C4 bench workload a few times (for ncu captures):
    ncu --set full --clock-control none --import-source on -k regex:tri_tile -s 2 -c 1 -o gpurun_out/X python profiles/ncu_tile.py [--ordering tiles] [--dtype f64]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--elems", type=int, default=10_000_000)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--ordering", default="tiles")
ap.add_argument("--tile-nodes", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.float64 if a.dtype == "f64" else torch.float32
m, model, loss_fn, _ = bench.make_workload(a, 0, 1, dev, dt, a.ordering, a.elems, a.tile_nodes)
print("plan", {k: model._plan().info[k] for k in ("n_tiles", "tile_ordered", "max_local", "max_entries")})
print("kernel ms", bench.time_kernel(model, loss_fn, 4, 2))
