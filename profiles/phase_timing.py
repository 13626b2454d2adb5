"""Per-phase cycle breakdown of the persistent tile kernel (clock64 inside the kernel, hidenn_debug_tile_timing).
    python profiles/phase_timing.py [--elems N] [--dtype f64|f32] [--tile-nodes T] [--ordering morton]
Layout per tile (16 int64): thread 0 -> [0..7], thread 255 -> [8..15]:
  0 tile start (after the barrier), 1 elements done (before barrier), 2 after barrier, 3 fold+stores done, 4 smid,
  5 after rotate, 6 after cp.async.wait_all, 7 after the closing barrier."""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hidenn_fem_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--elems", type=int, default=10_000_000)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--tile-nodes", type=int, default=0)
ap.add_argument("--ordering", default="morton")
a = ap.parse_args()
dev = torch.device("cuda:0")
dt = torch.float64 if a.dtype == "f64" else torch.float32
m, model, loss_fn, _ = bench.make_workload(a, 0, 1, dev, dt, a.ordering, a.elems, a.tile_nodes)
plan = model._plan()
for _ in range(3):
    model.zero_grad(); loss_fn(model).backward()
buf = torch.zeros(plan.info["n_tiles"] * 16, dtype=torch.int64, device=dev)
L = _lib.lib()
L.hidenn_debug_tile_timing(C.c_void_p(buf.data_ptr()))
ms = bench.time_kernel(model, loss_fn, 5, 2)
L.hidenn_debug_tile_timing(C.c_void_p(0))
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 16)
t = t[(t[:, 7] > 0) & (t[:, 15] > 0)]          # tiles that are not the last of their CTA
print(f"kernel {ms*1e3:.1f} us, tiles {plan.info['n_tiles']}, elems/tile {plan.info['elem_visits']/plan.info['n_tiles']:.0f}")
def row(name, v, tot):
    print(f"  {name:44s} mean {v.mean():8.0f} cyc  p10 {np.percentile(v,10):8.0f}  p50 {np.percentile(v,50):8.0f}  p90 {np.percentile(v,90):8.0f}  ({100*v.mean()/tot:4.1f}%)")
for w, o in (("warp 0", 0), ("warp 7", 8)):
    tot = (t[:, o + 7] - t[:, o + 0]).mean()
    print(f" {w}: per-tile loop {tot:.0f} cycles")
    row("E: gathers issue + elements", t[:, o + 1] - t[:, o + 0], tot)
    row("   wait at barrier B2", t[:, o + 2] - t[:, o + 1], tot)
    row("F: metadata loads + fold + stores", t[:, o + 3] - t[:, o + 2], tot)
    row("   rotate (waits for the metadata loads)", t[:, o + 5] - t[:, o + 3], tot)
    row("   cp.async.wait_all", t[:, o + 6] - t[:, o + 5], tot)
    row("   wait at barrier B1", t[:, o + 7] - t[:, o + 6], tot)
