"""Per-phase cycle breakdown of the tile kernel (clock64 inside the kernel, hidenn_debug_tile_timing).
    python profiles/phase_timing.py [--elems N] [--dtype f64|f32] [--tile-nodes T] [--ordering morton]"""
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
buf = torch.zeros(plan.info["n_tiles"] * 8, dtype=torch.int64, device=dev)
L = _lib.lib()
L.hidenn_debug_tile_timing(C.c_void_p(buf.data_ptr()))
ms = bench.time_kernel(model, loss_fn, 5, 2)
L.hidenn_debug_tile_timing(C.c_void_p(0))
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 8)
p1, p2, p3 = t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2]
tot = t[:, 3] - t[:, 0]
print(f"kernel {ms*1e3:.1f} us, tiles {len(t)}, elems/tile {plan.info['elem_visits']/len(t):.0f}")
pa, pb, pc, pd = t[:, 6] - t[:, 0], t[:, 1] - t[:, 6], t[:, 5] - t[:, 2], t[:, 3] - t[:, 5]
for name, v in (("stage (loads->smem)", p1), ("  .. desc+slots arrive", pa), ("  .. gathers+barrier", pb), ("elements", p2), ("fold+store", p3),
                ("  .. fold loops+stores (thread 0)", pc), ("  .. prefetch+reduce+barrier", pd), ("CTA total", tot)):
    print(f"  {name:34s} mean {v.mean():8.0f} cyc  p10 {np.percentile(v,10):8.0f}  p50 {np.percentile(v,50):8.0f}  p90 {np.percentile(v,90):8.0f}  ({100*v.mean()/tot.mean():4.1f}%)")
# concurrency: per SM, sum of CTA lifetimes / wall span
sm = t[:, 4]
occ = []
for s_ in np.unique(sm)[:8]:
    sel = sm == s_
    span = t[sel, 3].max() - t[sel, 0].min()
    occ.append(tot[sel].sum() / span)
print("  mean resident CTAs per SM (first 8 SMs):", np.round(occ, 2))
