"""Host-vs-device time of one training step (torch.profiler); run under torchrun for N>1.
    python profiles/host_profile.py [--elems N]"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import torch.distributed as dist
ap = argparse.ArgumentParser(); ap.add_argument("--elems", type=int, default=10_000_000); ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN"); dist.init_process_group("nccl", device_id=dev)
dt = torch.float64 if a.dtype == "f64" else torch.float32
m, model, loss_fn, _ = bench.make_workload(a, rank, world, dev, dt, "morton", a.elems)
def step():
    model.zero_grad(set_to_none=True); l = loss_fn(model); l.backward(); return l
for _ in range(10): step()
torch.cuda.synchronize()
# pure host time: enqueue 100 steps without waiting (the GPU queue absorbs them), measure enqueue rate
t0 = time.perf_counter()
for _ in range(100): step()
t_enq = (time.perf_counter() - t0) / 100
torch.cuda.synchronize()
t_tot = (time.perf_counter() - t0) / 100
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(20): step()
    torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}: enqueue {t_enq*1e6:.0f} us/step (host), wall {t_tot*1e6:.0f} us/step")
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=44))
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
if world > 1: dist.destroy_process_group()
