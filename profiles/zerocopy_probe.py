"""Probe: run the resident entry point with the Parameter rows / gradient rows living in PINNED HOST memory (UVA zero-copy)."""
import ctypes as C, time, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hidenn_fem_b200 import _lib

class A: pass
args = A(); args.elems = 10_000_000; args.tile_nodes = 0
dev = torch.device("cuda:0")
for ordering in ("morton", "natural"):
    m, model, loss_fn, _ = bench.make_workload(args, 0, 1, dev, torch.float64, ordering, args.elems, 0)
    loss = loss_fn(model); loss.backward(); torch.cuda.synchronize()
    plan = model._plan()
    consts, hints = loss_fn._consts(model, None)
    xb, ub = model._fixed_pair()
    xf = model.node_coords_free.detach().cpu().pin_memory(); uf = model.u_free.detach().cpu().pin_memory()
    gx = torch.empty_like(xf).pin_memory(); gu = torch.empty_like(uf).pin_memory()
    out = torch.empty(4, dtype=torch.float64, device=dev)
    scratch = torch.zeros(plan.info["scratch"], dtype=torch.float64, device=dev)
    f = _lib.fn("hidenn_tri_energy", torch.float64)
    s = _lib.stream_ptr()
    def call(xin, uin, gxo, guo):
        _lib.check(f(plan.handle, _lib.ptr(xin), _lib.ptr(xb), _lib.ptr(uin), _lib.ptr(ub), _lib.ptr(consts), None, C.c_int(7 | hints),
                     _lib.ptr(out), _lib.ptr(gxo), _lib.ptr(guo), None, _lib.ptr(scratch), s))
        torch.cuda.synchronize()
    dgx = torch.empty_like(model.node_coords_free); dgu = torch.empty_like(model.u_free)
    for name, a in (("all zero-copy", (xf, uf, gx, gu)), ("reads zero-copy, writes HBM", (xf, uf, dgx, dgu)),
                    ("reads HBM, writes zero-copy", (model.node_coords_free.detach(), model.u_free.detach(), gx, gu))):
        for _ in range(2): call(*a)
        t0 = time.perf_counter()
        for _ in range(5): call(*a)
        t = (time.perf_counter() - t0) / 5
        ok = ""
        if a[2] is gx:
            ok = " bit-equal=%s" % (torch.equal(gx, model.node_coords_free.grad.cpu()) and torch.equal(gu, model.u_free.grad.cpu()))
        print(ordering, name, "%.3f ms" % (t * 1e3), "loss", out[0].item() == loss.item(), ok, flush=True)
    del model, loss_fn, plan
    torch.cuda.empty_cache()
