"""Multi-GPU parity check (run under torchrun on N GPUs of one box, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check_gpu.py [--elems 400000]

Each rank builds its column strip, runs the fused kernels + one packed NCCL all-reduce, and compares loss and
gradients with a single-GPU evaluation of the global mesh done on the same device (partition invariance,
SURVEY §4 (v)).  Also runs 5 Adam steps on both and checks the halo copies stay consistent."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)      # a hung collective must not hold the box
    ap = argparse.ArgumentParser()
    ap.add_argument("--elems", type=int, default=400_000)
    ap.add_argument("--dtype", default="f64")
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dt = torch.float64 if a.dtype == "f64" else torch.float32
    tol = 1e-10 if dt == torch.float64 else 1e-5

    class A:  # the knobs make_workload reads
        pass
    m, model, loss_fn, dims = bench.make_workload(A, rank, world, dev, dt, "morton", a.elems // world)
    mg, gmodel, gloss_fn, _ = bench.make_workload(A, 0, 1, dev, dt, "morton", a.elems)
    assert dims == _, (dims, _)

    def evaluate(mod, lf):
        mod.zero_grad(set_to_none=True)
        l = lf(mod)
        l.backward()
        return l.item(), mod.node_coords_free.grad.clone(), mod.u_free.grad.clone()

    l, gx, gu = evaluate(model, loss_fn)
    L, GX, GU = evaluate(gmodel, gloss_fn)
    # map local free rows to global free rows through the generating-grid ids
    pos = {int(g): i for i, g in enumerate(mg.global_node_id)}
    loc = np.array([pos[int(g)] for g in m.global_node_id])
    gfx = np.cumsum(~mg.boundary_mask) - 1
    gfu = np.cumsum(~mg.dirichlet_mask) - 1
    fx, fu = ~m.boundary_mask, ~m.dirichlet_mask
    assert np.array_equal(fx, ~mg.boundary_mask[loc]) and np.array_equal(fu, ~mg.dirichlet_mask[loc])
    rx = torch.from_numpy(gfx[loc][fx]).to(dev)
    ru = torch.from_numpy(gfu[loc][fu]).to(dev)
    rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
    el, ex, eu = abs(l - L) / abs(L), rel(gx, GX[rx]), rel(gu, GU[ru])
    ok = el < tol and ex < tol and eu < tol
    # a few Adam steps with the unchanged loop: halo copies must follow the global trajectory
    o1 = torch.optim.Adam([{"params": model.u_free, "lr": 1e-4}, {"params": model.node_coords_free, "lr": 1e-5}])
    o2 = torch.optim.Adam([{"params": gmodel.u_free, "lr": 1e-4}, {"params": gmodel.node_coords_free, "lr": 1e-5}])
    for _ in range(5):
        for mod, lf, o in ((model, loss_fn, o1), (gmodel, gloss_fn, o2)):
            o.zero_grad()
            lf(mod).backward()
            o.step()
    eu2 = rel(model.u_free.detach(), gmodel.u_free.detach()[ru])
    ex2 = rel(model.node_coords_free.detach(), gmodel.node_coords_free.detach()[rx])
    ok = ok and eu2 < (1e-7 if dt == torch.float64 else 1e-4) and ex2 < (1e-9 if dt == torch.float64 else 1e-5)
    # CUDA-graph replay of the rank-local step + eager halo exchange: same bits as the eager step, twice in a row
    from hidenn_fem_b200.graph import GraphedEnergyStep
    l_e, gx_e, gu_e = evaluate(model, loss_fn)
    step = GraphedEnergyStep(model, loss_fn)
    for _ in range(2):
        l_g = step()
        ok = ok and l_g.item() == l_e and torch.equal(model.node_coords_free.grad, gx_e) and torch.equal(model.u_free.grad, gu_e)
    # sharded L-BFGS (global inner products through the owner weights) vs the stock optimiser on one GPU
    from hidenn_fem_b200.optim import ShardedLBFGS
    ob = ShardedLBFGS(model.parameters(), max_iter=6, history_size=6, weights=loss_fn.halo.row_weights)
    og = torch.optim.LBFGS(gmodel.parameters(), max_iter=6, history_size=6)
    lb, lg = [], []
    for _ in range(2):
        def c1():
            ob.zero_grad(); l = loss_fn(model); l.backward(); return l
        def c2():
            og.zero_grad(); l = gloss_fn(gmodel); l.backward(); return l
        lb.append(float(ob.step(c1).detach())); lg.append(float(og.step(c2).detach()))
    eu3 = rel(model.u_free.detach(), gmodel.u_free.detach()[ru])
    el3 = max(abs(a - b) / abs(b) for a, b in zip(lb, lg))
    # (lr = 1 without a line search is not a contraction on this problem -- the reference's own setting -- so rounding
    # differences between the two summation orders grow along the trajectory: 4e-16 at N=2, 3e-6 at N=8 after 12 iterations)
    ok = ok and eu3 < (1e-4 if dt == torch.float64 else 1e-2) and el3 < (1e-8 if dt == torch.float64 else 1e-3)
    if rank == 0:
        print("sharded LBFGS vs single-GPU LBFGS: losses %s vs %s (rel %.2e), u rel %.2e" % (lb, lg, el3, eu3))
    res = torch.tensor([el, ex, eu, eu2, ex2, 0.0 if ok else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("dist_check world=%d elems=%d shared_nodes=%d: max rel err loss %.2e gx %.2e gu %.2e | after 5 Adam steps u %.2e x %.2e -> %s"
              % (world, mg.connectivity.shape[0], loss_fn.halo.S, *res[:5].tolist(), "OK" if res[5].item() == 0 else "FAIL"))
    dist.destroy_process_group()
    sys.exit(0 if res[5].item() == 0 else 1)


if __name__ == "__main__":
    main()
