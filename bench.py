#!/usr/bin/env python
"""Benchmark of the HiDeNN-FEM quadrature hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f64|f32]

A "step" = one evaluation of EnergyLoss2D (domain + Neumann edges) forward AND backward
(r-adaptive: gradients w.r.t. nodal values and nodal coordinates) over the whole mesh, through the
drop-in Python API with the reference's own loop body (zero_grad -> loss_fn(model) -> backward).
Workload C4: 2x1 plate with three holes, >= 10 M unstructured triangles per GPU
(jittered nodes, hashed diagonals), gauss_order=4 (ng=4), FP64.  N>1: the global plate has N x 10 M
elements split into column strips (contiguous element blocks + halo nodes), one packed all-reduce
per step ("weak" scaling).

Prints ONE JSON line (rank 0).  metric = element-quadrature evals/s = Ne*ng/t_step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

NG = 4
METRIC = "element-quadrature evals/s (fwd+bwd)"
WORKLOAD = ("C4 examples/example4.py 2D plate linear elasticity, EnergyLoss2D energy+gradient (fwd+bwd, r-adaptive), "
            "unstructured triangles (jitter 0.25, hashed diagonals), gauss_order=4")
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--elems", type=int, default=10_000_000, help="elements per GPU")
    ap.add_argument("--ordering", default="tiles", choices=["tiles", "natural", "morton", "random"],
                    help="node numbering of the synthetic mesh; 'tiles' = passed through meshgen.reorder_for_locality "
                         "(the package's mesh-ingestion helper), which FP64 plans recognise and stage with bulk copies")
    ap.add_argument("--tile-nodes", type=int, default=0)
    ap.add_argument("--cpu-sample-elems", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph replay of the step")
    ap.add_argument("--watchdog-s", type=int, default=1500, help="dump all thread stacks and exit if the run takes longer")
    ap.add_argument("--extra", action="store_true", help="also time the random-ordering and FP32 variants (N=1)")
    ap.add_argument("--no-grid", action="store_true", help="skip the C2 / C3 lines (other_configs) of the default run")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 1 s back-to-back leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --elems per GPU (driver default); strong: --elems-total over all GPUs (C5 as north_star words it)")
    ap.add_argument("--elems-total", type=int, default=80_000_000, help="total elements of the strong-scaling run")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:       # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------------------------------------
def balanced_splits(nx, ny, world, holes, length=2.0, height=1.0):
    """Column split points giving every rank about the same number of kept elements."""
    if world == 1:
        return [0, nx - 1]
    hx, hy = length / (nx - 1), height / (ny - 1)
    y = np.arange(ny) * hy
    cnt = np.zeros(nx - 1, np.int64)
    prev = None
    for ix in range(nx):
        x = ix * hx
        ins = np.zeros(ny, bool)
        for (cx, cy, r) in holes:
            ins |= (x - cx) ** 2 + (y - cy) ** 2 <= r * r
        if prev is not None:
            bad = prev[:-1] | prev[1:] | ins[:-1] | ins[1:]      # cell touches an inside node -> at most 2 elements lost
            cnt[ix - 1] = 2 * (ny - 1) - 2 * int(bad.sum())
        prev = ins
    cum = np.concatenate([[0], np.cumsum(cnt)])
    splits = [0]
    for r in range(1, world):
        splits.append(int(np.searchsorted(cum, cum[-1] * r / world)))
    splits.append(nx - 1)
    return splits


def make_workload(args, rank, world, device, dtype, ordering, n_elems_per_gpu, tile_nodes=0):
    from hidenn_fem_b200 import meshgen, dist as hd
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN2D
    from hidenn_fem_b200.loss import EnergyLoss2D
    nx, ny = meshgen.plate_dims_for_elements(n_elems_per_gpu * world)
    splits = balanced_splits(nx, ny, world, meshgen.DEFAULT_HOLES)
    # "tiles": generator output (Z-curve numbered, cheap to build) passed through the ingestion helper;
    # "random+reorder": the same helper on a randomly numbered mesh (what a gmsh mesh looks like)
    # "tiles+rotated": every element's corners rotated at random (an arbitrary mesher's corner order: all nine pair classes
    # occur, so the plan keeps one element per entry -- DESIGN.md 3.1)
    gen_order = {"tiles": "morton", "random+reorder": "random", "tiles+rotated": "morton"}.get(ordering, ordering)
    m = meshgen.plate_mesh(nx, ny, jitter=0.25, diag="random", seed=0, ordering=gen_order,
                           col_range=(splits[rank], splits[rank + 1]))
    if ordering == "tiles+rotated":
        rot = np.random.default_rng(0).integers(0, 3, m.connectivity.shape[0])
        m.connectivity = np.take_along_axis(m.connectivity, (np.arange(3)[None, :] + rot[:, None]) % 3, axis=1)
    if ordering in ("tiles", "random+reorder", "tiles+rotated"):
        m = meshgen.reorder_mesh(m, mode="tiles", tile_nodes=tile_nodes)
    T = torch.tensor
    torch.manual_seed(0)
    model = PiecewiseLinearShapeNN2D(T(m.node_coords, dtype=dtype), T(m.connectivity), T(m.boundary_mask),
                                     T(m.dirichlet_mask), 0.0, T(m.neumann_edges))
    # u_free = 1e-5*randn as in the reference ctor (models.py:274), seeded per global node so strips agree
    u0 = 1e-5 * (2.0 * meshgen._hash_u01(np.stack([m.global_node_id * 2, m.global_node_id * 2 + 1], 1) + (1 << 50), 7) - 1.0)
    with torch.no_grad():
        model.u_free.copy_(T(u0[~m.dirichlet_mask], dtype=torch.float32))
    if dtype == torch.float64:
        model = model.double()
    model.tile_nodes = tile_nodes
    model = model.to(device)
    if world > 1:
        halo = hd.setup_strip_halo(m, m.boundary_mask, m.dirichlet_mask, device, dtype)
        model.priority_nodes = halo.priority_nodes      # tiles owning shared nodes run first: the exchange overlaps the rest
        loss_fn = hd.DistributedEnergyLoss2D(E=10e9, nu=0.3, length=2.0, height=1.0, device=device, dtype=dtype, halo=halo)
    else:
        loss_fn = EnergyLoss2D(E=10e9, nu=0.3, length=2.0, height=1.0, device=device, dtype=dtype)
    return m, model, loss_fn, (nx, ny)


def make_step(model, loss_fn, graph=True):
    if graph:
        # public API: hidenn_fem_b200.graph.GraphedEnergyStep == zero_grad + loss_fn(model) + backward, replayed
        from hidenn_fem_b200.graph import GraphedEnergyStep
        return GraphedEnergyStep(model, loss_fn)

    def step():
        model.zero_grad(set_to_none=True)
        loss = loss_fn(model)
        loss.backward()
        return loss
    return step


def timed_loop(step, steps, world):
    """`steps` calls bracketed by barrier + synchronize, CUDA events on the launch stream, max over ranks -> (ms total, last loss)."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        ms = float(t.item())
    return ms, loss


def time_steps(step, steps, warmup, world):
    for _ in range(warmup):
        step()
    ms, loss = timed_loop(step, steps, world)
    return ms / steps, float(loss.item())


def time_e2e_dist(model, step, steps, warmup, world):
    """End to end at N > 1 through the Python API: every step copies the rank's Parameter rows from pinned host memory,
    runs the (graphed) step incl. the halo exchange, and reads the loss and both gradients back into pinned host memory.
    Wall clock between barriers, max over ranks."""
    import torch.distributed as dist
    xf_h = model.node_coords_free.detach().cpu().pin_memory()
    uf_h = model.u_free.detach().cpu().pin_memory()
    gx_h, gu_h = torch.empty_like(xf_h).pin_memory(), torch.empty_like(uf_h).pin_memory()
    loss_h = torch.empty(1, dtype=xf_h.dtype).pin_memory()

    def call():
        with torch.no_grad():
            model.node_coords_free.copy_(xf_h, non_blocking=True)
            model.u_free.copy_(uf_h, non_blocking=True)
        loss = step()
        gx_h.copy_(model.node_coords_free.grad, non_blocking=True)
        gu_h.copy_(model.u_free.grad, non_blocking=True)
        loss_h.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(max(1, min(warmup, 3))):
        call()
    n = max(3, min(steps, 10))
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    t = torch.tensor([(time.perf_counter() - t0) / n], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d = (xf_h.numel() + uf_h.numel()) * xf_h.element_size()
    d2h = (gx_h.numel() + gu_h.numel() + 1) * xf_h.element_size()
    b = torch.tensor([h2d, d2h], device="cuda", dtype=torch.float64)
    dist.all_reduce(b)
    return float(t.item()) * 1e3, int(b[0].item()), int(b[1].item()), float(loss_h[0])


def time_kernel(model, loss_fn, steps, warmup):
    """Dominant kernel alone (tile kernel, HIDENN_TILES_ONLY) with CUDA events on the launch stream."""
    import ctypes as C
    from hidenn_fem_b200 import _lib
    plan = model._plan()
    dt = model.dtype
    consts, hints = loss_fn._consts(model, None)
    xb, ub = model._fixed_pair()
    xf, uf = model.node_coords_free.detach(), model.u_free.detach()
    gx, gu = torch.empty_like(xf), torch.empty_like(uf)
    out = torch.empty(4, device=xf.device, dtype=dt)
    scratch = loss_fn._scratch(plan, xf.device, dt)
    f = _lib.fn("hidenn_tri_energy", dt)
    s = _lib.stream_ptr()

    def launch():
        _lib.check(f(plan.handle, _lib.ptr(xf), _lib.ptr(xb), _lib.ptr(uf), _lib.ptr(ub), _lib.ptr(consts), _lib.ptr(None),
                     C.c_int(1 | 2 | 4 | 16 | hints), _lib.ptr(out), _lib.ptr(gx), _lib.ptr(gu), _lib.ptr(None), _lib.ptr(scratch), s))
    for _ in range(warmup):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def time_e2e(model, loss_fn, steps, warmup):
    """Through the C-ABI host-buffer entry point: pinned host parameters in, loss + both gradients out,
    every step (hidenn_tri_energy_host_*)."""
    import ctypes as C
    from hidenn_fem_b200 import _lib
    plan = model._plan()
    dt = model.dtype
    consts, hints = loss_fn._consts(model, None)
    consts = consts.cpu()
    xb, ub = model._fixed_pair()
    xf = model.node_coords_free.detach().cpu().pin_memory()
    uf = model.u_free.detach().cpu().pin_memory()
    xb, ub = xb.cpu().pin_memory(), ub.cpu().pin_memory()
    out = torch.empty(4, dtype=dt).pin_memory()
    gx, gu = torch.empty_like(xf).pin_memory(), torch.empty_like(uf).pin_memory()
    f = _lib.fn("hidenn_tri_energy_host", dt)
    s = _lib.stream_ptr()
    h2d = (xf.numel() + uf.numel() + xb.numel() + ub.numel() + consts.numel()) * xf.element_size()
    d2h = (gx.numel() + gu.numel() + 4) * xf.element_size()

    def call():
        _lib.check(f(plan.handle, _lib.ptr(xf), _lib.ptr(xb), _lib.ptr(uf), _lib.ptr(ub), _lib.ptr(consts), C.c_int(7 | hints),
                     _lib.ptr(out), _lib.ptr(gx), _lib.ptr(gu), s))
    for _ in range(max(1, min(warmup, 3))):
        call()
    torch.cuda.synchronize()
    n = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for _ in range(n):
        call()          # synchronous: returns when loss and gradients are in host memory
    t = (time.perf_counter() - t0) / n
    loss = float(out[0])
    # same call with the three-stream pipeline switched off (copies and kernels back to back), for comparison
    os.environ["HIDENN_HOST_CHUNKS"] = "1"
    try:
        call()
        t0 = time.perf_counter()
        for _ in range(3):
            call()
        t_serial = (time.perf_counter() - t0) / 3
    finally:
        del os.environ["HIDENN_HOST_CHUNKS"]
    # same step with only the loss read back (gradients computed, left on the device)
    def call_loss_only():
        _lib.check(f(plan.handle, _lib.ptr(xf), _lib.ptr(xb), _lib.ptr(uf), _lib.ptr(ub), _lib.ptr(consts), C.c_int(7 | hints),
                     _lib.ptr(out), None, None, s))
    call_loss_only()
    t0 = time.perf_counter()
    for _ in range(3):
        call_loss_only()
    t_loss = (time.perf_counter() - t0) / 3
    return t * 1e3, h2d, d2h, loss, t_serial * 1e3, t_loss * 1e3


def time_loop(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_grid_paths(device, steps, warmup, peak, full_c3=False):
    """Configs C2 (1D bar energy, 1 M elements, FP64) and C3 (structured Q1 L2 projection) of SURVEY §8(d):
    one step = loss forward + backward through the drop-in API.  Informational lines next to the C4 headline."""
    from hidenn_fem_b200.models import PiecewiseLinearShapeNN, StructuredShapeNN2D
    from hidenn_fem_b200 import models_grid as mg
    from hidenn_fem_b200.utils import interval_gauss_points
    out = {}
    # C2: examples/example3.py at 1M elements, ng=2, fused kernel with the built-in body force
    N = 1_000_001
    m1 = PiecewiseLinearShapeNN(torch.linspace(0, 10.0, N, dtype=torch.float64), r_adapt=True, u0=0.0, uN=0.0).double().to(device)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        m1.u.copy_((1e-2 * torch.randn(N - 2, generator=g, dtype=torch.float64)).to(device))
    xi, wi = interval_gauss_points(2, device=device, dtype=torch.float64)

    def step_c2():
        m1.zero_grad(set_to_none=True)
        mg.bar_energy_loss(m1, xi, wi, None, 175.0, b_builtin=True).backward()
    ms = time_loop(step_c2, steps, warmup)
    mg._bar_state.check(block=True)
    out["C2_bar_1M_f64_fused"] = {"ms_per_step": ms, "evals_per_s": (N - 1) * 2 / (ms * 1e-3),
                                  "hbm_frac_informational": 4 * 8 * (N - 1) / (ms * 1e-3) / 1e9 / peak,
                                  "note": "eager: fused bar step = 3 launches (+ ones_like fill and grad_output scale of autograd)"}

    from hidenn_fem_b200.graph import GraphedStep
    g2 = GraphedStep(m1, lambda: mg.bar_energy_loss(m1, xi, wi, None, 175.0, b_builtin=True))
    ms = time_loop(lambda: g2(), steps, warmup)
    mg._bar_state.check(block=True)
    out["C2_bar_1M_f64_fused_graph_replay"] = {"ms_per_step": ms, "evals_per_s": (N - 1) * 2 / (ms * 1e-3),
                                               "hbm_frac_informational": 4 * 8 * (N - 1) / (ms * 1e-3) / 1e9 / peak,
                                               "note": "32 MB working set is L2-resident; the step is bound by the FP64 exponentials of "
                                                       "the example's body force (4 per element) and the softplus / sigmoid of the grid"}
    del g2

    def step_c2g():
        m1.zero_grad(set_to_none=True)
        mg.energy_loss_generic(m1, xi, wi, mg.example3_b_force, E=175.0).backward()
    ms = time_loop(step_c2g, max(3, steps // 4), 2)
    out["C2_bar_1M_f64_generic_double_backward"] = {"ms_per_step": ms, "evals_per_s": (N - 1) * 2 / (ms * 1e-3)}
    del m1
    # C3: examples/example2.py scaled up; full batch of tensor-product samples through the generic forward
    Ng, Ms = (4097, 8192) if full_c3 else (1025, 2048)
    gx = torch.linspace(0, 1, Ng, dtype=torch.float64)
    m2 = StructuredShapeNN2D(gx, gx.clone(), r_adapt=True).double().to(device)
    xs = torch.linspace(0, 1, Ms, dtype=torch.float64, device=device)
    XX, YY = torch.meshgrid(xs, xs, indexing="ij")
    x_train = torch.stack([XX.flatten(), YY.flatten()], dim=1)
    u_true = torch.sin(2 * torch.pi * x_train[:, 0]) * torch.cos(2 * torch.pi * x_train[:, 1])
    del XX, YY

    def step_c3():
        m2.zero_grad(set_to_none=True)
        ((m2(x_train) - u_true) ** 2).mean().backward()
    ms = time_loop(step_c3, max(2, steps // 5), 1)
    M = x_train.shape[0]
    out["C3_structured_l2_%dx%d_nodes_%d_samples_f64" % (Ng, Ng, M)] = {
        "ms_per_step": ms, "evals_per_s": M / (ms * 1e-3),
        "hbm_frac_informational": (3 * 8 * M + 2 * 8 * Ng * Ng) / (ms * 1e-3) / 1e9 / peak,
        "note": "reference expression ((model(x)-u)**2).mean(): shared-memory-staged lookup forward + sort-free deterministic "
                "binning + fused cell fold backward"}

    def step_c3f():
        m2.zero_grad(set_to_none=True)
        mg.l2_projection_loss(m2, x_train, u_true).backward()
    ms = time_loop(step_c3f, max(2, steps // 5), 1)
    out["C3_structured_l2_fused_loss_f64"] = {
        "ms_per_step": ms, "evals_per_s": M / (ms * 1e-3),
        "hbm_frac_informational": (3 * 8 * M + 2 * 8 * Ng * Ng) / (ms * 1e-3) / 1e9 / peak,
        "note": "models_grid.l2_projection_loss: residual weights and loss produced by the forward pass"}
    return out


def kernel_source_hash():
    """sha256 (first 16 hex digits) over the triangle kernel / plan sources: stamps profiles/traffic.json."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "hidenn-fem_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.startswith("tri_") or f == "common.cuh":
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def fp64_peak(device):
    """DFMA instructions / s of this GPU measured by the library's microbenchmark (None if the symbol is missing)."""
    import ctypes as C
    from hidenn_fem_b200 import _lib
    L = _lib.lib()
    if not hasattr(L, "hidenn_fp64_peak"):
        return None
    out = C.c_double(0.0)
    L.hidenn_fp64_peak.restype = C.c_int
    with torch.cuda.device(device):
        rc = L.hidenn_fp64_peak(C.byref(out), _lib.stream_ptr())
    return float(out.value) if rc == 0 else None


# ------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(n_elems, dtype):
    """The reference's CPU torch path on the same generator at a bounded size: the UNMODIFIED reference classes when their
    snapshot exists (oracle/_ref, made by oracle/make_ref.py at build time; kind "reference"), else the op-for-op
    restatement oracle/torch_port.py (kind "port").  Returns (step, elements, kind)."""
    from hidenn_fem_b200 import meshgen
    from oracle import torch_port as tp
    from oracle import closed_form as cf
    from oracle import make_ref
    torch.set_num_threads(os.cpu_count() or 1)
    nx, ny = meshgen.plate_dims_for_elements(n_elems)
    m = meshgen.plate_mesh(nx, ny, jitter=0.25, diag="random", seed=0, ordering="morton")
    T = torch.tensor
    npdt = np.float64 if dtype == torch.float64 else np.float32
    u0 = 1e-5 * (2.0 * meshgen._hash_u01(np.stack([m.global_node_id * 2, m.global_node_id * 2 + 1], 1) + (1 << 50), 7) - 1.0)
    ref = make_ref.load()
    if ref is not None:
        ref_models, ref_loss = ref
        torch.manual_seed(0)
        model = ref_models.PiecewiseLinearShapeNN2D(T(m.node_coords.astype(npdt)), T(m.connectivity), T(m.boundary_mask),
                                                    T(m.dirichlet_mask), 0.0, T(m.neumann_edges))
        with torch.no_grad():
            model.u_free.copy_(T(u0[~m.dirichlet_mask].astype(np.float32)))
        if dtype == torch.float64:
            model = model.double()
        loss_fn = ref_loss.EnergyLoss2D(E=10e9, nu=0.3, length=2.0, height=1.0, device=torch.device("cpu"), dtype=dtype)

        def step():
            model.zero_grad()
            loss = loss_fn(model)
            loss.backward()
            return float(loss.detach())
        return step, m.connectivity.shape[0], "reference"
    port = tp.TriPort(T(m.node_coords.astype(npdt)), T(m.connectivity), T(m.boundary_mask), T(m.dirichlet_mask), 0.0,
                      T(m.neumann_edges), u_free=T(u0[~m.dirichlet_mask].astype(npdt)))
    xg, wg = cf.triangle_gauss_points(4, npdt)
    xi1, w1 = cf.interval_gauss_points(2, npdt)
    Cm = T(cf.plane_stress_C(10e9, 0.3, npdt))
    xg, wg, xi1, w1 = T(xg), T(wg), T(xi1), T(w1)

    def step():
        port.x_free.grad = None
        port.u_free.grad = None
        loss = tp.tri_energy(port, Cm, xg, wg, xi1, w1)
        loss.backward()
        return float(loss.detach())
    return step, m.connectivity.shape[0], "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    step, ne, kind = cpu_reference_step_factory(args.cpu_sample_elems, dtype)
    # --steps / --warmup are honoured as given: one step of the 1 M-element sample is ~1 s of CPU work on 16 cores,
    # so the driver's 20 + 5 fit in half a minute
    k, w = max(1, args.steps), max(0, args.warmup)
    for _ in range(w):
        step()
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    t = (time.perf_counter() - t0) / k
    val = ne * NG / t
    cores = torch.get_num_threads()
    sample = f"{ne} triangles of the same plate generator (C4 scaled down), 1 step = loss+backward, mean of {k}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
        "warmup": w, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_total": args.elems, "sample_elements": ne, "gauss_points": NG,
                   "note": "CPU arm: every step is a bounded sample of the workload (same generator, scaled down)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    if args.watchdog_s > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog_s, exit=True)      # a hung collective must not hold the GPU box
    if args.impl == "reference":
        run_reference(args)
        return
    # stdout carries exactly one JSON line: anything libraries print (NCCL's version banner, warnings) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    sz = 8 if dtype == torch.float64 else 4

    t0 = time.time()
    elems_per_gpu = args.elems if args.scaling == "weak" else max(1, args.elems_total // world)
    m, model, loss_fn, dims = make_workload(args, rank, world, device, dtype, args.ordering, elems_per_gpu, args.tile_nodes)
    plan = model._plan()
    setup_s = time.time() - t0
    ne_local = m.connectivity.shape[0]
    nn_local = m.node_coords.shape[0]

    step = make_step(model, loss_fn, graph=not args.no_graph)
    sampler = ClockSampler(physical_index(local))
    sampler.start()
    ms_step, loss_val = time_steps(step, args.steps, args.warmup, world)
    clocks = sampler.stop()
    # sustained leg: >= 1 s of back-to-back steps with its own clock record (the K-step figure above is a burst of a few ms)
    sustained = None
    if not args.no_sustained:
        n_sus = int(min(20000, max(args.steps, np.ceil(1100.0 / max(ms_step, 1e-3)))))
        s2 = ClockSampler(physical_index(local))
        s2.start()
        ms_sus, _ = timed_loop(step, n_sus, world)
        sustained = {"steps": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus, "clocks": s2.stop()}
    ms_kernel = time_kernel(model, loss_fn, args.steps, args.warmup)

    ne_tot = torch.tensor([ne_local, nn_local], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ne_tot)
    ne_total, nn_total = int(ne_tot[0].item()), int(ne_tot[1].item())
    value = ne_total * NG / (ms_step * 1e-3)
    if sustained is not None:
        sustained["value"] = ne_total * NG / (sustained["ms_per_step"] * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    # algorithmic bytes of one launch: int32 connectivity + read x,u + write dx,du of the free rows
    nfree_x, nfree_u = plan.info["n_free_x"], plan.info["n_free_u"]
    alg_bytes = 12 * ne_local + 2 * sz * (2 * nn_local) + 2 * sz * (nfree_x + nfree_u)
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": {9: ("tri_tile9_kernel<double> (warp-specialised: 16 element warps on edge-sharing element pairs, "
                                                "6 fold warps, 2 loader warps; bulk-copy stage ring; tile-ordered numbering)"
                                                if plan.info.get("n_pairs", 0) > 0 else
                                                "tri_tile9_kernel<double> (warp-specialised, one element per entry: 12 element warps, 10 fold warps, 2 loader warps; "
                                                "bulk-copy stage ring; tile-ordered numbering)"),
                                            8: "tri_tile8_kernel<double> (two CTAs per SM, bulk-copy staging; tile-ordered numbering)"}.get(
                                                plan.info.get("kernel"), "tri_tile_persistent_kernel<%s>" % ("double" if sz == 8 else "float")),
                "kernel_ms": ms_kernel, "algorithmic_bytes_per_launch": alg_bytes,
                "bytes_per_element": alg_bytes / ne_local, "peak_source": peak_src}
    # DRAM bytes of one launch from the committed ncu capture -- only if that capture was made from the kernel sources
    # that are running now (profiles/traffic.json carries their hash), else null
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    src_hash = kernel_source_hash()
    roofline["kernel_source_sha256_16"] = src_hash
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get(args.dtype + ("_tiles" if plan.info.get("tile_ordered") else ""))
            if isinstance(ent, dict) and ent.get("source_sha256_16") == src_hash:
                roofline["traffic"] = ent.get("dram_bytes_per_launch")
                roofline["traffic_source"] = ent.get("from")
        except Exception:
            pass
    fp = fp64_peak(device)
    if fp is not None and sz == 8:
        # FP64 instructions one launch issues (thread level): 68 per element visit (closed form, SASS count) + 4 per fold slot
        # + 8 per merged pair partial; against the DFMA issue rate measured by hidenn_fp64_peak on this GPU
        n_slots = plan.info.get("fold_slots", 0) or 3 * plan.info["elem_visits"]
        fp64_instr = 68.0 * plan.info["elem_visits"] + 4.0 * n_slots
        roofline["fp64_pipe"] = {"measured_peak_dfma_per_s": fp, "measured_peak_tflops": 2.0 * fp / 1e12,
                                 "kernel_fp64_instr_per_launch": fp64_instr,
                                 "frac": fp64_instr / (ms_kernel * 1e-3) / fp}

    e2e = None
    if not args.no_e2e and world == 1:
        ms_e2e, h2d, d2h, l_e2e, ms_serial, ms_lossonly = time_e2e(model, loss_fn, args.steps, args.warmup)
        e2e = {"value": ne_local * NG / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ms_e2e, "ms_per_step_unpipelined": ms_serial, "ms_per_step_loss_only_readback": ms_lossonly,
               "api": "hidenn_tri_energy_host_%s (pinned host buffers; rows in / tiles / gradient rows out overlapped on 3 streams)" % args.dtype,
               "loss_matches_resident": bool(abs(l_e2e - loss_val) <= 1e-9 * abs(loss_val))}
    elif not args.no_e2e and world > 1:
        ms_e2e, h2d, d2h, l_e2e = time_e2e_dist(model, step, args.steps, args.warmup, world)
        e2e = {"value": ne_total * NG / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
               "api": "pinned host Parameter rows -> device, GraphedEnergyStep incl. halo exchange, loss + both gradients -> pinned host "
                      "(all ranks, bytes summed over ranks)",
               "loss_matches_resident": bool(abs(l_e2e - loss_val) <= 1e-9 * abs(loss_val))}

    extra = {}
    if args.extra and world == 1:
        for tag, dt2, ordr in (("f64_morton", torch.float64, "morton"), ("f64_natural", torch.float64, "natural"),
                               ("f64_random_numbering", torch.float64, "random"),
                               ("f64_random_numbering_after_reorder_for_locality", torch.float64, "random+reorder"),
                               ("f64_tiles_corners_rotated_at_random", torch.float64, "tiles+rotated"),
                               ("f32_tiles", torch.float32, "tiles"), ("f32_morton", torch.float32, "morton")):
            del model, loss_fn
            torch.cuda.empty_cache()
            m2, model, loss_fn, _ = make_workload(args, 0, 1, device, dt2, ordr, args.elems, args.tile_nodes)
            ms2, _ = time_steps(make_step(model, loss_fn, graph=not args.no_graph), args.steps, args.warmup, 1)
            k2 = time_kernel(model, loss_fn, args.steps, args.warmup)
            s2 = 8 if dt2 == torch.float64 else 4
            p2 = model._plan()
            ab = 12 * m2.connectivity.shape[0] + 2 * s2 * 2 * m2.node_coords.shape[0] + 2 * s2 * (p2.info["n_free_x"] + p2.info["n_free_u"])
            extra[tag] = {"ms_per_step": ms2, "kernel_ms": k2, "evals_per_s": m2.connectivity.shape[0] * NG / (ms2 * 1e-3),
                          "roofline_frac": ab / (k2 * 1e-3) / 1e9 / peak}

    other = None
    if not args.no_grid and world == 1:
        try:
            del model, loss_fn, step
            torch.cuda.empty_cache()
            other = bench_grid_paths(device, args.steps, args.warmup, peak, full_c3=True)
        except Exception as e:      # the headline must survive a failure of these lines
            other = {"error": repr(e)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        step, ne_s, cpu_kind = cpu_reference_step_factory(args.cpu_sample_elems, dtype)
        step()
        best = 1e30
        for _ in range(3):
            t1 = time.perf_counter()
            step()
            best = min(best, time.perf_counter() - t1)
        cpu_baseline = {"value": ne_s * NG / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu_kind,
                        "sample": f"{ne_s} triangles, same generator, "
                                  + ("the unmodified reference classes (oracle/_ref snapshot)" if cpu_kind == "reference"
                                     else "torch-autograd port of the reference path (oracle/torch_port.py)")
                                  + f", 1 warm-up + best of 3, {best * 1e3:.0f} ms/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "elements_total": ne_total, "nodes_total": nn_total, "elements_per_gpu": ne_local,
                       "grid_nodes": list(dims), "ordering": args.ordering, "partition": "column strips + halo nodes" if world > 1 else "none",
                       "halo_exchange": (getattr(getattr(loss_fn, "halo", None), "backend", None) if world > 1 else None),
                       "first_tiles": plan.info.get("n_first_tiles", 0),
                       "launch": "eager" if args.no_graph else "CUDA-graph replay of zero_grad+loss+backward (hidenn_fem_b200.graph.GraphedEnergyStep)",
                       "l2_policy": "inputs+outputs+plan per launch (> 400 MB) exceed the 126 MB L2; no flush needed",
                       "tile_nodes": plan.info["max_local"], "n_tiles": plan.info["n_tiles"],
                       "halo_recompute": plan.info["elem_visits"] / max(1, ne_local), "setup_s": setup_s},
            "loss": loss_val,
            "sustained": sustained,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": args.steps * ((1 if plan.info.get("tile_ordered") else 2) + (2 if world > 1 else 0)),
            "gpu_launches_note": ("per replayed step: the tile kernel (Neumann edges, final reduction and the loss exchange inside)"
                                  if plan.info.get("tile_ordered") else
                                  "per replayed step: tri_tile_persistent_kernel + tri_edge_finalize_kernel")
                                 + (" + halo_p2p push + pull on a side stream (peer-memory exchange; pack_all + NCCL all-reduce + "
                                    "unpack_all after the replay with HIDENN_HALO=nccl)" if world > 1 else ""),
        }
        if other is not None:
            line["other_configs"] = other      # BASELINE configs C2 (1D bar, 1 M elements) and C3 (structured L2, 4097^2 nodes)
        if extra:
            line["extra"] = extra
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
