"""CUDA-graph replay of one energy step (loss forward + backward).

A step of the fused path is one launch on one GPU for tile-ordered FP64 plans (tile kernel with the Neumann edges, the
final reduction and, with several GPUs, the loss exchange in its tail) plus the put / complete kernels of the peer-memory
halo exchange on a side stream; at ~0.2 ms of GPU time per 10 M elements the Python / autograd enqueue cost (~0.2 ms) is
of the same size.  Capturing the step once and replaying it removes that cost.  The peer-memory halo exchange
(dist.HaloP2P) is captured with the step; the NCCL variant stays outside the graph (pack launch, eager all-reduce, unpack
launch after every replay).
The parameters keep their storage (optimisers update them in place), the gradients live in buffers owned by the
graph and are overwritten by every replay -- the usual whole-step capture contract of torch.cuda.graphs.
"""
from __future__ import annotations

import torch


class GraphedStep:
    """step = GraphedStep(model, fn);  loss = step()  ==  model.zero_grad(); loss = fn(); loss.backward()

    `fn` is any capture-safe closure over the model's parameters built from this package's losses
    (`lambda: loss_fn(model)`, `lambda: bar_energy_loss(model, xi, wi, None, E, b_builtin=True)`,
    `lambda: l2_projection_loss(model, x, u)`): no host synchronisation, fixed shapes.  After each call `p.grad` of
    every trainable parameter holds the fresh gradient (same tensors every time) and the returned 0-dim tensor the
    loss.  Gradient accumulation across calls is not available in this mode."""

    def __init__(self, model, fn, warmup: int = 3, one=None):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs CUDA parameters (no CPU fallback)")
        self.model = model

        if one is None:
            def one():
                loss = fn()
                loss.backward()
                return loss

        # warm-up on a side stream (constant tables, scratch, plan upload), as capture requires
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                model.zero_grad(set_to_none=True)
                one()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss = one()
        self.loss = loss.detach()
        # the gradient tensors the graph writes; re-attached on every call so that an optimizer.zero_grad() in between
        # (set_to_none=True is torch's default) cannot detach the parameters from them
        self._static = [(p, p.grad) for p in model.parameters() if p.grad is not None]

    def __call__(self):
        self.graph.replay()
        for p, g in self._static:
            p.grad = g
        return self.loss


class GraphedEnergyStep(GraphedStep):
    """step = GraphedEnergyStep(model, loss_fn);  loss = step()  ==  zero_grad(); loss = loss_fn(model); loss.backward()

    `loss_fn_args` are forwarded to the loss (b_force / t_force callables must be capture-safe: no host syncs).
    With a distributed loss (dist.DistributedEnergyLoss2D) the graph holds the rank-local part and the halo exchange
    (pack, NCCL all-reduce, unpack) runs eagerly after every replay."""

    def __init__(self, model, loss_fn, *loss_fn_args, warmup: int = 3, **loss_fn_kw):
        self.loss_fn = loss_fn
        self._halo = getattr(loss_fn, "halo", None)
        if self._halo is not None and getattr(self._halo, "capturable", False):
            self._halo = None             # peer-memory exchange (dist.HaloP2P): its kernels are part of the captured step
        if self._halo is not None:
            loss_fn.halo = None           # NCCL exchange: the graph holds the rank-local part, the exchange follows each replay
        def one():
            # zero_grad(); loss = loss_fn(model); loss.backward() with the implicit grad_output = 1: the fused launch has
            # already produced d loss / d parameters, so they are attached directly -- no autograd pass, no ones_like fill, no
            # scale kernel in the replayed step (a frozen Parameter gets no gradient, as in the eager path)
            loss = loss_fn(model, *loss_fn_args, **loss_fn_kw)
            gx, gu = loss_fn.last_grads
            if model.node_coords_free.requires_grad:
                model.node_coords_free.grad = gx
            if model.u_free.requires_grad:
                model.u_free.grad = gu
            return loss
        try:
            super().__init__(model, None, warmup=warmup, one=one)
            self._parts = loss_fn.last_parts          # [loss, domain, edge, 0] written by the finalize kernel;
        finally:                                      # self.loss is a view of _parts[0]: the exchange completes it in place
            if self._halo is not None:
                loss_fn.halo = self._halo
        gx = getattr(model, "node_coords_free", None)
        gu = getattr(model, "u_free", None)
        self._gx = None if gx is None else gx.grad
        self._gu = None if gu is None else gu.grad

    def __call__(self):
        super().__call__()
        if self._halo is not None:
            self._halo.exchange(self._parts, self._gx, self._gu)
        return self.loss
