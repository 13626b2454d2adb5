"""L-BFGS for parameters that are sharded across ranks with replicated halo entries.

The reference trains with `torch.optim.LBFGS(model.parameters())` (examples/example4.py:68-80: lr=1, max_iter=20,
history 100, no line search).  With the mesh partitioned over GPUs (dist.py) every rank holds its own rows plus copies
of the rows it shares with neighbours, so the optimiser's inner products and norms must be GLOBAL and must count every
shared row once.  `ShardedLBFGS` is the same algorithm (two-loop recursion, first step min(1, 1/|g|_1), the same
stopping tests in the same order) with every reduction routed through `weights` (1 for rows this rank owns, 0 for
copies owned elsewhere) and one all-reduce; since the halo exchange leaves bit-identical gradients on every copy and
all step coefficients are global scalars, the copies of a shared row stay bit-identical on all ranks.
With `weights=None` and no process group it is a drop-in for the stock optimiser on one GPU.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import torch
import torch.distributed as dist


class ShardedLBFGS(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1.0, max_iter: int = 20, max_eval: Optional[int] = None,
                 tolerance_grad: float = 1e-7, tolerance_change: float = 1e-9, history_size: int = 100,
                 weights: Optional[Sequence[Optional[torch.Tensor]]] = None, group=None):
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("ShardedLBFGS supports a single parameter group (like torch.optim.LBFGS)")
        self._params = self.param_groups[0]["params"]
        self._group = group
        self._distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self._w = None
        if weights is not None:
            if len(weights) != len(self._params):
                raise ValueError("weights: one entry (tensor or None) per parameter")
            parts = []
            for p, w in zip(self._params, weights):
                if w is None:
                    parts.append(torch.ones(p.numel(), device=p.device, dtype=p.dtype))
                else:                     # one weight per row (leading dimension) or per entry
                    w = w.to(device=p.device, dtype=p.dtype)
                    if w.numel() != p.numel():
                        w = w.reshape([-1] + [1] * (p.dim() - 1)).expand_as(p)
                    parts.append(w.reshape(-1))
            self._w = torch.cat(parts)

    # -- flat views -----------------------------------------------------------------------------
    def _flat_grad(self) -> torch.Tensor:
        return torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1) for p in self._params])

    def _add(self, step: float, direction: torch.Tensor) -> None:
        off = 0
        with torch.no_grad():
            for p in self._params:
                n = p.numel()
                p.add_(direction[off:off + n].view_as(p), alpha=step)
                off += n

    # -- global reductions ------------------------------------------------------------------------
    def _reduce(self, t: torch.Tensor, op) -> float:
        t = t.to(torch.float64).reshape(1)
        if self._distributed:
            dist.all_reduce(t, op=op, group=self._group)
        return float(t.item())

    def _dot(self, a: torch.Tensor, b: torch.Tensor) -> float:
        prod = a * b
        if self._w is not None:
            prod = prod * self._w
        return self._reduce(prod.sum(dtype=torch.float64), dist.ReduceOp.SUM)

    def _abs_sum(self, a: torch.Tensor) -> float:
        v = a.abs()
        if self._w is not None:
            v = v * self._w
        return self._reduce(v.sum(dtype=torch.float64), dist.ReduceOp.SUM)

    def _abs_max(self, a: torch.Tensor) -> float:
        return self._reduce(a.abs().max() if a.numel() else a.new_zeros(()), dist.ReduceOp.MAX)

    # -- one optimiser step (up to max_iter inner iterations, like the stock LBFGS) ----------------
    @torch.no_grad()
    def step(self, closure: Callable[[], torch.Tensor]):
        g0 = self.param_groups[0]
        lr, max_iter, max_eval = g0["lr"], g0["max_iter"], g0["max_eval"]
        tol_g, tol_x, hist = g0["tolerance_grad"], g0["tolerance_change"], g0["history_size"]
        closure = torch.enable_grad()(closure)
        st = self.state[self._params[0]]
        st.setdefault("func_evals", 0)
        st.setdefault("n_iter", 0)

        first_loss = closure()
        loss = float(first_loss)
        evals = 1
        st["func_evals"] += 1
        g = self._flat_grad()
        if self._abs_max(g) <= tol_g:
            return first_loss

        d, t = st.get("d"), st.get("t")
        ys_hist, s_hist, rho = st.get("ys_hist", []), st.get("s_hist", []), st.get("rho", [])
        g_prev, h_diag = st.get("g_prev"), st.get("h_diag", 1.0)
        loss_prev = st.get("loss_prev")

        it = 0
        while it < max_iter:
            it += 1
            st["n_iter"] += 1
            if st["n_iter"] == 1:
                d = g.neg()
                ys_hist, s_hist, rho, h_diag = [], [], [], 1.0
            else:
                y = g - g_prev
                s = d * t
                ys = self._dot(y, s)
                if ys > 1e-10:                       # curvature pair accepted
                    if len(ys_hist) == hist:
                        ys_hist.pop(0); s_hist.pop(0); rho.pop(0)
                    ys_hist.append(y); s_hist.append(s); rho.append(1.0 / ys)
                    h_diag = ys / self._dot(y, y)
                k = len(ys_hist)
                alpha = [0.0] * k
                q = g.neg()
                for i in range(k - 1, -1, -1):
                    alpha[i] = self._dot(s_hist[i], q) * rho[i]
                    q.add_(ys_hist[i], alpha=-alpha[i])
                d = q.mul_(h_diag)
                for i in range(k):
                    beta = self._dot(ys_hist[i], d) * rho[i]
                    d.add_(s_hist[i], alpha=alpha[i] - beta)
            g_prev = g.clone()
            loss_prev = loss
            t = min(1.0, 1.0 / self._abs_sum(g)) * lr if st["n_iter"] == 1 else lr
            gtd = self._dot(g, d)
            if gtd > -tol_x:                         # not a descent direction any more
                break
            self._add(t, d)
            new_evals = 0
            if it != max_iter:                       # the last inner iteration leaves re-evaluation to the next step()
                loss = float(closure())
                g = self._flat_grad()
                new_evals = 1
            evals += new_evals
            st["func_evals"] += new_evals
            if it == max_iter or evals >= max_eval:
                break
            if self._abs_max(g) <= tol_g:
                break
            if self._abs_max(d) * t <= tol_x:
                break
            if abs(loss - loss_prev) < tol_x:
                break

        st.update(d=d, t=t, ys_hist=ys_hist, s_hist=s_hist, rho=rho, g_prev=g_prev, h_diag=h_diag, loss_prev=loss_prev)
        return first_loss
