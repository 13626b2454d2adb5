"""L-BFGS for parameters that are sharded across ranks with replicated halo entries.

The reference trains with `torch.optim.LBFGS(model.parameters())` (examples/example4.py:68-80: lr=1, max_iter=20,
history 100, no line search).  With the mesh partitioned over GPUs (dist.py) every rank holds its own rows plus copies
of the rows it shares with neighbours, so the optimiser's inner products and norms must be GLOBAL and must count every
shared row once.  `ShardedLBFGS` is the same algorithm (two-loop recursion, first step min(1, 1/|g|_1), the same
stopping tests in the same order) with every reduction routed through `weights` (1 for rows this rank owns, 0 for
copies owned elsewhere) and one all-reduce; since the halo exchange leaves bit-identical gradients on every copy and
all step coefficients are global scalars, the copies of a shared row stay bit-identical on all ranks.
With `weights=None` and no process group it is a drop-in for the stock optimiser on one GPU.
`line_search_fn="strong_wolfe"` adds the strong-Wolfe line search of the stock optimiser (bracketing + zoom with safeguarded
cubic interpolation, Nocedal & Wright Alg. 3.5 / 3.6) with global inner products: it makes the iteration a descent method,
so rounding differences between summation orders no longer grow along the trajectory as they do with the reference's
fixed step lr = 1.  `history_dtype=torch.float32` stores the (s, y) history in single precision (inner products still
accumulate in FP64): the reference's history_size = 100 costs 200 x the parameter memory -- 32 GB per GPU at 10 M elements
in FP64, 16 GB in FP32.

Two forms of the same iteration: the textbook two-loop recursion (one all-reduce per inner product, 2·history + 5 per
iteration) and the vector-free form (default when distributed): the inner products of the new (s, y, g) with all stored
pairs travel in ONE all-reduce, the recursion runs on the (2m+1)² Gram matrix on the host, and the direction is one
linear combination of the stored vectors -- two passes over the history instead of four, two collectives per iteration.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import torch
import torch.distributed as dist


class ShardedLBFGS(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1.0, max_iter: int = 20, max_eval: Optional[int] = None,
                 tolerance_grad: float = 1e-7, tolerance_change: float = 1e-9, history_size: int = 100,
                 weights: Optional[Sequence[Optional[torch.Tensor]]] = None, group=None, vector_free: Optional[bool] = None,
                 line_search_fn: Optional[str] = None, history_dtype: Optional[torch.dtype] = None):
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        if line_search_fn not in (None, "strong_wolfe"):
            raise RuntimeError("only 'strong_wolfe' is supported")          # the stock optimiser's message
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size, line_search_fn=line_search_fn)
        self._hist_dtype = history_dtype
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("ShardedLBFGS supports a single parameter group (like torch.optim.LBFGS)")
        self._params = self.param_groups[0]["params"]
        self._group = group
        self._distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        # vector-free form (Chen et al., "Large-scale L-BFGS using MapReduce", 2014): all inner products of an iteration
        # travel in ONE all-reduce and the two-loop recursion runs on the (2m+1)^2 Gram matrix; default when distributed
        self._vector_free = self._distributed if vector_free is None else bool(vector_free)
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

    # -- strong-Wolfe line search (global reductions) ----------------------------------------------------
    @staticmethod
    def _cubic_min(x1, f1, g1, x2, f2, g2, bounds=None):
        """Minimiser of the cubic through (x1, f1, g1), (x2, f2, g2), clipped to `bounds` (midpoint if it has none)."""
        lo, hi = bounds if bounds is not None else ((x1, x2) if x1 <= x2 else (x2, x1))
        d1 = g1 + g2 - 3.0 * (f1 - f2) / (x1 - x2)
        sq = d1 * d1 - g1 * g2
        if sq >= 0.0:
            d2 = sq ** 0.5
            if x1 <= x2:
                m = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2))
            else:
                m = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.0 * d2))
            return min(max(m, lo), hi)
        return 0.5 * (lo + hi)

    def _strong_wolfe(self, evaluate, t, d, f, g, gtd, c1=1e-4, c2=0.9, tol_x=1e-9, max_ls=25):
        """evaluate(t) -> (loss, flat grad) at x0 + t d.  Returns (loss, grad, t, evaluations)."""
        d_norm = self._abs_max(d)
        f_new, g_new = evaluate(t)
        n_eval = 1
        gtd_new = self._dot(g_new, d)
        t_prev, f_prev, g_prev, gtd_prev = 0.0, f, g, gtd
        done = False
        it = 0
        br = br_f = br_g = br_gtd = None
        while it < max_ls:                                   # bracketing phase
            if f_new > f + c1 * t * gtd or (it > 1 and f_new >= f_prev):
                br, br_f, br_g, br_gtd = [t_prev, t], [f_prev, f_new], [g_prev, g_new.clone()], [gtd_prev, gtd_new]
                break
            if abs(gtd_new) <= -c2 * gtd:
                br, br_f, br_g, br_gtd = [t], [f_new], [g_new], [gtd_new]
                done = True
                break
            if gtd_new >= 0:
                br, br_f, br_g, br_gtd = [t_prev, t], [f_prev, f_new], [g_prev, g_new.clone()], [gtd_prev, gtd_new]
                break
            lo, hi = t + 0.01 * (t - t_prev), t * 10.0
            t_next = self._cubic_min(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(lo, hi))
            t_prev, f_prev, g_prev, gtd_prev = t, f_new, g_new.clone(), gtd_new
            t = t_next
            f_new, g_new = evaluate(t)
            n_eval += 1
            gtd_new = self._dot(g_new, d)
            it += 1
        if it == max_ls:                                     # no bracket found: keep the last point
            br, br_f, br_g, br_gtd = [0.0, t], [f, f_new], [g, g_new], [gtd, gtd_new]
        stalled = False
        lo_i, hi_i = (0, 1) if len(br) == 2 and br_f[0] <= br_f[-1] else (1, 0)
        while not done and it < max_ls:                      # zoom phase
            if abs(br[1] - br[0]) * d_norm < tol_x:
                break
            t = self._cubic_min(br[0], br_f[0], br_gtd[0], br[1], br_f[1], br_gtd[1])
            bmax, bmin = max(br), min(br)
            eps = 0.1 * (bmax - bmin)
            if min(bmax - t, t - bmin) < eps:                # too close to an end: make progress, or move 10 % inside
                if stalled or t >= bmax or t <= bmin:
                    t = bmax - eps if abs(t - bmax) < abs(t - bmin) else bmin + eps
                    stalled = False
                else:
                    stalled = True
            else:
                stalled = False
            f_new, g_new = evaluate(t)
            n_eval += 1
            gtd_new = self._dot(g_new, d)
            it += 1
            if f_new > f + c1 * t * gtd or f_new >= br_f[lo_i]:
                br[hi_i], br_f[hi_i], br_g[hi_i], br_gtd[hi_i] = t, f_new, g_new.clone(), gtd_new
                lo_i, hi_i = (0, 1) if br_f[0] <= br_f[1] else (1, 0)
            else:
                if abs(gtd_new) <= -c2 * gtd:
                    done = True
                elif gtd_new * (br[hi_i] - br[lo_i]) >= 0:
                    br[hi_i], br_f[hi_i], br_g[hi_i], br_gtd[hi_i] = br[lo_i], br_f[lo_i], br_g[lo_i], br_gtd[lo_i]
                br[lo_i], br_f[lo_i], br_g[lo_i], br_gtd[lo_i] = t, f_new, g_new.clone(), gtd_new
        if len(br) == 1:
            lo_i = 0
        return br_f[lo_i], br_g[lo_i], br[lo_i], n_eval

    def _search(self, closure, t, d, loss, g, gtd, tol_x):
        """Line search along d from the current parameters; leaves them at the accepted point."""
        cur = [0.0]

        def evaluate(tt):
            self._add(tt - cur[0], d)
            cur[0] = tt
            with torch.enable_grad():
                l = float(closure().detach())
            return l, self._flat_grad()
        loss, g, t, n = self._strong_wolfe(evaluate, t, d, loss, g, gtd, tol_x=tol_x)
        self._add(t - cur[0], d)                              # the accepted point is in general not the last one evaluated
        return loss, g, t, n

    # -- one optimiser step (up to max_iter inner iterations, like the stock LBFGS) ----------------
    @torch.no_grad()
    def step(self, closure: Callable[[], torch.Tensor]):
        if self._vector_free:
            return self._step_vector_free(closure)
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
                    hd = self._hist_dtype or y.dtype
                    ys_hist.append(y.to(hd)); s_hist.append(s.to(hd)); rho.append(1.0 / ys)
                    h_diag = ys / self._dot(y, y)
                k = len(ys_hist)
                alpha = [0.0] * k
                q = g.neg()
                for i in range(k - 1, -1, -1):
                    alpha[i] = self._dot(s_hist[i].to(q.dtype), q) * rho[i]
                    q.add_(ys_hist[i].to(q.dtype), alpha=-alpha[i])
                d = q.mul_(h_diag)
                for i in range(k):
                    beta = self._dot(ys_hist[i].to(d.dtype), d) * rho[i]
                    d.add_(s_hist[i].to(d.dtype), alpha=alpha[i] - beta)
            g_prev = g.clone()
            loss_prev = loss
            t = min(1.0, 1.0 / self._abs_sum(g)) * lr if st["n_iter"] == 1 else lr
            gtd = self._dot(g, d)
            if gtd > -tol_x:                         # not a descent direction any more
                break
            new_evals = 0
            if g0["line_search_fn"] == "strong_wolfe":
                loss, g, t, new_evals = self._search(closure, t, d, loss, g, gtd, tol_x)
            else:
                self._add(t, d)
                if it != max_iter:                   # the last inner iteration leaves re-evaluation to the next step()
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


    # -- vector-free variant -------------------------------------------------------------------------
    def _reduce_vec(self, t: torch.Tensor, op) -> torch.Tensor:
        t = t.to(torch.float64).contiguous()
        if self._distributed:
            dist.all_reduce(t, op=op, group=self._group)
        return t.cpu()

    @torch.no_grad()
    def _step_vector_free(self, closure: Callable[[], torch.Tensor]):
        """Same iteration as `step`, but per inner iteration: one SUM all-reduce of 3(2k+1)+9 inner products (new s, y, g
        against the stored pairs and themselves), the two-loop recursion on the Gram matrix (host, scalars only), the
        direction as one linear combination of the stored vectors, and one MAX all-reduce for the two stopping norms."""
        import numpy as np
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
        n = g.numel()
        w = self._w

        S, Y = st.get("vf_S"), st.get("vf_Y")                  # [hist, n] ring storage, logical order kept in `order`
        if S is None:                                          # capacity grows 8 -> 16 -> ... -> hist rows as pairs arrive
            S = torch.empty(min(hist, 8), n, device=g.device, dtype=self._hist_dtype or g.dtype)
            Y = torch.empty_like(S)
        order = st.get("vf_order", [])                         # physical rows of the stored pairs, oldest first
        G = st.get("vf_G", np.zeros((0, 0)))                   # Gram of [s_0..s_{k-1}, y_0..y_{k-1}] (logical order)
        d, t = st.get("d"), st.get("t")
        g_prev, h_diag, loss_prev = st.get("g_prev"), st.get("h_diag", 1.0), st.get("loss_prev")

        it = 0
        while it < max_iter:
            it += 1
            st["n_iter"] += 1
            if st["n_iter"] == 1:
                d = g.neg()
                order, G, h_diag = [], np.zeros((0, 0)), 1.0
                gtd = -self._dot(g, g)
            else:
                y = g - g_prev
                s = d * t
                k = len(order)
                V = torch.stack([s, y, g])                      # [3, n]
                Vw = V * w if w is not None else V
                parts = [V @ Vw.t()]                            # 3 x 3
                if k:
                    idx = torch.tensor(order, device=g.device)
                    parts += [S[idx].to(g.dtype) @ Vw.t(), Y[idx].to(g.dtype) @ Vw.t()]   # k x 3 each
                red = self._reduce_vec(torch.cat([p.reshape(-1) for p in parts]), dist.ReduceOp.SUM).numpy()
                self_d = red[:9].reshape(3, 3)
                sh = red[9:9 + 3 * k].reshape(k, 3) if k else np.zeros((0, 3))
                yh = red[9 + 3 * k:].reshape(k, 3) if k else np.zeros((0, 3))
                ys = self_d[0, 1]
                sg = sh[:, 2].copy()
                yg = yh[:, 2].copy()
                if ys > 1e-10:                                  # accept the pair: extend (or rotate) the Gram matrix
                    if k == hist:
                        keep = [i for i in range(2 * k) if i not in (0, k)]
                        G = G[np.ix_(keep, keep)]
                        row = order.pop(0)
                        sh, yh, sg, yg = sh[1:], yh[1:], sg[1:], yg[1:]
                        k -= 1
                    else:
                        if k == S.shape[0]:                     # all allocated rows in use: double the capacity
                            cap = min(hist, 2 * S.shape[0])
                            S2 = torch.empty(cap, n, device=g.device, dtype=S.dtype)
                            Y2 = torch.empty_like(S2)
                            S2[:S.shape[0]].copy_(S)
                            Y2[:Y.shape[0]].copy_(Y)
                            S, Y = S2, Y2
                        row = next(r for r in range(S.shape[0]) if r not in order)
                    S[row].copy_(s)
                    Y[row].copy_(y)
                    order.append(row)
                    Gn = np.zeros((2 * k + 2, 2 * k + 2))
                    old = list(range(k)) + list(range(k + 1, 2 * k + 1))          # positions of the old s / y in the new matrix
                    if k:
                        Gn[np.ix_(old, old)] = G
                    # new s at position k, new y at position 2k+1
                    Gn[k, old] = Gn[old, k] = np.concatenate([sh[:, 0], yh[:, 0]])
                    Gn[2 * k + 1, old] = Gn[old, 2 * k + 1] = np.concatenate([sh[:, 1], yh[:, 1]])
                    Gn[k, k] = self_d[0, 0]
                    Gn[2 * k + 1, 2 * k + 1] = self_d[1, 1]
                    Gn[k, 2 * k + 1] = Gn[2 * k + 1, k] = ys
                    G = Gn
                    sg = np.append(sg, self_d[0, 2])
                    yg = np.append(yg, self_d[1, 2])
                    k += 1
                    h_diag = ys / self_d[1, 1]
                # two-loop recursion in coefficient space: direction = sum_i ds[i] s_i + dy[i] y_i + dg g
                bg = np.concatenate([sg, yg])                  # inner products of the basis with g
                gg = self_d[2, 2]
                ds, dy, dg = np.zeros(k), np.zeros(k), -1.0
                alpha = np.zeros(k)

                def dot_with(col):                              # (current combination) . basis[col]
                    v = ds @ G[:k, col] + dy @ G[k:, col] if k else 0.0
                    return v + dg * bg[col]
                for i in range(k - 1, -1, -1):
                    alpha[i] = dot_with(i) / G[i, k + i]
                    dy[i] -= alpha[i]
                ds *= h_diag; dy *= h_diag; dg *= h_diag
                for i in range(k):
                    beta = dot_with(k + i) / G[i, k + i]
                    ds[i] += alpha[i] - beta
                d = g * dg
                if k:
                    idx = torch.tensor(order, device=g.device)
                    coef = torch.tensor(np.concatenate([ds, dy]), device=g.device, dtype=g.dtype)
                    d = d + coef[:k] @ S[idx].to(g.dtype) + coef[k:] @ Y[idx].to(g.dtype)
                gtd = float(ds @ sg + dy @ yg + dg * gg)
            g_prev = g.clone()
            loss_prev = loss
            t = min(1.0, 1.0 / self._abs_sum(g)) * lr if st["n_iter"] == 1 else lr
            if gtd > -tol_x:
                break
            new_evals = 0
            if g0["line_search_fn"] == "strong_wolfe":
                loss, g, t, new_evals = self._search(closure, t, d, loss, g, gtd, tol_x)
            else:
                self._add(t, d)
                if it != max_iter:
                    loss = float(closure())
                    g = self._flat_grad()
                    new_evals = 1
            evals += new_evals
            st["func_evals"] += new_evals
            if it == max_iter or evals >= max_eval:
                break
            norms = self._reduce_vec(torch.stack([g.abs().max(), d.abs().max()]), dist.ReduceOp.MAX)
            if float(norms[0]) <= tol_g:
                break
            if float(norms[1]) * t <= tol_x:
                break
            if abs(loss - loss_prev) < tol_x:
                break

        st.update(d=d, t=t, vf_S=S, vf_Y=Y, vf_order=order, vf_G=G, g_prev=g_prev, h_diag=h_diag, loss_prev=loss_prev)
        return first_loss
