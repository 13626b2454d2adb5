"""Quadrature tables and the gradient smoke check -- mirror of /root/reference/src/utils.py.

Same names, arguments and values as the reference (including its quirks, SURVEY Appendix B
Q2/Q3: order-4/6/7 triangle weights are multiplied by 0.5 so orders 4 and 6 sum to 0.25, and
`interval_gauss_points` returns the raw Gauss-Legendre rule on [-1,1] despite its docstring).
Constants only; no kernels involved.
"""
import numpy as np
import torch


def interval_gauss_points(order=1, device=None, dtype=torch.float32, unit_interval=False):
    """Gauss-Legendre points/weights exactly as reference utils.py:4-11 (raw leggauss, [-1,1]).
    unit_interval=True (correct-math switch, default off): the rule its docstring promises, mapped to [0,1]."""
    xi, wi = np.polynomial.legendre.leggauss(order)
    if unit_interval:
        xi, wi = 0.5 * (xi + 1.0), 0.5 * wi
    return torch.tensor(xi, dtype=dtype, device=device), torch.tensor(wi, dtype=dtype, device=device)


_TRI_RULES = {
    1: ([[1 / 3, 1 / 3]], [0.5], 1.0),
    3: ([[1 / 6, 1 / 6], [4 / 6, 1 / 6], [1 / 6, 4 / 6]], [1 / 6, 1 / 6, 1 / 6], 1.0),
    4: ([[1 / 3, 1 / 3], [0.6, 0.2], [0.2, 0.6], [0.2, 0.2]], [-27 / 96, 25 / 96, 25 / 96, 25 / 96], 0.5),
    6: ([[0.445948490915965, 0.445948490915965], [1 - 2 * 0.445948490915965, 0.445948490915965],
         [0.445948490915965, 1 - 2 * 0.445948490915965], [0.091576213509771, 0.091576213509771],
         [1 - 2 * 0.091576213509771, 0.091576213509771], [0.091576213509771, 1 - 2 * 0.091576213509771]],
        [0.111690794839005] * 3 + [0.054975871827661] * 3, 0.5),
    7: ([[1 / 3, 1 / 3], [0.0597158717, 0.4701420641], [0.4701420641, 0.0597158717], [0.4701420641, 0.4701420641],
         [0.7974269853, 0.1012865073], [0.1012865073, 0.7974269853], [0.1012865073, 0.1012865073]],
        [0.225, 0.1323941527, 0.1323941527, 0.1323941527, 0.1259391805, 0.1259391805, 0.1259391805], 0.5),
}


def triangle_gauss_points(order=1, device=None, dtype=torch.float32, fix_weights=False):
    """Points (r,s) and weights on the reference triangle, values of reference utils.py:13-81.
    fix_weights=True (correct-math switch, default off): orders 4 and 6 sum to the triangle area 0.5 instead of 0.25."""
    if device is None:
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    if order not in _TRI_RULES:
        raise NotImplementedError("Supported orders: 1, 3, 4, 6, 7")
    pts, wts, scale = _TRI_RULES[order]
    rs = torch.tensor(pts, dtype=dtype, device=device)
    w = torch.tensor(wts, dtype=dtype, device=device)
    if scale != 1.0 and not (fix_weights and order in (4, 6)):
        w = scale * w
    return rs, w


def test_gradients(model, loss_fn):
    """Smoke check of reference utils.py:83-97: both Parameters receive finite gradients."""
    loss = loss_fn(model)
    loss.backward()
    assert model.u_free.grad is not None
    assert not torch.isnan(model.u_free.grad).any()
    assert model.node_coords_free.grad is not None
    assert not torch.isnan(model.node_coords_free.grad).any()
    print("Gradient magnitudes:")
    print(f"u_free: {model.u_free.grad.norm()}")
    print(f"node_coords: {model.node_coords_free.grad.norm()}")


test_gradients.__test__ = False   # not a pytest test
