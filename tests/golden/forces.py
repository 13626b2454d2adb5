"""User-force callables used by the golden fixtures (torch for the models, numpy for the closed-form oracle)."""
import numpy as np
import torch


def b_force_test(x):
    """Smooth function of the *reference* coordinates (loss.py:80 passes x_eval)."""
    return torch.stack([3.0e6 * (1.0 + x[:, 0]) * torch.cos(x[:, 1]), -2.0e6 * (x[:, 0] - 0.5 * x[:, 1] ** 2)], dim=1)


def t_force_test(x):
    """Traction that depends on the physical edge point (loss.py:106)."""
    return torch.stack([1.0e5 * (1.0 + 0.3 * x[:, 1] ** 2), 2.0e4 * torch.sin(3.0 * x[:, 1]) + 1.0e3 * x[:, 0]], dim=1)


def b_force_np(x):
    return np.stack([3.0e6 * (1.0 + x[:, 0]) * np.cos(x[:, 1]), -2.0e6 * (x[:, 0] - 0.5 * x[:, 1] ** 2)], axis=1)


def t_force_np(x):
    return np.stack([1.0e5 * (1.0 + 0.3 * x[..., 1] ** 2), 2.0e4 * np.sin(3.0 * x[..., 1]) + 1.0e3 * x[..., 0]], axis=-1)


def dt_dx_np(x):
    """d t_i / d x_j, shape [...,2,2]."""
    out = np.zeros(x.shape[:-1] + (2, 2), dtype=x.dtype)
    out[..., 0, 1] = 1.0e5 * 0.6 * x[..., 1]
    out[..., 1, 0] = 1.0e3
    out[..., 1, 1] = 2.0e4 * 3.0 * np.cos(3.0 * x[..., 1])
    return out
