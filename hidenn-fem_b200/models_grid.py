"""1D and structured-2D models (placeholder until the grid kernels land)."""
import torch.nn as nn


class PiecewiseLinearShapeNN(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError


class StructuredShapeNN2D(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError
