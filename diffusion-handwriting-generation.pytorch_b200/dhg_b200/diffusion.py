"""Diffusion schedule and the two posterior updates (utils/nn.py:19-39,64-112),
computed on the GPU by the library's fused update kernel."""
import math

import torch

NUM_STEPS = 60


def explin(min_val, max_val, n):
    return torch.exp(torch.linspace(math.log(min_val), math.log(max_val), n))


def get_beta_set():
    """0.02 + explin(1e-5, 0.4, 60) -- same torch ops and order as utils/nn.py:19-39,
    so the schedule handed to the library is bit-identical to the reference's."""
    return 0.02 + explin(1e-5, 0.4, NUM_STEPS)


def get_alpha_bar(beta=None):
    beta = get_beta_set() if beta is None else beta
    return torch.cumprod(1 - beta, dim=0)  # inference.py:81
