"""Shared helpers for the parity tests (test infrastructure)."""
import math

import numpy as np
import torch

DT = torch.float64


def rel(a, b):
    a = torch.as_tensor(a, dtype=DT).detach().cpu().reshape(-1)
    b = torch.as_tensor(b, dtype=DT).detach().cpu().reshape(-1)
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def theta_state(Q, freq_scale, shift=0, seed=None):
    """Deterministic non-degenerate kernel parameters (the S1 recipe of SURVEY 8d)."""
    q = torch.arange(Q, dtype=DT)
    return {"log-w": math.log(1.0 / Q) - 0.05 * torch.cos(q + shift),
            "log-ls": 0.1 * torch.sin(q + shift),
            "freq": freq_scale * q / max(Q - 1, 1) + 0.05 * torch.sin(2 * (q + shift))}


def params_to_numpy(params):
    if isinstance(params, dict):
        return {k: params_to_numpy(v) for k, v in params.items()}
    return torch.as_tensor(params).detach().cpu().numpy()


def tree_flatten(tree, prefix=""):
    out = []
    for k in sorted(tree.keys()):
        v = tree[k]
        if isinstance(v, dict):
            out += tree_flatten(v, prefix + k + "/")
        else:
            out.append((prefix + k, torch.as_tensor(v, dtype=DT).detach().cpu()))
    return out


def nonuniform_grid(n, scale, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(0, 1, n, dtype=DT)
    h = 1.0 / (n - 1)
    x = x + (torch.rand(n, generator=g, dtype=DT) - 0.5) * 0.6 * h
    x[0], x[-1] = 0.0, 1.0
    return x * scale
