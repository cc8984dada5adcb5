"""TEST INFRASTRUCTURE: optax 0.1.4 `adam(learning_rate)` / `apply_updates` on torch pytrees (third-party, not in
the reference tree; published algorithm: scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0) with bias correction,
then scale(-lr)).  Pinned by the reference's own golden runs (tests/test_oracle_golden.py)."""
import torch


def _map(f, *trees):
    if isinstance(trees[0], dict):
        return {k: _map(f, *[t[k] for t in trees]) for k in trees[0]}
    return f(*trees)


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float64)


class _Adam(object):
    def __init__(self, lr, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    def init(self, params):
        z = lambda p: torch.zeros_like(_t(p))
        return {"count": 0, "mu": _map(z, params), "nu": _map(z, params)}

    def update(self, grads, state, params=None):
        c = state["count"] + 1
        mu = _map(lambda m, g: self.b1 * m + (1 - self.b1) * g, state["mu"], grads)
        nu = _map(lambda v, g: self.b2 * v + (1 - self.b2) * g * g, state["nu"], grads)
        upd = _map(lambda m, v: -self.lr * (m / (1 - self.b1 ** c)) / (torch.sqrt(v / (1 - self.b2 ** c)) + self.eps), mu, nu)
        return upd, {"count": c, "mu": mu, "nu": nu}


def adam(learning_rate):
    return _Adam(learning_rate)


def apply_updates(params, updates):
    return _map(lambda p, u: _t(p) + u, params, updates)
