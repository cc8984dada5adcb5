"""GPU: the ensemble path (CUDA-graph replay of every member's gphm_step on several streams) gives
each member exactly what stepping it alone gives, and the members match the oracle."""
import importlib
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

E = importlib.import_module("gaussian-process-slover-for-high-freq-pde_b200.ensemble")


def _config(G, equation, N, Q=6, kernel="Matern52_Cos_1d"):
    tp = G.configs.load_config(equation, "/nonexistent")
    tp.update(equation=equation, kernel=kernel, Q=Q, N_col=N, scale=2 * math.pi if tp["scale"] == "2pi" else 1.0)
    return tp


@pytest.mark.parametrize("equation,N", [("poisson_2d-sin_add_cos", 48), ("poisson_1d-single_sin", 64), ("advection-sin", 40),
                                        ("allencahn_1d-sin_cos", 50)])
@pytest.mark.parametrize("graph", [True, False])
def test_ensemble_matches_members_stepped_alone(gphm, equation, N, graph):
    tp = _config(gphm, equation, N)
    members = E.ensemble_members(3, (5, 8))
    ens, idx = E.build_ensemble(tp, members, streams=4, graph=graph)
    assert idx == list(range(6)) and len(ens) == 6
    solo, _ = E.build_ensemble(tp, members, streams=1, graph=False)
    steps = 4
    ens.step(steps)
    for m, st in zip(solo.models, solo.states):                 # one member after the other, plain launches
        for _ in range(steps):
            m.core.step_inplace(st, m.lr)
    torch.cuda.synchronize()
    ens.raise_on_bad_status()
    for a, b in zip(ens.states, solo.states):                   # same kernels on the same inputs: bitwise
        assert torch.equal(a.U, b.U) and torch.equal(a.small, b.small) and torch.equal(a.terms, b.terms)
        assert int(a.count) == steps
    losses = ens.losses()
    assert losses.shape == (6,) and bool(torch.isfinite(losses).all())
    assert len(set(losses.tolist())) == 6                       # members really differ
    err = ens.errors()
    assert err.shape == (6,) and bool(torch.isfinite(err).all())


def test_ensemble_member_matches_oracle(gphm, oracle):
    """Member (seed 1, fs 8) of a 2-D Poisson ensemble, 2 steps, against the oracle from the same init."""
    O = oracle
    N, Q = 40, 6
    tp = _config(gphm, "poisson_2d-sin_add_cos", N, Q)
    members = [(0, 5), (1, 8), (2, 5)]
    ens, _ = E.build_ensemble(tp, members, streams=2)
    p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", "Matern52_Cos_1d", N, 2 * math.pi, M=300)
    init = E.member_init(ens.models[1], 1, 8)
    to_t = lambda t: {k: to_t(v) for k, v in t.items()} if isinstance(t, dict) else torch.as_tensor(np.asarray(t), dtype=torch.float64)
    params = to_t(init)
    ost = O.adam_init(params)
    for _ in range(2):
        params, ost, info = O.step(p, params, ost, tp["lr"], "efficient")
    ens.step(2)
    got = ens.params(1)
    assert abs(float(ens.losses()[1]) - info["loss"]) <= 1e-6 * abs(info["loss"])
    assert float((got["U"].cpu() - params["U"]).abs().max()) <= 1e-7
    for k in ("kernel_paras_1", "kernel_paras_2"):
        for leaf in ("log-w", "log-ls", "freq"):
            assert float((got[k][leaf].cpu() - params[k][leaf]).abs().max()) <= 1e-6


def test_ensemble_sharded_partition_runs_every_member_once(gphm):
    tp = _config(gphm, "poisson_1d-single_sin", 64)
    members = E.ensemble_members(2, (5, 8, 11))
    whole, _ = E.build_ensemble(tp, members, streams=2)
    whole.step(2)
    rows = []
    for r in range(4):                                          # the four "ranks" one after the other on this GPU
        part, idx = E.build_ensemble(tp, members, rank=r, world=4, streams=2)
        assert len(part) == len(idx) and len(idx) in (1, 2)
        part.step(2)
        rows.append(part.losses())
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(rows), whole.losses())
