"""CPU, world_size 1 and 2 over gloo: the sharded step's host logic (row/column block layouts,
all-to-all transposes, partial theta-gradient reductions, boundary bookkeeping) reproduces the
unsharded oracle.  The numerical primitives are the CPU stand-in of tests/cpu_ops.py."""
import importlib
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

DT = torch.float64
CASES = [("poisson_2d-sin_add_cos", "poisson", "Matern52_Cos_1d", 1.0), ("allencahn_2d-mix-sincos", "allencahn", "SE_Cos_1d", 1.0),
         ("advection-sin", "advection", "Matern52_Cos_1d", 5.0)]


def _small_from(params, Q):
    s = torch.zeros(6 * Q + 2, dtype=DT)
    for a, key in enumerate(("kernel_paras_1", "kernel_paras_2")):
        for j, leaf in enumerate(("log-w", "log-ls", "freq")):
            s[(3 * a + j) * Q:(3 * a + j + 1) * Q] = params[key][leaf]
    s[6 * Q], s[6 * Q + 1] = params["log_tau"], params["log_v"]
    return s


def _run_case(equation, eq_name, kernel, beta, N1, N2, Q, steps, gs=True):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import gphm_oracle as O
    from cpu_ops import CpuOps
    D = importlib.import_module("gaussian-process-slover-for-high-freq-pde_b200.dist")
    p, _, _ = O.make_problem_2d(equation, kernel, N1, 2 * math.pi, beta=beta, M=8, N2=N2)
    params = O.state_S1(p, Q=Q, freq_scale=4.0)
    ops = CpuOps(kernel, eq_name, p.x, p.y, p.llk_weight, Q, beta, gs=gs)
    solver = D.ShardedSolver2D(kernel, eq_name, p.x.numpy(), p.y.numpy(), p.src.numpy(), p.bvals.numpy(), p.llk_weight,
                               1.0, beta, 1e-6, Q, 0.01, ops=ops)
    solver.set_state(params["U"], _small_from(params, Q))
    te, ge = O.loss_and_grad_efficient(p, params)
    terms, gU_r, gs = solver.value_and_grad()
    r0, h = solver.rank * solver.h, solver.h
    want_terms = [te["loss"], te["logdet1"], te["logdet2"], te["quad"], te["bgap"], te["eqgap"],
                  float(ge["log_tau"]), float(ge["log_v"])]
    for got, want in zip(terms.tolist(), want_terms):
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (got, want)
    wantU = ge["U"][r0:r0 + h]
    assert float((gU_r - wantU).norm()) <= 1e-9 * float(wantU.norm())
    assert float((gs - _small_from(ge, Q)).norm()) <= 1e-9 * float(_small_from(ge, Q).norm())
    ost = O.adam_init(params)
    for _ in range(steps):
        params, ost, _ = O.step(p, params, ost, 0.01, "efficient")
        solver.step()
    U = solver.gather_U()
    assert float((U - params["U"]).abs().max()) <= 1e-8
    assert float((solver.small - _small_from(params, Q)).abs().max()) <= 1e-8
    assert int(solver.count) == steps
    # step_host on host copies of this rank's params (plain semantics on the CPU stand-in): the same update as step()
    names = ("U", "small", "mU", "vU", "msmall", "vsmall", "count")
    saved = {k: getattr(solver, k).clone() for k in names}
    solver.step()
    want = {k: getattr(solver, k).clone() for k in names}
    for k in names:
        getattr(solver, k).copy_(saved[k])
    hU, hs, hloss = saved["U"].clone(), saved["small"].clone(), torch.zeros(1, dtype=DT)
    solver.step_host(hU, hs, hloss)
    assert all(torch.equal(getattr(solver, k), want[k]) for k in names)
    assert torch.equal(hU, want["U"]) and torch.equal(hs, want["small"]) and float(hloss) == float(solver.last_loss())
    return True


def _worker(rank, world, port, case):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _run_case(*case)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


# gs = True: the all-FFT step's host logic (4 exchanges, transposed column blocks); False: the general one (7 exchanges)
@pytest.mark.parametrize("gs", [True, False])
@pytest.mark.parametrize("equation,eq_name,kernel,beta", CASES)
def test_sharded_step_world1(equation, eq_name, kernel, beta, gs):
    assert _run_case(equation, eq_name, kernel, beta, 24, 20, 4, 2, gs)


@pytest.mark.parametrize("gs", [True, False])
@pytest.mark.parametrize("equation,eq_name,kernel,beta", CASES)
def test_sharded_step_world2_gloo(equation, eq_name, kernel, beta, gs):
    mp.spawn(_worker, args=(2, _free_port(), (equation, eq_name, kernel, beta, 24, 20, 4, 2, gs)), nprocs=2, join=True)


@pytest.mark.parametrize("gs", [True, False])
def test_sharded_step_world4_gloo(gs):
    mp.spawn(_worker, args=(4, _free_port(), ("poisson_2d-sin_add_cos", "poisson", "Matern52_Cos_1d", 1.0, 24, 20, 4, 1, gs)),
             nprocs=4, join=True)


def test_local_boundary_covers_every_edge_point_once_per_edge():
    D = importlib.import_module("gaussian-process-slover-for-high-freq-pde_b200.dist")
    N1, N2, P = 12, 8, 4
    bvals = list(range(2 * N1 + 2 * N2))
    seen = []
    for r in range(P):
        idx, val, nseg0 = D.local_boundary(r, P, N1, N2, bvals)
        h = N1 // P
        assert len(set(idx[:nseg0])) == nseg0 and len(set(idx[nseg0:])) == len(idx) - nseg0
        seen += [(r * h * N2 + i, v) for i, v in zip(idx, val)]
    assert sorted(v for _, v in seen) == bvals                     # every boundary value used exactly once
    corners = [0, N2 - 1, (N1 - 1) * N2, N1 * N2 - 1]
    flat = [i for i, _ in seen]
    assert all(flat.count(c) == 2 for c in corners)                # corners appear in two edges (reference hstack)
