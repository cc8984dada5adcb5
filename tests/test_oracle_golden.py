"""CPU: the oracle replays the reference's own golden runs (the pin for every parity claim) and
its two formulations agree.  Fixtures: tests/golden/*.npz (made by tests/golden/make_golden.py from
the reference's code/result_log/**.pkl)."""
import math
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _loglike(loss):
    return math.log(loss) if loss > 1 else loss     # model_GP_solver_2d.py:311


def test_golden_1d_trajectory(oracle):
    O = oracle
    g = np.load(os.path.join(GOLD, "poisson_1d_single_sin_matern52cos_e100.npz"))
    p, xte, yte = O.make_problem_1d("poisson_1d-single_sin", "Matern52_Cos_1d", 400, 2 * math.pi)
    params = O.init_params_1d(400, 30, 20)
    st = O.adam_init(params)
    for i in range(100):
        params, st, terms = O.step(p, params, st, 0.01, "efficient")
        if i % 5 == 0:
            j = i // 5
            assert abs(_loglike(terms["loss"]) - g["log_loss_list"][j]) <= 1e-8 * abs(g["log_loss_list"][j]), (i,)
            err = O.rel_l2(O.preds_1d(p, params, xte), yte)
            assert abs(err - g["log_err_list"][j]) <= 1e-7 * g["log_err_list"][j], (i, err)
            assert np.allclose(torch.exp(params["kernel_paras"]["log-w"]).numpy(), g["log_w_list"][j], rtol=1e-6, atol=0)
            assert np.allclose(params["kernel_paras"]["freq"].numpy(), g["log_freq_list"][j], rtol=1e-6, atol=1e-9)
    assert np.linalg.norm(params["u"].numpy() - g["u"]) <= 1e-6 * np.linalg.norm(g["u"])
    assert abs(float(params["log_tau"]) - float(g["log_tau"])) < 1e-8
    assert abs(float(params["log_v"]) - float(g["log_v"])) < 1e-8


def test_golden_1d_first_steps_literal(oracle):
    """The literal (autograd / LU / slogdet) formulation on the same fixture, first checkpoints."""
    O = oracle
    g = np.load(os.path.join(GOLD, "poisson_1d_single_sin_matern52cos_e100.npz"))
    p, xte, yte = O.make_problem_1d("poisson_1d-single_sin", "Matern52_Cos_1d", 400, 2 * math.pi)
    params = O.init_params_1d(400, 30, 20)
    st = O.adam_init(params)
    for i in range(6):
        params, st, terms = O.step(p, params, st, 0.01, "literal")
        if i % 5 == 0:
            assert abs(_loglike(terms["loss"]) - g["log_loss_list"][i // 5]) <= 1e-9 * abs(g["log_loss_list"][i // 5])


def test_golden_2d_trajectory(oracle):
    O = oracle
    g = np.load(os.path.join(GOLD, "poisson_2d_sin_sin_matern52cos_e100.npz"))
    p, xt, ut = O.make_problem_2d("poisson_2d-sin_sin", "Matern52_Cos_1d", 400, 2 * math.pi)
    params = O.init_params_2d(400, 400, 30, 20)
    st = O.adam_init(params)
    for i in range(100):
        params, st, terms = O.step(p, params, st, 0.01, "efficient")
        if i % 5 == 0:
            j = i // 5
            tol = 1e-12 if i == 0 else (1e-7 if i <= 5 else 1e-4)       # trajectories are chaotic
            assert abs(_loglike(terms["loss"]) - g["log_loss_list"][j]) <= tol * abs(g["log_loss_list"][j]), (i,)
            if j in (0, 1, 2, 5, 10, 15, 19):           # the prediction is 5x a step's cost: a spread of checkpoints
                err = O.rel_l2(O.preds_2d(p, params, xt[0], xt[1]), ut)
                assert abs(err - g["log_err_list"][j]) <= (1e-7 if i <= 5 else 1e-3) * g["log_err_list"][j], (i, err)
    final = O.rel_l2(O.preds_2d(p, params, xt[0], xt[1]), ut)
    assert abs(final - 0.46758844) <= 0.05 * 0.46758844                   # log.txt:2-3


def test_golden_2d_step0_literal(oracle):
    O = oracle
    g = np.load(os.path.join(GOLD, "poisson_2d_sin_sin_matern52cos_e100.npz"))
    p, _, _ = O.make_problem_2d("poisson_2d-sin_sin", "Matern52_Cos_1d", 400, 2 * math.pi)
    terms, _ = O.loss_and_grad_literal(p, O.init_params_2d(400, 400, 30, 20))
    assert abs(math.log(terms["loss"]) - g["log_loss_list"][0]) < 1e-12


@pytest.mark.parametrize("kernel", ["SE_Cos_1d", "Matern52_Cos_1d", "Matern52_1d", "SE_1d"])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_closed_forms_match_autograd_of_kappa(oracle, kernel, order):
    """k, k', k'' closed forms == nested autograd of the literal scalar kappa (x1 != y1), and the
    theta partials == autograd of the closed forms."""
    O = oracle
    torch.manual_seed(0)
    Q = 5
    th = {"log-w": torch.randn(Q, dtype=O.DT) * 0.3, "log-ls": torch.randn(Q, dtype=O.DT) * 0.3,
          "freq": torch.rand(Q, dtype=O.DT) * 3}
    for x1v, y1v in [(0.3, 1.1), (1.7, 0.2), (0.05, 0.0)]:
        x1 = torch.tensor(x1v, dtype=O.DT, requires_grad=True)
        y1 = torch.tensor(y1v, dtype=O.DT)
        val = O.kappa_scalar(kernel, x1, y1, th)
        for _ in range(order):
            (val,) = torch.autograd.grad(val, x1, create_graph=True)
        G = O.gram(kernel, x1.detach().reshape(1), y1.reshape(1), th, order)
        assert abs(float(val) - float(G)) <= 1e-10 * max(1.0, abs(float(val)))
    d = torch.tensor([0.0, 0.013, 0.4, 2.5], dtype=O.DT)
    leaves = [th[k].clone().requires_grad_(True) for k in ("log-w", "log-ls", "freq")]
    term, parts = O.kernel_terms(kernel, d, leaves[0], leaves[1], leaves[2], order, True)
    for m in range(d.numel()):
        grads = torch.autograd.grad(term[m].sum(), leaves, retain_graph=True, allow_unused=True)
        for gr, pt in zip(grads, parts):
            gr = torch.zeros(Q, dtype=O.DT) if gr is None else gr
            assert torch.allclose(gr, pt[m].detach(), rtol=1e-10, atol=1e-12)


def test_second_derivative_diagonal_is_analytic(oracle):
    """sgn(0)=+1 convention: diag of the DD Gram = k''(0) = -sum w (a^2/3 + omega^2) for Matern*cos."""
    O = oracle
    th = {"log-w": torch.tensor([0.1, -0.4], dtype=O.DT), "log-ls": torch.tensor([0.2, -0.1], dtype=O.DT),
          "freq": torch.tensor([1.5, 7.0], dtype=O.DT)}
    x = torch.linspace(0, 1, 7, dtype=O.DT)
    D = O.gram("Matern52_Cos_1d", x, x, th, 2)
    w, a, om = torch.exp(th["log-w"]), math.sqrt(5) * torch.exp(th["log-ls"]), 2 * math.pi * th["freq"]
    assert torch.allclose(torch.diagonal(D), torch.full((7,), float(-(w * (a * a / 3 + om * om)).sum()), dtype=O.DT))
    D1 = O.gram("Matern52_Cos_1d", x, x, th, 1)
    assert torch.allclose(D1, -D1.T) and float(torch.diagonal(D1).abs().max()) == 0.0


CASES_2D = [("poisson_2d-sin_add_cos", "Matern52_Cos_1d", 1.0), ("allencahn_2d-mix-sincos", "SE_Cos_1d", 1.0),
            ("advection-sin", "SE_1d", 7.0), ("poisson_2d-sin_sin", "Matern52_1d", 1.0)]


@pytest.mark.parametrize("equation,kernel,beta", CASES_2D)
def test_efficient_equals_literal_2d(oracle, equation, kernel, beta):
    O = oracle
    p, _, _ = O.make_problem_2d(equation, kernel, 40, 2 * math.pi, beta=beta, N2=33)
    s = O.state_S1(p, Q=6, freq_scale=3.0)
    tl, gl = O.loss_and_grad_literal(p, s)
    te, ge = O.loss_and_grad_efficient(p, s)
    # LU (literal) vs Cholesky (efficient): agreement is limited by cond(K) ~ 1e7 for the plain SE
    # kernel, so the bound is the 1e-6 parity bound of north_star, not machine precision.
    for k in tl:
        assert abs(tl[k] - te[k]) <= 1e-6 * max(1.0, abs(tl[k])), k
    for (ka, a), (kb, b) in zip(O.flatten(gl), O.flatten(ge)):
        assert float((a - b).norm()) <= 1e-6 * max(float(a.norm()), 1e-30), ka


@pytest.mark.parametrize("equation,kernel", [("poisson_1d-sin_cos", "Matern52_Cos_1d"), ("allencahn_1d-single_sin", "SE_Cos_1d")])
def test_efficient_equals_literal_1d(oracle, equation, kernel):
    O = oracle
    p, _, _ = O.make_problem_1d(equation, kernel, 60, 2 * math.pi)
    params = O.init_params_1d(60, 6, 3.0)
    params["u"] = (0.5 * torch.sin(3 * p.x) + 0.1).reshape(-1, 1)
    params["log_tau"], params["log_v"] = torch.tensor(0.3, dtype=O.DT), torch.tensor(-0.2, dtype=O.DT)
    tl, gl = O.loss_and_grad_literal(p, params)
    te, ge = O.loss_and_grad_efficient(p, params)
    for k in tl:
        assert abs(tl[k] - te[k]) <= 1e-9 * max(1.0, abs(tl[k])), k
    for (ka, a), (kb, b) in zip(O.flatten(gl), O.flatten(ge)):
        assert float((a - b).norm()) <= 1e-8 * max(float(a.norm()), 1e-30), ka


def test_nonuniform_grid_general_path(oracle):
    from helpers import nonuniform_grid
    O = oracle
    p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", "Matern52_Cos_1d", 30, 2 * math.pi, N2=24)
    p.x = nonuniform_grid(30, 2 * math.pi, 1)
    p.y = nonuniform_grid(24, 2 * math.pi, 2)
    s = O.state_S1(p, Q=5, freq_scale=2.0)
    tl, gl = O.loss_and_grad_literal(p, s)
    te, ge = O.loss_and_grad_efficient(p, s)
    assert abs(tl["loss"] - te["loss"]) <= 1e-9 * abs(tl["loss"])
    for (ka, a), (kb, b) in zip(O.flatten(gl), O.flatten(ge)):
        assert float((a - b).norm()) <= 1e-8 * max(float(a.norm()), 1e-30), ka


def test_known_answer_S1_N200(oracle):
    """SURVEY App. G.4 row N=200."""
    O = oracle
    p, _, _ = O.make_problem_2d("poisson_2d-sin_add_cos", "Matern52_Cos_1d", 200, 2 * math.pi)
    te, ge = O.loss_and_grad_efficient(p, O.state_S1(p))
    assert abs(te["loss"] - 2.83539134746924e08) <= 1e-9 * 2.83539134746924e08
    assert abs(float(ge["U"].norm()) - 1.0290228319e08) <= 1e-8 * 1.0290228319e08
    assert abs(float(ge["log_tau"]) + 7.5164805751e04) <= 1e-8 * 7.5164805751e04
