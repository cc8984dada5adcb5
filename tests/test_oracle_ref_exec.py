"""CPU: the oracle against golden vectors produced by EXECUTING the reference's own unmodified solver sources
(tests/golden/make_ref_exec_golden.py: /root/reference/code/*.py run with a torch-backed stand-in for the jax /
optax API, because JAX cannot be installed in the build container).  Every kernel class x every equation family:
loss, every gradient leaf, two Adam steps and the prediction.  This is the pin for what the two shipped result
logs do not cover (SE_Cos_1d, Matern52_1d, SE_1d, Allen-Cahn, advection)."""
import math
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_exec.npz"))
TAGS = sorted({"|".join(k.split("|")[:3]) for k in GOLD.files if not k.startswith("x1d")})
SIZES = {"s": (12, 10, 20, 4), "g": (40, 33, 60, 6)}                    # N1, N2, N (1-D), Q - as in the generator
FS, LR, M_TEST = 5.0, 0.01, 7


def _tree(tag, prefix, like):
    if isinstance(like, dict):
        return {k: _tree(tag, prefix + k + "/", v) for k, v in like.items()}
    return torch.as_tensor(GOLD[tag + "|" + prefix[:-1]], dtype=torch.float64)


def _rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64).reshape(-1), torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm()) / max(float(b.norm()), 1e-300)


def _problem(O, tag):
    """(problem, test grid, pytree skeleton) rebuilt from the generator's recipe; the stored data must match."""
    dim, eq, kname = tag.split("|")
    N1, N2, N1D, Q = SIZES[dim[0]]
    if dim[1:] == "2d":
        adv = eq.startswith("advection")
        beta = float(GOLD[tag + "|beta"])
        p, (xt, yt), ut = O.make_problem_2d(eq, kname, N1, 1.0 if adv else 2 * math.pi, beta=beta, M=M_TEST, N2=N2,
                                            llk_weight=500.0 if adv else 200.0)
        assert np.array_equal(p.src.numpy().reshape(N1, N2), GOLD[tag + "|src"]) and np.array_equal(p.bvals.numpy(), GOLD[tag + "|bvals"])
        return p, (xt, yt), O.state_S1(p, Q=Q, freq_scale=FS)
    p, xte, _ = O.make_problem_1d(eq, kname, N1D, 2 * math.pi, M=M_TEST)
    assert np.array_equal(p.src.numpy(), GOLD[tag + "|src"]) and np.array_equal(p.yb.numpy(), GOLD[tag + "|yb"])
    return p, xte, O.init_params_1d(N1D, Q, FS)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_executed_reference(oracle, tag):
    O = oracle
    p, xte, like = _problem(O, tag)
    two = tag.split("|")[0].endswith("2d")
    params = _tree(tag, "params0/", like)                     # the state the reference was evaluated at
    want_loss = float(GOLD[tag + "|loss"])
    want_grad = _tree(tag, "grad/", like)
    # tiny set: rounding level.  "g" set (N = 40 x 33, cond(K) ~ 1e7): LU solves and differently ordered autodiff agree to
    # ~1e-7 in the gradients (observed <= 1e-7; bound = north_star's 1e-6), Adam's first steps amplify leaves whose
    # gradient is at rounding level (observed 5e-8 absolute after two steps)
    small = tag[0] == "s"
    tol_loss, tol_grad, tol_par, tol_pred = (1e-10, 1e-8, 1e-9, 1e-8) if small else (1e-9, 1e-6, 1e-6, 1e-7)
    for name, fn in (("literal", O.loss_and_grad_literal), ("efficient", O.loss_and_grad_efficient)):
        terms, grads = fn(p, params)
        assert abs(terms["loss"] - want_loss) <= tol_loss * abs(want_loss), (name, terms["loss"], want_loss)
        for (ka, a), (kb, b) in zip(O.flatten(grads), O.flatten(want_grad)):
            assert ka == kb
            if float(b.norm()) == 0.0:
                assert float(a.norm()) == 0.0, (name, ka)         # freq of the kernels without a cosine factor
            else:
                assert _rel(a, b) <= tol_grad, (name, ka, _rel(a, b))
    st, pr = O.adam_init(params), params
    for k in range(2):
        pr, st, info = O.step(p, pr, st, LR, "efficient")
        assert abs(info["loss"] - float(GOLD[tag + "|step_losses"][k])) <= tol_loss * abs(info["loss"])
    want2 = _tree(tag, "params2/", like)
    for (ka, a), (kb, b) in zip(O.flatten(pr), O.flatten(want2)):
        assert float((a.reshape(-1) - b.reshape(-1)).abs().max()) <= tol_par, ka           # Adam steps are ~lr: absolute
    pred = O.preds_2d(p, pr, xte[0], xte[1]) if two else O.preds_1d(p, pr, xte)
    assert float((pred.reshape(-1) - torch.as_tensor(GOLD[tag + "|pred2"]).reshape(-1)).abs().max()) <= tol_pred


XTAGS = sorted({"|".join(k.split("|")[:3]) for k in GOLD.files if k.startswith("x1d")})


@pytest.mark.parametrize("tag", XTAGS)
def test_oracle_extra_gp_matches_executed_reference(oracle, tag):
    """Second stage of GP_solver_1d_extra (model_GP_solver_1d_extra.py:107-152) executed by the reference itself:
    loss_extra, its gradient, two step_extra updates (oracle side: autograd of loss_extra_literal + the oracle's Adam)."""
    O = oracle
    _, eq, _ = tag.split("|")
    N, Qg = SIZES["g"][2], SIZES["g"][3]
    p, xte, _ = O.make_problem_1d(eq, "SE_Cos_1d", N, 2 * math.pi, M=M_TEST)
    assert np.array_equal(p.src.numpy(), GOLD[tag + "|src"])
    first = _tree(tag, "first/", O.init_params_1d(N, Qg, FS))
    like = {"u": 0, "kernel_paras": {"log-w": 0, "log-ls": 0}, "log_tau": 0, "log_v": 0}
    pe = _tree(tag, "params0/", like)

    def value_and_grad(params_extra):
        leaves = O._clone_leaves(params_extra, True)
        val = O.loss_extra_literal(p, "Matern52_1d", first, leaves)
        flat = O.flatten(leaves)
        grads = torch.autograd.grad(val, [t for _, t in flat])
        return float(val.detach()), dict(zip([k for k, _ in flat], grads))

    val, grads = value_and_grad(pe)
    want = float(GOLD[tag + "|loss"])
    assert abs(val - want) <= 1e-10 * abs(want)
    for path, w in O.flatten(_tree(tag, "grad/", like)):
        assert _rel(grads[path], w) <= 1e-6, (path, _rel(grads[path], w))
    st, cur = O.adam_init(pe), pe
    for k in range(2):
        v, g = value_and_grad(cur)
        assert abs(v - float(GOLD[tag + "|step_losses"][k])) <= 1e-9 * abs(v)
        gtree = {"u": g["u"], "kernel_paras": {"log-w": g["kernel_paras/log-w"], "log-ls": g["kernel_paras/log-ls"]},
                 "log_tau": g["log_tau"], "log_v": g["log_v"]}
        cur, st = O.adam_update(cur, gtree, st, LR)
    for (ka, a), (kb, b) in zip(O.flatten(cur), O.flatten(_tree(tag, "params2/", like))):
        assert float((a.reshape(-1) - b.reshape(-1)).abs().max()) <= 1e-6, ka
