"""GPU: the product path (reference-named classes -> C-ABI -> sm_100a kernels) against golden vectors produced by
EXECUTING the reference's own unmodified solver sources (tests/golden/make_ref_exec_golden.py, "g" size set:
40 x 33 grids, N = 60 in 1-D, Q = 6): loss and every gradient leaf within north_star's 1e-6 relative, two Adam
steps, the prediction - for every kernel class and every equation family.  (Runs last: the file name sorts after
the other GPU tests.)"""
import math
import os

import numpy as np
import pytest
import torch

from helpers import tree_flatten
from test_gpu_solver import TOL, make_1d, make_2d

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_exec.npz"))
TAGS = sorted({"|".join(k.split("|")[:3]) for k in GOLD.files if k.startswith("g")})
N1, N2, N1D, Q, FS, LR, M_TEST = 40, 33, 60, 6, 5.0, 0.01, 7


def _tree(tag, prefix, like):
    if isinstance(like, dict):
        return {k: _tree(tag, prefix + k + "/", v) for k, v in like.items()}
    return torch.as_tensor(GOLD[tag + "|" + prefix[:-1]], dtype=torch.float64)


@pytest.mark.parametrize("tag", TAGS)
def test_product_matches_executed_reference(gphm, oracle, tag):
    O = oracle
    dim, eq, kname = tag.split("|")
    two = dim.endswith("2d")
    if two:
        adv = eq.startswith("advection")
        beta = float(GOLD[tag + "|beta"])
        p, model, xte, _ = make_2d(gphm, O, eq, kname, N1, N2, Q, FS, 1.0 if adv else 2 * math.pi, beta=beta,
                                   llk=500.0 if adv else 200.0, M=M_TEST)
        assert np.array_equal(p.src.numpy().reshape(N1, N2), GOLD[tag + "|src"])        # same inputs as the reference run
        like = O.state_S1(p, Q=Q, freq_scale=FS)
    else:
        p, model, xte, _ = make_1d(gphm, O, eq, kname, N1D, Q, FS, 2 * math.pi)
        assert np.array_equal(p.src.numpy(), GOLD[tag + "|src"])
        like = O.init_params_1d(N1D, Q, FS)
    params = _tree(tag, "params0/", like)
    want_loss = float(GOLD[tag + "|loss"])
    loss, grads = model.value_and_grad(params)
    assert abs(float(loss) - want_loss) <= TOL * abs(want_loss)
    want = dict(tree_flatten(_tree(tag, "grad/", like)))
    for path, g in tree_flatten(grads):
        w = want[path]
        assert float((g.reshape(-1) - w.reshape(-1)).norm()) <= TOL * float(w.norm()) + 1e-300, (path, float(g.norm()), float(w.norm()))
    gparams, gst = params, model.core.init_opt_state(params)
    for k in range(2):
        gparams, gst, l = model.step(gparams, gst)
        assert abs(float(l) - float(GOLD[tag + "|step_losses"][k])) <= TOL * abs(float(l))
    ukey = "U" if two else "u"
    wantU = torch.as_tensor(GOLD[tag + "|params2/" + ukey])
    # Adam's first updates are +-lr*sign(g): the field agrees to ~1e-7 absolute (leaves with rounding-level gradients do not)
    assert float((gparams[ukey].cpu().reshape(-1) - wantU.reshape(-1)).abs().max()) <= 1e-6 * max(1.0, float(wantU.abs().max()))
    model.core.raise_on_bad_status()
