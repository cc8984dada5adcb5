"""CPU: the C-ABI library loads and exports every symbol include/gphm.h declares (no compute
calls - there is no GPU here), and the host-side logic (configs, problem setup, result formats,
pytree packing) behaves like the reference's."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gphm.h")).read()
    return sorted(set(re.findall(r"GPHM_API\s+[\w\s\*]+?\b(gphm_\w+)\s*\(", src)))


def test_header_symbols_are_exported(gphm):
    lib = gphm._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libgphm.so does not export " + name
    assert sorted(gphm._lib.EXPORTED_SYMBOLS) == declared, "ctypes table and header disagree"
    assert lib.gphm_version() == 100


def test_argument_validation_without_gpu(gphm):
    """Usage errors are reported through status codes + gphm_last_error, never by aborting."""
    lib = gphm._lib.load()
    d = gphm._lib.ProblemDesc()
    d.dim, d.kernel_id, d.eq_type, d.n1, d.n2, d.Q, d.nb = 2, 9, 0, 8, 8, 4, 32
    assert lib.gphm_workspace_bytes(ctypes.byref(d)) == 0
    assert b"Invalid Kernel" in lib.gphm_last_error()
    d.kernel_id = 1
    nbytes = lib.gphm_workspace_bytes(ctypes.byref(d))
    assert nbytes > 0
    d.nb = 7
    assert lib.gphm_workspace_bytes(ctypes.byref(d)) == 0
    d.nb, d.eq_type, d.dim, d.n2 = 2, 2, 1, 1
    assert lib.gphm_workspace_bytes(ctypes.byref(d)) == 0          # advection needs dim == 2
    assert lib.gphm_gram(0, 0, None, 4, None, 4, None, 3, 0.0, None, None) == -1
    with pytest.raises(Exception, match="Invalid Kernel"):
        gphm._lib.check(lib.gphm_kappa_pairs(7, 0, 1, 1, 1, 1, 3, 1, None), "x")


def test_workspace_size_scales(gphm):
    lib = gphm._lib.load()
    d = gphm._lib.ProblemDesc()
    d.dim, d.kernel_id, d.eq_type, d.Q = 2, 1, 0, 30
    d.n1 = d.n2 = 4096
    d.nb = 4 * 4096
    big = lib.gphm_workspace_bytes(ctypes.byref(d))
    assert 25 * 4096 * 4096 * 8 < big < 40 * 4096 * 4096 * 8       # ~3.5-5.4 GB, far below 180 GB HBM
    d.n1 = d.n2 = 400
    d.nb = 1600
    assert lib.gphm_workspace_bytes(ctypes.byref(d)) < big / 50


def test_configs_match_reference_yaml_keys(gphm):
    cfg = gphm.configs.load_config("poisson_2d-sin_add_cos", config_dir="/nonexistent")
    assert cfg["Q"] == 30 and cfg["lr"] == 0.01 and cfg["llk_weight"] == 200 and cfg["freq_scale"] == 20
    assert cfg["N_col"] == 200 and cfg["scale"] == "2pi" and cfg["logdet"] is True and cfg["tol"] == -1
    adv = gphm.configs.load_config("advection-sin", config_dir="/nonexistent")
    assert adv["beta"] == 200 and adv["llk_weight"] == 500 and adv["freq_scale"] == 40 and adv["scale"] == "1"
    assert set(gphm.configs.CONFIGS) >= set(gphm.model_GP_solver_1d.EQUATIONS) | {"advection-sin"}
    c = gphm.model_GP_solver_2d.make_config("poisson_2d-sin_sin", "Matern52_Cos_1d", 100, config_dir="/nonexistent")
    assert c["scale"] == 2 * np.pi and c["nepoch"] == 100 and c["other_paras"] == "-x-2pi-Ncol-400"
    assert c["kernel"] is gphm.Matern52_Cos_1d
    with pytest.raises(Exception, match="Invalid Kernel"):
        gphm.model_GP_solver_2d.make_config("poisson_2d-sin_sin", "RBF", 1, config_dir="/nonexistent")
    with pytest.raises(AssertionError):
        gphm.model_GP_solver_2d.make_config("poisson_9d", "SE_1d", 1, config_dir="/nonexistent")


def test_yaml_override(gphm, tmp_path):
    (tmp_path / "poisson_2d-sin_sin.yaml").write_text("Q: 7\nscale: '1'\nN_col: 50\nother_paras: '-x-1'\nfreq_scale: 3\n")
    c = gphm.model_GP_solver_2d.make_config("poisson_2d-sin_sin", "SE_1d", 5, config_dir=str(tmp_path))
    assert c["Q"] == 7 and c["scale"] == 1.0 and c["other_paras"] == "-x-1-Ncol-50"


def test_problem_setup_matches_oracle(gphm, oracle):
    O = oracle
    cfg = gphm.model_GP_solver_2d.make_config("allencahn_2d-mix-sincos", "SE_Cos_1d", 10, config_dir="/nonexistent")
    cfg["N_col"] = 37
    bvals, X_col, src, X_test, u_test = gphm.model_GP_solver_2d.build_problem(cfg, M=21)
    p, xt, ut = O.make_problem_2d("allencahn_2d-mix-sincos", "SE_Cos_1d", 37, 1.0, M=21)
    assert np.allclose(src, p.src.numpy(), rtol=1e-12, atol=1e-9)
    assert np.allclose(bvals, p.bvals.numpy()) and np.allclose(u_test, ut.numpy())
    cfg1 = gphm.model_GP_solver_2d.make_config("allencahn_1d-sin_cos", "SE_Cos_1d", 10,
                                               allowed=gphm.model_GP_solver_1d.EQUATIONS, config_dir="/nonexistent")
    Xind, y, X_col1, src1, X_test1, Y_test1 = gphm.model_GP_solver_1d.build_problem(cfg1)
    p1, xte1, yte1 = O.make_problem_1d("allencahn_1d-sin_cos", "SE_Cos_1d", 400, 2 * math.pi)
    assert np.allclose(src1, p1.src.numpy(), rtol=1e-12, atol=1e-7) and np.allclose(y, p1.yb.numpy())
    adv = gphm.model_GP_solver_2d.make_config("advection-sin", "SE_Cos_1d", 10, allowed=["advection-sin"],
                                              config_dir="/nonexistent")
    adv["N_col"] = 25
    b2, Xc2, src2, _, _ = gphm.model_GP_solver_advection.build_problem(adv, M=11)
    pa, _, _ = O.make_problem_2d("advection-sin", "SE_Cos_1d", 25, 1.0, beta=200.0, M=11)
    assert np.allclose(src2, pa.src.numpy(), atol=1e-9) and np.allclose(b2, pa.bvals.numpy())


def test_result_formats(gphm, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    tp = {"equation": "poisson_2d-sin_sin", "nepoch": 100, "Q": 30, "llk_weight": 200, "num_u_trick": 1, "lr": 0.01,
          "freq_scale": 20, "logdet": True, "other_paras": "-x-2pi-Ncol-400", "kernel": gphm.Matern52_Cos_1d}

    class M:
        cov_func = gphm.Matern52_Cos_1d()
        params = {"U": torch.zeros(2, 2), "log_tau": torch.tensor(0.5)}
    # same file name as the fixture the reference ships
    assert gphm.utils.get_save_name(tp) == "llk_weight-200.0-nu-1-Q-30-epoch-100-lr-0.0100-freqscale=20-logdet-1-x-2pi-Ncol-400"
    path = gphm.utils.store_model(M, {"loss_list": [1.0]}, tp)
    assert path == "result_log/poisson_2d-sin_sin/kernel_Matern52_Cos_1d/epoch_100/Q30/" + gphm.utils.get_save_name(tp) + ".pkl"
    import pickle
    params, log_dict, trick = pickle.load(open(path, "rb"))
    assert params["U"].shape == (2, 2) and trick["kernel"] == "Matern52_Cos_1d"
    log = gphm.utils.wrirte_log(M, {"mean": 0.4676, "std": 0.0, "used_time": 9.0589, "avg_time": 9.0589,
                                    "stop_epoch_mean": 100, "err_list": [0.46758844]}, tp)
    lines = open(log).read().splitlines()
    assert lines[0] == "llk_weight-200.0--nu-1-Q-30-epoch-100-lr-0.0100-freqscale=20-logdet-1-x-2pi-Ncol-400"
    assert lines[1] == "err_mean: 0.4676, err_std: 0.0000, used_time: 9.0589, avg_time: 9.0589, avg_epochs 100 "


def test_no_product_import_of_oracle():
    """The shipped package must never reach into oracle/ (test infrastructure)."""
    pkg = os.path.join(ROOT, "gaussian-process-slover-for-high-freq-pde_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f
