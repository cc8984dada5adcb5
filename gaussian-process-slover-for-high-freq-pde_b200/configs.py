"""Per-equation hyper-parameters: the same keys and values as the reference's
code/config/<equation>.yaml files, which `evals()` merges into `trick_paras`
(model_GP_solver_2d.py:476-491).  A user-supplied ./config/<equation>.yaml (the reference's
cwd-relative convention) overrides this table when present."""
import os

_COMMON = dict(num_u_trick=1, Q=30, lr=0.01, logdet=True, num_fold=1, tol=-1)


def _cfg(equation, llk_weight, freq_scale, N_col, scale, nepoch, **extra):
    c = dict(_COMMON, equation=equation, llk_weight=llk_weight, freq_scale=freq_scale, N_col=N_col, scale=scale,
             other_paras="-x-2pi" if scale == "2pi" else "-x-1", nepoch=nepoch)
    c.update(extra)
    return c


CONFIGS = {c["equation"]: c for c in [
    _cfg("poisson_1d-single_sin", 200, 20, 400, "2pi", 100000),
    _cfg("poisson_1d-mix_sin", 200, 30, 900, "1", 100000, change_point=0.5),
    _cfg("poisson_1d-sin_cos", 200, 20, 400, "2pi", 100000),
    _cfg("poisson_1d-x_time_sinx", 200, 50, 900, "2pi", 100000),
    _cfg("poisson_1d-x2_add_sinx", 200, 100, 400, "1", 1000000, change_point=0.01),
    _cfg("allencahn_1d-single_sin", 200, 20, 400, "2pi", 100000),
    _cfg("allencahn_1d-sin_cos", 200, 20, 400, "2pi", 100000),
    _cfg("poisson_2d-sin_sin", 200, 20, 400, "2pi", 100000),
    _cfg("poisson_2d-sin_add_cos", 200, 20, 200, "2pi", 1000000),
    _cfg("allencahn_2d-mix-sincos", 200, 30, 400, "1", 1000000),
    _cfg("advection-sin", 500, 40, 200, "1", 200000, beta=200),
]}


def load_config(equation, config_dir="./config"):
    """dict for `equation`; reads <config_dir>/<equation>.yaml if it exists (reference layout)."""
    path = os.path.join(config_dir, equation + ".yaml")
    if os.path.exists(path):
        import yaml
        with open(path, "r") as f:
            return yaml.safe_load(f)
    if equation not in CONFIGS:
        raise FileNotFoundError(path)
    return dict(CONFIGS[equation])
