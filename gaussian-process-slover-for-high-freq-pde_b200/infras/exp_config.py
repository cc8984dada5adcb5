"""CLI surface of the reference (infras/exp_config.py:33-55): `equation`, `kernel`, `nepoch`.
As in the reference, `nepoch` defaults to 1000000 and therefore always overrides the YAML value
(model_GP_solver_2d.py:490-491)."""


class ExpConfig(object):
    equation = None
    kernel = None
    nepoch = 1000000
    config_name = "Exp Config"

    def parse(self, kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
        print("=" * 33)
        print("*", self.config_name)
        print("-" * 33)
        for k in ("equation", "kernel", "nepoch"):
            print("-", k, ":", getattr(self, k))
        print("=" * 33)
        return self
