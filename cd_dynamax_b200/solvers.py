"""Solver tokens accepted in `diffeqsolve_settings["solver"]` (mirrors the diffrax constructors the reference passes,
src/utils/diffrax_utils.py:121-127; `RK4` is the classical tableau named by the north star, absent from diffrax)."""


class _Solver:
    def __init__(self, *args, **kwargs):
        pass

    def __repr__(self):
        return f"{type(self).__name__}()"


class Euler(_Solver):
    pass


class Heun(_Solver):
    pass


class Midpoint(_Solver):
    pass


class Ralston(_Solver):
    pass


class Bosh3(_Solver):
    pass


class RK4(_Solver):
    pass


class Dopri5(_Solver):
    pass


class ConstantStepSize:
    def __init__(self, *args, **kwargs):
        pass
