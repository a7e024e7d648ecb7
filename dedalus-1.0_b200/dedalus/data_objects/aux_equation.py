"""Auxiliary ODE advanced alongside the fields (reference: data_objects/aux_equation.py)."""


class AuxEquation(object):
    def __init__(self, RHS, kwargs, init_cond=0.):
        self._RHS = RHS
        self.kwargs = kwargs
        self.value = init_cond

    def RHS(self, value):
        return self._RHS(value, **self.kwargs)
