"""``freud.order.Steinhardt`` on the GPU path (reference ``freud/order.py:375-642``)."""

import numpy as np

from .locality import _computed, _ext, _PairCompute


class Steinhardt(_PairCompute):
    def __init__(self, l, average=False, wl=False, weighted=False, wl_normalize=False):  # noqa: E741
        ls = [int(v) for v in np.atleast_1d(l)]
        if any(v < 0 for v in ls):
            raise ValueError("l must be a non-negative integer.")
        self._scalar_l = np.ndim(l) == 0
        self._cpp_obj = _ext()._order.Steinhardt(ls, bool(average), bool(wl), bool(weighted), bool(wl_normalize))

    average = property(lambda self: self._cpp_obj.isAverage())
    wl = property(lambda self: self._cpp_obj.isWl())
    weighted = property(lambda self: self._cpp_obj.isWeighted())
    wl_normalize = property(lambda self: self._cpp_obj.isWlNormalized())

    @property
    def l(self):  # noqa: E743
        ls = self._cpp_obj.getL()
        return ls[0] if self._scalar_l else ls

    def compute(self, system, neighbors=None):
        nq, nlist, qargs, _ = self._preprocess_arguments(system, None, neighbors)
        self._cpp_obj.compute(nlist, nq._cpp_obj, qargs)
        self._called_compute = True
        return self

    @_computed
    def order(self):
        o = self._cpp_obj.getOrder()
        return o[0] if self._scalar_l else o

    @_computed
    def particle_order(self):
        a = self._cpp_obj.getParticleOrder()
        return a[:, 0] if self._scalar_l else a

    ql = _computed(lambda self: self._cpp_obj.getQl()[:, 0] if self._scalar_l else self._cpp_obj.getQl())

    @_computed
    def particle_harmonics(self):
        q = self._cpp_obj.getQlm()
        return q[0] if self._scalar_l else q
