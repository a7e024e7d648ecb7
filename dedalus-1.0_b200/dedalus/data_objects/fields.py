"""Scalar / vector / tensor field containers (reference: dedalus/data_objects/fields.py).

A field is a list of representation objects (its components).  `create_field_classes` binds
the representation class, shape and length into concrete ScalarField / VectorField /
TensorField classes, which is how a Physics object hands the representation plug-in to its
StateData (fields.py:35-57)."""
import weakref

import torch

from ..utils.logger import mylog
from ..utils.parallelism import reduce_max
from .representations import FourierRepresentation


def create_field_classes(representation, shape, length):
    bound = {"representation": representation, "shape": shape, "length": length}
    return {"TensorField": type("TensorField", (TensorFieldBase,), dict(bound)),
            "VectorField": type("VectorField", (VectorFieldBase,), dict(bound)),
            "ScalarField": type("ScalarField", (ScalarFieldBase,), dict(bound))}


class BaseField(object):
    """ncomp components of the bound representation."""

    def __init__(self, sd, ncomp):
        try:
            self.sd = weakref.proxy(sd)
        except TypeError:
            self.sd = sd
        self.ncomp = ncomp
        self.ndim = len(self.shape)
        self.components = [self.representation(self.sd, self.shape, self.length) for _ in range(ncomp)]
        self.ctrans = {}

    def _index(self, item):
        return self.ctrans[item] if isinstance(item, str) else item

    def __getitem__(self, item):
        return self.components[self._index(item)]

    def __setitem__(self, item, data):
        self.components[self._index(item)] = data

    def __iter__(self):
        return iter(enumerate(self.components))

    def zero(self, comp, space="kspace"):
        self.components[self._index(comp)][space] = 0.

    def zero_all(self, space="kspace"):
        for c in self.components:
            c[space] = 0.

    def save(self, group):
        group.attrs["representation"] = self.representation.__name__
        group.attrs["type"] = self.__class__.__name__
        for i, c in self:
            dset = group.create_dataset(str(i), tuple(int(n) for n in c.local_shape[c._curr_space]),
                                        dtype=c.dtype[c._curr_space])
            c.save(dset)

    def report_counts(self):
        for i, c in self:
            mylog.debug("component[%i] forward count = %i" % (i, c.fwd_count))
            mylog.debug("component[%i] rev count = %i" % (i, c.rev_count))


class TensorFieldBase(BaseField):
    def __init__(self, sd):
        BaseField.__init__(self, sd, sd.ndim ** 2)


class VectorFieldBase(BaseField):
    """One component per dimension; ctrans maps 'x','y','z' to 0,1,2."""

    def __init__(self, sd):
        BaseField.__init__(self, sd, sd.ndim)
        names = "xyz"[:self.ncomp]
        for i, n in enumerate(names):
            self.ctrans[n] = i
            self.ctrans[i] = n

    def max_square(self):
        """max over components and space of comp**2, reduced over ranks (fields.py:153-157)."""
        c2 = torch.stack([(c["xspace"] ** 2).max() for _, c in self])
        return reduce_max(c2, reduce_all=True)

    def l2norm(self):
        acc = torch.zeros_like(self.components[0]["xspace"])
        for _, c in self:
            acc += c["xspace"] ** 2
        return acc.sqrt()

    def div_free(self):
        """Remove the irrotational part: F_i -= k_i (k.F) / k^2 (fields.py:169-195).  Used for
        initial conditions; runs as a handful of tensor operations."""
        if not issubclass(self.representation, FourierRepresentation):
            raise NotImplementedError("Solenoidal projection not implemented for this representation.")
        mylog.debug("Performing solenoidal projection.")
        divF = 0
        for i, c in self:
            divF = divF + c.deriv(self.ctrans[i])
        k2 = self.components[0].k2(no_zero=True)
        for i, c in self:
            c["kspace"].sub_(-1j * c.k[self.ctrans[i]] * divF / k2)


class ScalarFieldBase(BaseField):
    """Single component; item and attribute access fall through to it, so a scalar field
    behaves like its representation (fields.py:197-245)."""

    def __init__(self, sd):
        BaseField.__init__(self, sd, 1)
        self.ctrans = {"": 0, 0: ""}

    def __getitem__(self, item):
        if isinstance(item, int) and item == 0:
            return self.components[0]
        return self.components[0][item]

    def __setitem__(self, item, data):
        if isinstance(item, int) and item == 0:
            self.components[0] = data
        else:
            self.components[0][item] = data

    def __getattr__(self, attr):
        if attr in ("components", "__setstate__", "__getstate__"):
            raise AttributeError(attr)
        return getattr(self.components[0], attr)

    def zero(self, space="kspace"):
        self.components[0][space] = 0.
