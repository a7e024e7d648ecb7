"""StateData: the ordered collection of fields that makes up the state of a run
(reference: dedalus/data_objects/state_data.py:45-186)."""
from collections import OrderedDict

from ..utils.api import Timer
from ..utils.logger import mylog
from .fields import create_field_classes


class StateData(object):
    timer = Timer()

    def __init__(self, time, shape, length, field_class_dict, field_list=(), params=None):
        self.time = time
        self.shape = shape
        self.ndim = len(shape)
        self.length = length
        self._field_classes = field_class_dict
        self.parameters = params if params is not None else {}
        self.fields = OrderedDict()
        for name, ftype in field_list:
            self.add_field(name, ftype)

    def __getitem__(self, item):
        return self.fields[item]

    def __iter__(self):
        return iter(self.fields.items())

    def clone(self):
        """Empty StateData with the same grid and field classes (no fields)."""
        return self.__class__(self.time, self.shape, self.length, self._field_classes, params=self.parameters)

    def set_time(self, time):
        # shearing-box components follow the time: their wavenumbers drift with it (state_data.py:115-121)
        self.time = time
        for f in self._cached()[4]:
            for _, c in f:
                c._update_k()

    def add_field(self, name, fieldtype):
        if name in self.fields:
            raise ValueError("Field with this name already exists.")
        self.fields[name] = self._field_classes[fieldtype](self)
        self._cache = None

    # Per-step host overhead matters for the small, launch-bound grids (tens of microseconds of kernels per step): the
    # integrators and the fused RHS walk the components through these cached lists instead of the generators.  Valid because a
    # component's k-space buffer never changes identity (representations.py:144-149) and fields are only ever added.
    _cache = None

    def _cached(self):
        c = self._cache
        if c is None:
            comps = [comp for _, f in self.fields.items() for _, comp in f]
            ks = [comp._k for comp in comps]
            import ctypes as C
            arr = (C.c_void_p * len(ks))(*[t.data_ptr() for t in ks])
            dyn = [f for _, f in self.fields.items() if not f.representation._static_k]
            c = self._cache = (comps, ks, arr, C.cast(arr, C.c_void_p), dyn)
        return c

    def comp_list(self):
        """All components, insertion order (cached)."""
        return self._cached()[0]

    def components(self):
        """(field name, component index, representation) in insertion order."""
        for name, f in self.fields.items():
            for i, c in f:
                yield name, i, c

    def snapshot(self, root_grp):
        for name, f in self.fields.items():
            f.save(root_grp.create_group(name))

    def create_tmp_data(self, space):
        c = next(iter(self.fields.values()))[0]
        import torch
        return torch.zeros(tuple(int(n) for n in c.local_shape[space]), dtype=getattr(torch, c.dtype[space]),
                           device=c.kdata.device)

    def report_counts(self):
        for name, f in self.fields.items():
            mylog.debug("field %s" % name)
            f.report_counts()

    def __reduce__(self):
        state = {k: v for k, v in self.__dict__.items() if k not in ("fields", "_field_classes", "_cache", "_subs")}
        rep = next(iter(self._field_classes.values())).representation
        keys = [(n, f.__class__.__name__) for n, f in self.fields.items()]
        return (_rebuild_state, (self.__class__, state, rep, keys))


def _rebuild_state(cls, state, rep, keys):
    obj = cls(state["time"], state["shape"], state["length"], create_field_classes(rep, state["shape"], state["length"]))
    obj.__dict__.update(state)
    for name, ftype in keys:
        obj.add_field(name, ftype)
    return obj
