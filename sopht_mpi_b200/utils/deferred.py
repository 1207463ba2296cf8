"""A number whose value is fetched from the device when it is first USED.

``UnboundedFlowSimulator3D.compute_stable_timestep`` returns one: the reference's examples call it at
the top of every step (``dt = flow_sim.compute_stable_timestep(); interactor(); interactor.time_step(dt);
flow_sim.time_step(dt)``, e.g. ``examples/3d_examples/FlowPastSphereCase/flow_past_sphere_case.py:120-140``),
and returning a plain float there makes the host wait for the GPU to drain before it can enqueue
anything of the new step.  The interaction does not need dt, so with the deferred value its kernels are
queued behind the previous step and the wait (``float(dt)`` inside ``time_step``) finds work in flight.
The object resolves itself on any arithmetic, comparison, conversion or formatting, and yields exactly
the numpy scalar the eager code path would have returned; ``SB200_EAGER_DT=1`` switches the deferral off.
(The 2D simulator stays eager: its 512 x 256 step is bound by host launches, there is no GPU idle time to
win back, and the deferral's own Python costs 0.02 ms of a 0.235 ms step: measured 0.255 ms.)
"""
import operator

import numpy as np


def _value(x):
    return x._get() if isinstance(x, DeferredScalar) else x


def _binary(op):
    def forward(self, other):
        return op(self._get(), _value(other))

    def reflected(self, other):
        return op(_value(other), self._get())

    return forward, reflected


class DeferredScalar:
    __slots__ = ("_fn", "_val")

    def __init__(self, resolve):
        self._fn, self._val = resolve, None

    @property
    def resolved(self):
        return self._fn is None

    def _get(self):
        if self._fn is not None:
            self._val = self._fn()
            self._fn = None
        return self._val

    def __getattr__(self, name):
        # everything else a numpy scalar offers (.item(), .dtype, .astype ...) comes from the value itself
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self._get(), name)

    # conversions
    def __float__(self):
        return float(self._get())

    def __int__(self):
        return int(self._get())

    def __bool__(self):
        return bool(self._get())

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._get(), dtype=dtype)

    def __repr__(self):
        return repr(self._get())

    def __str__(self):
        return str(self._get())

    def __format__(self, spec):
        return format(self._get(), spec)

    def __hash__(self):
        return hash(self._get())

    def __neg__(self):
        return -self._get()

    def __pos__(self):
        return +self._get()

    def __abs__(self):
        return abs(self._get())

    def __round__(self, n=None):
        return round(self._get(), n) if n is not None else round(self._get())

    __add__, __radd__ = _binary(operator.add)
    __sub__, __rsub__ = _binary(operator.sub)
    __mul__, __rmul__ = _binary(operator.mul)
    __truediv__, __rtruediv__ = _binary(operator.truediv)
    __floordiv__, __rfloordiv__ = _binary(operator.floordiv)
    __mod__, __rmod__ = _binary(operator.mod)
    __pow__, __rpow__ = _binary(operator.pow)
    __lt__ = _binary(operator.lt)[0]
    __le__ = _binary(operator.le)[0]
    __gt__ = _binary(operator.gt)[0]
    __ge__ = _binary(operator.ge)[0]
    __eq__ = _binary(operator.eq)[0]
    __ne__ = _binary(operator.ne)[0]
