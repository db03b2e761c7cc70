"""Stand-in for the (absent) `numba_progress` package, used ONLY by the oracle harness.

TEST INFRASTRUCTURE. The reference's element loop takes a progress-bar argument in its
njit signature (fem/physics/edm/optimized_assembly.py:66-75); this stub supplies a no-op
object of an opaque numba type whose `.update(n)` compiles to nothing.
"""
from numba import types
from numba.extending import (typeof_impl, register_model, models, unbox, NativeValue,
                             overload_method)


class _PBType(types.Opaque):
    def __init__(self):
        super().__init__(name="ProgressBarStubType")


ProgressBarType = _PBType()
register_model(_PBType)(models.OpaqueModel)


class ProgressBar:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, n=1):
        pass


@typeof_impl.register(ProgressBar)
def _typeof_pb(val, c):
    return ProgressBarType


@unbox(_PBType)
def _unbox_pb(typ, obj, c):
    return NativeValue(c.context.get_dummy_value())


@overload_method(_PBType, "update")
def _pb_update(self, n):
    def impl(self, n):
        return None
    return impl
