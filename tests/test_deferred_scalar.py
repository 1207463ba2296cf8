"""utils/deferred.py: the stable timestep is read from the device when it is first used; it must then
behave like the numpy scalar the eager code returned."""
import ctypes

import numpy as np

from sopht_mpi_b200.utils.deferred import DeferredScalar


def test_deferred_scalar_resolves_once_and_behaves_like_a_number():
    calls = []
    d = DeferredScalar(lambda: (calls.append(1), np.float32(0.125))[1])
    assert not d.resolved and not calls  # nothing is fetched until the value is needed
    assert abs(d - 0.125) <= 1e-9 and len(calls) == 1 and d.resolved
    assert 2 * d == 0.25 and d * 2 == 0.25 and d / 2 == 0.0625 and 1 / d == 8 and 1 + d == 1.125 and 1 - d == 0.875
    assert float(d) == 0.125 and ctypes.c_double(d).value == 0.125 and np.float64(d) == 0.125
    assert np.float32(2) * d == 0.25 and np.sqrt(d) == np.sqrt(np.float32(0.125))
    assert np.asarray(d).dtype == np.float32 and isinstance(d * 1, np.floating)
    assert min(d, 1.0) is d and max(d, 1.0) == 1.0 and d < 1 and d >= 0.125 and d == 0.125 and d != 1
    assert f"{d:.3f}" == "0.125" and str(d) == "0.125"
    t = 0.0
    t += d
    assert t == 0.125 and len(calls) == 1
    assert d.item() == 0.125 and d.dtype == np.float32 and d.astype(np.float64) == 0.125  # numpy scalar methods
    assert np.isclose(d, 0.125) and not np.isnan(d)
    import copy
    import math

    assert math.sqrt(d) == math.sqrt(0.125) and copy.copy(d) == 0.125
    e = DeferredScalar(lambda: np.float64(3.0))
    assert d * e == 0.375 and e ** 2 == 9 and -e == -3 and abs(-e) == 3 and 2 ** e == 8
