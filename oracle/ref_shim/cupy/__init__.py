"""Import shim (test infrastructure only): lets the reference's CuPy-based block
functions run on CPU by aliasing the `cupy` namespace to NumPy.  Used only by
oracle/make_golden.py inside the build container, never by the product."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import random, linalg, fft  # noqa: F401

ndarray = _np.ndarray
float32 = _np.float32
bool_ = _np.bool_


def asnumpy(a, *args, **kwargs):
    return _np.asarray(a)


def get_array_module(*args):
    return _np
