"""ctypes front-end of oracle/simplex_oracle.c (TEST INFRASTRUCTURE; see oracle/__init__.py).

Mirrors the numpy-facing surface of the reference ``Simplex_CLASS``
(/root/reference/simplex.py:14-93) so parity tests read like calls into the reference.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build():
    """Compile the C oracle in-tree (gcc only; seconds)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_simplex.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_simplex.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_simplex_init.argtypes = [ctypes.c_int64, _i64p, _i64p]
        L.oracle_simplex_init.restype = None
        L.oracle_noise3.argtypes = [ctypes.c_double] * 3 + [_i64p, _i64p]
        L.oracle_noise3.restype = ctypes.c_double
        L.oracle_noise2.argtypes = [ctypes.c_double] * 2 + [_i64p]
        L.oracle_noise2.restype = ctypes.c_double
        L.oracle_noise3_grid.argtypes = [_f64p, ctypes.c_int, _f64p, ctypes.c_int, _f64p, ctypes.c_int,
                                         _i64p, _i64p, _f64p]
        L.oracle_noise3_grid.restype = None
        L.oracle_octaves.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_double, ctypes.c_double, _i64p, _i64p, _f64p]
        L.oracle_octaves.restype = None
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def wrap_seed(seed):
    """Python int -> the int64 the reference's ``overflow`` sees after the first LCG step
    (simplex.py:166-171): only the value mod 2**64 matters."""
    return ((int(seed) + (1 << 63)) % (1 << 64)) - (1 << 63)


def init_tables(seed):
    perm = np.zeros(256, dtype=np.int64)
    pgi3 = np.zeros(256, dtype=np.int64)
    lib().oracle_simplex_init(wrap_seed(seed), _p(perm, _i64p), _p(pgi3, _i64p))
    return perm, pgi3


class OracleSimplex:
    """Same call surface as the reference Simplex_CLASS (numpy float64 results)."""

    def __init__(self, seed=3, threads=1):
        self.threads = threads
        self.newSeed(seed)

    def newSeed(self, seed=None):
        if not seed:  # simplex.py:20-21
            seed = np.random.randint(-10000000000, 10000000000)
        self._perm, self._perm_grad_index3 = init_tables(seed)

    def noise3(self, x, y, z):
        return lib().oracle_noise3(x, y, z, _p(self._perm, _i64p), _p(self._perm_grad_index3, _i64p))

    def noise2(self, x, y):
        return lib().oracle_noise2(x, y, _p(self._perm, _i64p))

    def noise3array(self, x, y, z):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        out = np.empty((z.size, y.size, x.size), dtype=np.float64)
        lib().oracle_noise3_grid(_p(x, _f64p), x.size, _p(y, _f64p), y.size, _p(z, _f64p), z.size,
                                 _p(self._perm, _i64p), _p(self._perm_grad_index3, _i64p), _p(out, _f64p))
        return out

    def _octaves(self, H, W, zvals, octaves, persistence, frequency):
        zvals = np.ascontiguousarray(zvals, dtype=np.float64)
        out = np.empty((zvals.size, H, W), dtype=np.float64)
        L = lib()
        args = (_p(zvals, _f64p), zvals.size, int(octaves), float(persistence), float(frequency),
                _p(self._perm, _i64p), _p(self._perm_grad_index3, _i64p), _p(out, _f64p))
        nt = max(1, min(self.threads, H))
        if nt == 1:
            L.oracle_octaves(W, H, 0, H, *args)
        else:
            bounds = np.linspace(0, H, nt + 1).astype(int)
            with ThreadPoolExecutor(nt) as ex:  # ctypes drops the GIL during the call
                list(ex.map(lambda i: L.oracle_octaves(W, H, int(bounds[i]), int(bounds[i + 1]), *args), range(nt)))
        return out

    def rand_3d_octaves(self, shape, octaves=1, persistence=0.5, frequency=32):
        assert len(shape) == 3
        return self._octaves(shape[1], shape[2], np.arange(shape[0]), octaves, persistence, frequency)

    def rand_3d_fixed_T_octaves(self, shape, T, octaves=1, persistence=0.5, frequency=32):
        assert len(shape) == 2
        return self._octaves(shape[0], shape[1], np.atleast_1d(T), octaves, persistence, frequency)
