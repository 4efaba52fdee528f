"""Drop-in for the reference module ``simplex`` (reference simplex.py): ``Simplex_CLASS`` with the same
methods, numpy-float64 return values and public state (``_perm``, ``_perm_grad_index3``), computed by
the sm_100a kernels of csrc/simplex.cu through the C ABI (no numba, no CPU path).
"""
from ctypes import c_int64

import numpy as np
import torch

from . import _lib

# ---- module constants of the reference (simplex.py:96-163) ----
DEFAULT_SEED = 3
GRADIENTS2 = np.array([5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5], dtype=np.int64)
_G3_ROWS = [(-11, 4, 4), (-4, 11, 4), (-4, 4, 11), (11, 4, 4), (4, 11, 4), (4, 4, 11),
            (-11, -4, 4), (-4, -11, 4), (-4, -4, 11), (11, -4, 4), (4, -11, 4), (4, -4, 11),
            (-11, 4, -4), (-4, 11, -4), (-4, 4, -11), (11, 4, -4), (4, 11, -4), (4, 4, -11),
            (-11, -4, -4), (-4, -11, -4), (-4, -4, -11), (11, -4, -4), (4, -11, -4), (4, -4, -11)]
GRADIENTS3 = np.array(_G3_ROWS, dtype=np.int64).reshape(-1)


def _gradients4():
    # 64 gradients: for each sign pattern (w,z,y,x bits), the four placements of the "3"
    out = []
    for s in range(16):
        sg = [-1 if (s >> k) & 1 else 1 for k in range(4)]
        for big in range(4):
            out.extend(sg[k] * (3 if k == big else 1) for k in range(4))
    return np.array(out, dtype=np.int64)


GRADIENTS4 = _gradients4()
STRETCH_CONSTANT2 = -0.211324865405187
SQUISH_CONSTANT2 = 0.366025403784439
STRETCH_CONSTANT3 = -1.0 / 6
SQUISH_CONSTANT3 = 1.0 / 3
STRETCH_CONSTANT4 = -0.138196601125011
SQUISH_CONSTANT4 = 0.309016994374947
NORM_CONSTANT2 = 47
NORM_CONSTANT3 = 103
NORM_CONSTANT4 = 30


def overflow(x):
    """Wrap a Python int to int64 (simplex.py:166-171)."""
    return c_int64(x).value


def _wrap64(seed):
    return ((int(seed) + (1 << 63)) % (1 << 64)) - (1 << 63)


def init_tables(seeds):
    """Host permutation tables for a list of seeds -> int16 array (n, 512): [perm | perm_grad_index3]."""
    seeds = np.asarray([_wrap64(s) for s in seeds], dtype=np.int64)
    tables = np.empty((seeds.size, 512), dtype=np.int16)
    _lib.call("anoddpm_simplex_init_batch", seeds.ctypes.data, int(seeds.size), tables.ctypes.data)
    return tables


def _init(seed=DEFAULT_SEED):
    """(perm, perm_grad_index3) as int64[256] arrays (simplex.py:174-192)."""
    t = init_tables([seed])[0].astype(np.int64)
    return t[:256].copy(), t[256:].copy()


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("anoddpm_b200.simplex needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Simplex_CLASS:
    """OpenSimplex noise generator, API of reference simplex.py:14-93."""

    def __init__(self):
        self.newSeed()

    def newSeed(self, seed=None):
        if not seed:  # simplex.py:20-21 (so newSeed(0) re-randomises, like the reference)
            seed = np.random.randint(-10000000000, 10000000000)
        self._load_seed(seed)

    def _load_seed(self, seed):
        self._tables = init_tables([seed])
        t = self._tables[0].astype(np.int64)
        self._perm, self._perm_grad_index3 = t[:256].copy(), t[256:].copy()
        self._dev_tables = None

    def _tables_dev(self):
        # rebuilt from the public arrays so that user edits of _perm/_perm_grad_index3 are honoured
        host = np.concatenate([np.asarray(self._perm), np.asarray(self._perm_grad_index3)]).astype(np.int16)
        return torch.from_numpy(host).to(_device())

    # ---- scalar / array evaluation ----
    def noise3array(self, x, y, z):
        dev = _device()
        X, Y, Z = (torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1), device=dev) for a in (x, y, z))
        out = torch.empty((Z.numel(), Y.numel(), X.numel()), dtype=torch.float64, device=dev)
        tab = self._tables_dev()
        _lib.call("anoddpm_simplex_grid3", _lib.ptr(out), _lib.ptr(X), X.numel(), _lib.ptr(Y), Y.numel(), _lib.ptr(Z),
                  Z.numel(), _lib.ptr(tab), _lib.stream())
        return out.cpu().numpy()

    def noise3(self, x, y, z):
        return float(self.noise3array([x], [y], [z])[0, 0, 0])

    def _grid2(self, x, y):
        dev = _device()
        X, Y = (torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64).reshape(-1), device=dev) for a in (x, y))
        out = torch.empty((Y.numel(), X.numel()), dtype=torch.float64, device=dev)
        tab = self._tables_dev()
        _lib.call("anoddpm_simplex_grid2", _lib.ptr(out), _lib.ptr(X), X.numel(), _lib.ptr(Y), Y.numel(),
                  _lib.ptr(tab), _lib.stream())
        return out.cpu().numpy()

    def noise2(self, x, y):
        return float(self._grid2([x], [y])[0, 0])

    def noise2array(self, x, y):
        """Reference _noise2a (simplex.py:312-318): flat[i*y.size + j] = noise2(x[j], y[i]), reshaped
        (x.size, y.size).  Identical to the reference for square grids; for non-square grids the
        reference's indexing quirk is reproduced with sequential write order."""
        x, y = np.asarray(x), np.asarray(y)
        m = self._grid2(x, y)  # m[i, j] = noise2(x[j], y[i])
        if x.size == y.size:
            return m
        flat = np.zeros(x.size * y.size, dtype=np.double)
        for i in range(y.size):
            flat[i * y.size:i * y.size + x.size] = m[i]
        return flat.reshape((x.size, y.size))

    # ---- fractal fields ----
    def _octaves(self, H, W, zvals, octaves, persistence, frequency):
        dev = _device()
        z = torch.as_tensor(np.ascontiguousarray(zvals, dtype=np.float64).reshape(-1), device=dev)
        out = torch.empty((z.numel(), H, W), dtype=torch.float64, device=dev)
        tab = self._tables_dev()
        _lib.call("anoddpm_simplex_octaves", _lib.ptr(out), 1, z.numel(), int(H), int(W), _lib.ptr(tab), _lib.ptr(z),
                  int(octaves), float(persistence), float(frequency), _lib.stream())
        return out.cpu().numpy()

    def rand_3d_octaves(self, shape, octaves=1, persistence=0.5, frequency=32):
        assert len(shape) == 3
        return self._octaves(shape[1], shape[2], np.arange(0, shape[0]), octaves, persistence, frequency)

    def rand_2d_octaves(self, shape, octaves=1, persistence=0.5, frequency=32):
        assert len(shape) == 2
        noise = np.zeros(shape)
        y, x = [np.arange(0, end) for end in shape]
        amplitude = 1
        for _ in range(octaves):
            noise += amplitude * self.noise2array(x / frequency, y / frequency)
            frequency /= 2
            amplitude *= persistence
        return noise

    def rand_3d_fixed_T_octaves(self, shape, T, octaves=1, persistence=0.5, frequency=32):
        assert len(shape) == 2
        noise = np.zeros((1, *shape))
        # like the reference (simplex.py:86-91) this only broadcasts for len(T) == 1
        noise += self._octaves(shape[0], shape[1], np.atleast_1d(T), octaves, persistence, frequency)
        return noise
