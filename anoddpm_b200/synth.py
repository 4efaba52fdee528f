"""Deterministic synthetic weights for benchmarks, smoke tests and golden generation (there is no network for checkpoints).

``fill_state_dict`` is key-addressed: every tensor is drawn from a generator seeded by its state_dict key, so the same
values can be regenerated on any machine, and the reference's zero-initialised tensors (conv2 / proj_out / head)
carry signal.  No arithmetic of the hot path lives here.
"""
import zlib

import torch


def fill_state_dict(sd, seed=0):
    """weight ~ U(-1,1) * fan_in^-1/2 ; norm weight ~ 1 + 0.2 U ; bias ~ 0.2 U"""
    out = {}
    for k in sd.keys():
        v = sd[k]
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) & 0x7FFFFFFF)
        u = torch.rand(v.shape, generator=g, dtype=torch.float32) * 2 - 1
        if v.dim() >= 2:
            fan_in = v[0].numel()
            u = u * fan_in ** -0.5
        elif k.endswith("weight"):
            u = 1 + 0.2 * u
        else:
            u = 0.2 * u
        out[k] = u
    return out
