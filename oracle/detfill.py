"""Deterministic, construction-order-independent weight fill.  TEST INFRASTRUCTURE ONLY.

The golden generator (tests/golden/make_golden.py) applies this to the REFERENCE modules and
the tests apply it to the drop-in modules; both then hold bit-identical weights without any
state_dict having to be committed (the reference cannot travel to the GPU box).  Every tensor
is drawn from a generator seeded by crc32(key), so the values depend only on the state_dict
key and shape, never on module construction order or on the global RNG.
"""
import zlib

import torch


def _gen(key, salt):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (salt * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


@torch.no_grad()
def fill_state_dict_(module, salt=0, gain=1.0):
    """In-place fill of every parameter and buffer of ``module``; returns module."""
    sd = module.state_dict()
    for key, t in sd.items():
        g = _gen(key, salt)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t.zero_()
        elif leaf == "running_mean":
            t.copy_(torch.randn(t.shape, generator=g) * 0.05)
        elif leaf == "running_var":
            t.copy_(torch.rand(t.shape, generator=g) * 0.5 + 0.75)
        elif t.dim() == 1 and leaf == "weight":
            # BN scale or PReLU slope (both 1-D 'weight'): keep in a tame positive range
            t.copy_(torch.rand(t.shape, generator=g) * 0.5 + 0.25)
        elif leaf == "bias":
            t.copy_(torch.randn(t.shape, generator=g) * 0.02)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            t.copy_(torch.randn(t.shape, generator=g) * (gain / fan_in ** 0.5))
        else:
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
    module.load_state_dict(sd)
    return module


def det_tensor(name, shape, scale=1.0, dtype=torch.float32):
    """Deterministic N(0, scale) tensor keyed by ``name``."""
    return (torch.randn(shape, generator=_gen(name, 17)) * scale).to(dtype)


def det_labels(name, n, num_classes):
    return torch.randint(0, num_classes, (n,), generator=_gen(name, 23), dtype=torch.int64)
