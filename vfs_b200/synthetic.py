"""Deterministic synthetic weights for benches and tools (there is no network for checkpoints): a parameter fill keyed
by state-dict NAME, so any module exposing the reference's state-dict names gets identical values from a seed alone.
``oracle.seeded_state_dict`` (test infrastructure) implements the same rule; tests/test_host_cpu.py keeps the two
in step."""
import zlib

import torch


def seeded_state_dict(module_or_sd, seed=0, bn_affine=True):
    """conv / linear weights ~ kaiming-normal(fan_out), BN gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1), running_mean ~
    N(0, 0.1), running_var ~ U(0.5, 1.5) (zero-init-residual makes the raw init degenerate: every block would be
    identity + ReLU)."""
    sd = module_or_sd if isinstance(module_or_sd, dict) else module_or_sd.state_dict()
    out = {}
    for name, t in sd.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2**31))
        if name.endswith('num_batches_tracked'):
            out[name] = torch.zeros_like(t)
        elif name.endswith('running_var'):
            out[name] = torch.rand(t.shape, generator=g) + 0.5
        elif name.endswith('running_mean'):
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        elif t.ndim == 1 and name.endswith('.weight'):
            out[name] = (torch.rand(t.shape, generator=g) + 0.5) if bn_affine else torch.ones_like(t)
        elif t.ndim == 1:
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        elif t.ndim >= 2:
            fan_out = t.shape[0] * (t[0, 0].numel() if t.ndim > 2 else 1)
            out[name] = torch.randn(t.shape, generator=g) * (2.0 / fan_out)**0.5
        else:
            out[name] = t.clone()
    return out
