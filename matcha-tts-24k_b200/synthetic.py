"""Deterministic synthetic weights and inputs for parity tests and the benchmark.

There is no network for checkpoints or corpora, so weights are random-init in the shape of
the reference's ``Decoder.initialize_weights`` (reference decoder.py:341-357): kaiming-normal
(relu gain) matrices / conv kernels, torch-default uniform for the ConvTranspose1d, and -- so
that the padding-dependent terms of the path are exercised -- N(0, 0.1) on every 1-D
parameter (biases, norm affine, SnakeBeta alpha/beta), as BASELINE.md section 3 prescribes.

Every tensor is drawn from its own CPU generator seeded by crc32(parameter name) ^ seed, so
the values depend only on (name, shape, seed): the reference modules, the oracle and the
product container all receive bit-identical weights regardless of construction order.
"""
from __future__ import annotations

import math
import zlib

import torch


def named_seed_tensor(name: str, shape, seed: int = 1234, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) >= 2:
        fan_in = shape[1] * math.prod(shape[2:])
        if len(shape) == 3 and shape[2] == 4:  # ConvTranspose1d [Cin, Cout, 4]: torch default init
            bound = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        else:
            w = torch.randn(shape, generator=g, dtype=torch.float64) * math.sqrt(2.0 / fan_in)
    else:
        w = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        if leaf == "weight":  # GroupNorm / LayerNorm scale
            w = w + 1.0
    return w.to(dtype)


def fill_named_seed(module: torch.nn.Module, seed: int = 1234) -> None:
    """Overwrite every parameter of ``module`` in place (names taken from its state_dict)."""
    with torch.no_grad():
        for name, p in module.state_dict().items():
            p.copy_(named_seed_tensor(name, p.shape, seed).to(p.dtype))


def sequence_mask(lengths: torch.Tensor, max_length: int) -> torch.Tensor:
    """Prefix mask, as the reference builds it (matcha/utils/model.py:7-9)."""
    return torch.arange(max_length, device=lengths.device).unsqueeze(0) < lengths.unsqueeze(1)


def make_inputs(lengths, n_feats: int = 100, seed: int = 0, T: int | None = None, device="cpu"):
    """mu = randn*mask, z = randn, mask = sequence_mask(L, T), T = 2*ceil(max L / 2)."""
    lengths = torch.as_tensor(lengths, dtype=torch.int64)
    if T is None:
        T = int(2 * ((int(lengths.max()) + 1) // 2))
    g = torch.Generator().manual_seed(seed)
    B = lengths.numel()
    mask = sequence_mask(lengths, T).unsqueeze(1).float()
    mu = torch.randn(B, n_feats, T, generator=g) * mask
    z = torch.randn(B, n_feats, T, generator=g)
    return mu.to(device), mask.to(device), z.to(device), lengths


# BASELINE.json configs -> concrete synthetic workloads.
PROD = dict(channels=(384, 384), dropout=0.05, attention_head_dim=64, n_blocks=2, num_mid_blocks=2, num_heads=6)
DEFAULT = dict(channels=(320, 320), dropout=0.05, attention_head_dim=64, n_blocks=2, num_mid_blocks=2, num_heads=5)


def config_lengths(name: str):
    if name == "cfg1":
        return [150]
    if name == "cfg2":
        return [938] * 32
    if name == "cfg3":
        g = torch.Generator().manual_seed(0)
        return torch.randint(188, 1126, (256,), generator=g).tolist()
    if name == "cfg4":
        return [2812] * 16
    if name == "cfg5":
        return [938] * 64
    raise KeyError(name)


def algorithmic_flops(lengths, channels: int, nfe: int) -> float:
    """per utterance per NFE  L*(274 C^2 + 1800 C) + 24 C L^2  (2 FLOP / MAC)."""
    c = channels
    return float(nfe) * sum(L * (274 * c * c + 1800 * c) + 24 * c * L * L for L in lengths)
