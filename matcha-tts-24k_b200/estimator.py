"""Parameter container for the U-Net estimator.

Holds the estimator's weights under exactly the reference's state-dict names so that
``MatchaTTSInfer.load_state_dict(ckpt["state_dict"], strict=False)`` (reference
matcha/inference.py:194) fills them, including the ``ff._orig_mod.`` infix created by
``torch.compile(self.ff)`` (reference transformer.py:219).  It contains no PyTorch arithmetic:
``forward`` hands one estimator evaluation to the CUDA library.  The name/shape table below is
the single source of truth shared with ``csrc/`` (the library looks weights up by these names).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn


@dataclass(frozen=True)
class EstimatorConfig:
    in_channels: int          # 2*n_feats (+ spk_dim when upstream-style spks conditioning is on)
    out_channels: int         # n_feats
    channels: int             # C (both U-Net levels; the reference is always 2-level)
    n_heads: int
    head_dim: int
    n_blocks: int
    n_mid_blocks: int
    dropout: float = 0.05

    @property
    def temb_dim(self) -> int:
        return 4 * self.channels

    @property
    def inner(self) -> int:
        return self.n_heads * self.head_dim


def config_from_decoder_params(in_channels, out_channels, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                               n_blocks=1, num_mid_blocks=2, num_heads=4, down_block_type="transformer",
                               mid_block_type="transformer", up_block_type="transformer") -> EstimatorConfig:
    """Same keyword surface as the reference ``Decoder.__init__`` (decoder.py:203-216)."""
    for kind in (down_block_type, mid_block_type, up_block_type):
        if kind != "transformer":
            raise ValueError(f"Unknown block type {kind}")  # reference decoder.py:337; conformer is unused
    channels = tuple(channels)
    if len(channels) != 2 or channels[0] != channels[1]:
        raise ValueError(f"this build supports the reference's 2-level U-Net with equal widths, got channels={channels}")
    return EstimatorConfig(int(in_channels), int(out_channels), int(channels[0]), int(num_heads),
                           int(attention_head_dim), int(n_blocks), int(num_mid_blocks), float(dropout))


def _resnet(prefix, cin, cout, temb):
    return [
        (f"{prefix}.mlp.1.weight", (cout, temb)), (f"{prefix}.mlp.1.bias", (cout,)),
        (f"{prefix}.block1.block.0.weight", (cout, cin, 3)), (f"{prefix}.block1.block.0.bias", (cout,)),
        (f"{prefix}.block1.block.1.weight", (cout,)), (f"{prefix}.block1.block.1.bias", (cout,)),
        (f"{prefix}.block2.block.0.weight", (cout, cout, 3)), (f"{prefix}.block2.block.0.bias", (cout,)),
        (f"{prefix}.block2.block.1.weight", (cout,)), (f"{prefix}.block2.block.1.bias", (cout,)),
        (f"{prefix}.res_conv.weight", (cout, cin, 1)), (f"{prefix}.res_conv.bias", (cout,)),
    ]


def _transformer(prefix, c, inner):
    return [
        (f"{prefix}.norm1.weight", (c,)), (f"{prefix}.norm1.bias", (c,)),
        (f"{prefix}.attn1.to_q.weight", (inner, c)), (f"{prefix}.attn1.to_k.weight", (inner, c)),
        (f"{prefix}.attn1.to_v.weight", (inner, c)),
        (f"{prefix}.attn1.to_out.0.weight", (c, inner)), (f"{prefix}.attn1.to_out.0.bias", (c,)),
        (f"{prefix}.norm3.weight", (c,)), (f"{prefix}.norm3.bias", (c,)),
        (f"{prefix}.ff._orig_mod.net.0.proj.weight", (4 * c, c)), (f"{prefix}.ff._orig_mod.net.0.proj.bias", (4 * c,)),
        (f"{prefix}.ff._orig_mod.net.0.alpha", (4 * c,)), (f"{prefix}.ff._orig_mod.net.0.beta", (4 * c,)),
        (f"{prefix}.ff._orig_mod.net.2.weight", (c, 4 * c)), (f"{prefix}.ff._orig_mod.net.2.bias", (c,)),
    ]


def weight_spec(cfg: EstimatorConfig):
    """Ordered (name, shape) list == the reference estimator's ``state_dict()`` keys."""
    C, T4 = cfg.channels, cfg.temb_dim
    spec = [("time_mlp.linear_1.weight", (T4, cfg.in_channels)), ("time_mlp.linear_1.bias", (T4,)),
            ("time_mlp.linear_2.weight", (T4, T4)), ("time_mlp.linear_2.bias", (T4,))]
    for i, cin in enumerate((cfg.in_channels, C)):
        spec += _resnet(f"down_blocks.{i}.0", cin, C, T4)
        for j in range(cfg.n_blocks):
            spec += _transformer(f"down_blocks.{i}.1.{j}", C, cfg.inner)
        tail = f"down_blocks.{i}.2.conv" if i == 0 else f"down_blocks.{i}.2"
        spec += [(f"{tail}.weight", (C, C, 3)), (f"{tail}.bias", (C,))]
    for i in range(cfg.n_mid_blocks):
        spec += _resnet(f"mid_blocks.{i}.0", C, C, T4)
        for j in range(cfg.n_blocks):
            spec += _transformer(f"mid_blocks.{i}.1.{j}", C, cfg.inner)
    for i in range(2):
        spec += _resnet(f"up_blocks.{i}.0", 2 * C, C, T4)
        for j in range(cfg.n_blocks):
            spec += _transformer(f"up_blocks.{i}.1.{j}", C, cfg.inner)
        if i == 0:
            spec += [("up_blocks.0.2.conv.weight", (C, C, 4)), ("up_blocks.0.2.conv.bias", (C,))]
        else:
            spec += [("up_blocks.1.2.weight", (C, C, 3)), ("up_blocks.1.2.bias", (C,))]
    spec += [("final_block.block.0.weight", (C, C, 3)), ("final_block.block.0.bias", (C,)),
             ("final_block.block.1.weight", (C,)), ("final_block.block.1.bias", (C,)),
             ("final_proj.weight", (cfg.out_channels, C, 1)), ("final_proj.bias", (cfg.out_channels,))]
    return spec


class _Node(nn.Module):
    """Anonymous container; exists only to give parameters their dotted names."""


class EstimatorWeights(nn.Module):
    """The ``CFM.estimator`` object: reference-named parameters, CUDA-backed ``forward``."""

    def __init__(self, cfg: EstimatorConfig):
        super().__init__()
        self.cfg = cfg
        self.in_channels, self.out_channels = cfg.in_channels, cfg.out_channels
        self._owner = None  # set by CFM; a list so the CFM is not registered as a submodule
        for name, shape in weight_spec(cfg):
            node, parts = self, name.split(".")
            for part in parts[:-1]:
                if part not in node._modules:
                    node.add_module(part, _Node())
                node = node._modules[part]
            node.register_parameter(parts[-1], nn.Parameter(torch.empty(shape)))
        self.initialize_weights()

    def initialize_weights(self):
        """Same distribution as reference decoder.py:341-357 (kaiming-normal relu gain, zero biases,
        unit norm scales, torch-default uniform for the ConvTranspose1d which the reference leaves alone)."""
        with torch.no_grad():
            for name, p in self.named_parameters():
                leaf = name.rsplit(".", 1)[-1]
                if p.ndim >= 2:
                    if name.startswith("up_blocks.0.2.conv"):
                        fan_in = p.shape[1] * p.shape[2]
                        nn.init.uniform_(p, -fan_in ** -0.5, fan_in ** -0.5)
                    else:
                        nn.init.kaiming_normal_(p, nonlinearity="relu")
                elif name == "up_blocks.0.2.conv.bias":
                    fan_in = self.cfg.channels * 4
                    nn.init.uniform_(p, -fan_in ** -0.5, fan_in ** -0.5)
                elif leaf == "weight":
                    nn.init.ones_(p)
                else:
                    nn.init.zeros_(p)

    def forward(self, x, mask, mu, t, spks=None):
        """One estimator evaluation v = f(x, mask, mu, t) on the GPU (reference decoder.py:359-426).
        Inference only: there is no autograd path through the CUDA library."""
        if self._owner is None:
            raise RuntimeError("EstimatorWeights is not attached to a CFM")
        return self._owner[0]._estimator_call(x, mask, mu, t, spks)
