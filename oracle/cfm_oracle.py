"""CPU oracle for the CFM decode path of faltiska/Matcha-TTS-24k.

TEST INFRASTRUCTURE ONLY.  Nothing under ``matcha-tts-24k_b200/`` may import this
file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the timed baseline.

PARITY STATUS: *unpinned by the reference's own tests* -- the reference holds no
test, fixture or golden vector for this path.  What pins this file:
  * ``tests/golden/make_golden.py`` imports the reference's OWN ``decoder.py``,
    ``transformer.py`` and ``flow_matching.py`` from /root/reference (third-party
    ``diffusers`` / ``torchdiffeq`` / ``conformer`` replaced by the small shims in that
    script, because those packages are not installed and there is no network) and
    stores its outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this
    restatement against those vectors.
  * The third-party arithmetic itself (diffusers ``Attention``/``AttnProcessor2_0``,
    torchdiffeq fixed-grid solvers; both unpinned in requirements.txt:3,22) is restated
    from the published algorithms -- that part stays unpinned.

Plain PyTorch, runs in fp32 or fp64 on the CPU.  Module / parameter names follow the
reference so that a reference ``state_dict`` (keys ``estimator.*``) loads unchanged,
including the ``ff._orig_mod.`` infix that ``torch.compile(self.ff)`` creates
(reference matcha/models/components/transformer.py:219).

All ``ref:`` citations are relative to /root/reference/matcha/models/components/.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

SOLVER_NFE = {"euler": 1, "midpoint": 2, "heun3": 3, "rk4": 4}


# --------------------------------------------------------------------------- time embedding
def sinusoidal_embedding(t: torch.Tensor, dim: int, scale: float = 1000.0) -> torch.Tensor:
    """ref: decoder.py:20-29.  ``t`` is 0-dim or (B,); returns (B, dim) = [sin | cos]."""
    if t.ndim < 1:
        t = t.unsqueeze(0)
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device).float() * -step)
    arg = scale * t.unsqueeze(1) * freqs.unsqueeze(0)
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


class TimeMLP(nn.Module):
    """ref: decoder.py:75-119 (TimestepEmbedding with act_fn="silu", nothing optional set)."""

    def __init__(self, in_channels: int, hidden: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, hidden)
        self.linear_2 = nn.Linear(hidden, hidden)

    def forward(self, e):
        return self.linear_2(F.silu(self.linear_1(e)))


# --------------------------------------------------------------------------- conv blocks
class ConvNormAct(nn.Module):
    """ref: decoder.py:32-45 (Block1D): Mish(GroupNorm8(Conv1d_k3(x*mask))) * mask."""

    def __init__(self, cin: int, cout: int, groups: int = 8):
        super().__init__()
        self.block = nn.Sequential(nn.Conv1d(cin, cout, 3, padding=1), nn.GroupNorm(groups, cout), nn.Mish())

    def forward(self, x, mask):
        return self.block(x * mask) * mask


class ResBlock(nn.Module):
    """ref: decoder.py:48-63 (ResnetBlock1D)."""

    def __init__(self, cin: int, cout: int, temb_dim: int):
        super().__init__()
        self.mlp = nn.Sequential(nn.Mish(), nn.Linear(temb_dim, cout))
        self.block1 = ConvNormAct(cin, cout)
        self.block2 = ConvNormAct(cout, cout)
        self.res_conv = nn.Conv1d(cin, cout, 1)

    def forward(self, x, mask, temb):
        h = self.block1(x, mask)
        h = h + self.mlp(temb).unsqueeze(-1)  # added to padded frames too (decoder.py:60)
        h = self.block2(h, mask)
        return h + self.res_conv(x * mask)


class StridedDown(nn.Module):
    """ref: decoder.py:66-72 (Downsample1D): Conv1d k3 s2 p1."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv1d(c, c, 3, 2, 1)

    def forward(self, x):
        return self.conv(x)


class TransposedUp(nn.Module):
    """ref: decoder.py:122-160 (Upsample1D, use_conv_transpose=True): ConvTranspose1d k4 s2 p1."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.ConvTranspose1d(c, c, 4, 2, 1)

    def forward(self, x):
        return self.conv(x)


# --------------------------------------------------------------------------- transformer block
class SnakeBetaProj(nn.Module):
    """ref: transformer.py:14-77.  Linear then h + sin^2(h*e^alpha) / (e^beta + 1e-9)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.proj = nn.Linear(cin, cout)
        self.alpha = nn.Parameter(torch.zeros(cout))
        self.beta = nn.Parameter(torch.zeros(cout))

    def forward(self, x):
        h = self.proj(x)
        a, b = torch.exp(self.alpha), torch.exp(self.beta)
        return h + (1.0 / (b + 1e-9)) * torch.pow(torch.sin(h * a), 2)


class FeedForwardNet(nn.Module):
    """ref: transformer.py:80-120.  net = [SnakeBeta(dim->4dim), Dropout, Linear(4dim->dim)]."""

    def __init__(self, dim: int, dropout: float):
        super().__init__()
        self.net = nn.ModuleList([SnakeBetaProj(dim, 4 * dim), nn.Dropout(dropout), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        for layer in self.net:
            x = layer(x)
        return x


class _CompiledNameShim(nn.Module):
    """Reproduces the key infix of ``torch.compile(self.ff)`` (ref: transformer.py:219)
    without compiling anything: parameters live under ``_orig_mod``."""

    def __init__(self, inner: nn.Module):
        super().__init__()
        self._orig_mod = inner

    def forward(self, x):
        return self._orig_mod(x)


class SelfAttention(nn.Module):
    """Restatement of diffusers ``Attention`` + ``AttnProcessor2_0`` for the constructor
    arguments used at ref: transformer.py:180-188 (self-attention, bias=False, no norms).

    q/k/v projections are bias-free, ``to_out[0]`` has a bias, scale = dim_head**-0.5.
    The (B, T) *float* mask is broadcast to (B, H, 1, T) and handed to
    ``scaled_dot_product_attention`` as an ADDITIVE bias: valid keys get +1, padded keys
    +0 -- padded frames are attended (the fork's author documents the
    same semantics at text_encoder.py:300-306).
    """

    def __init__(self, dim: int, heads: int, dim_head: int, dropout: float):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(dropout)])

    def forward(self, x, key_bias):
        B, T, _ = x.shape
        split = lambda y: y.view(B, T, self.heads, self.dim_head).transpose(1, 2)
        q, k, v = split(self.to_q(x)), split(self.to_k(x)), split(self.to_v(x))
        scores = torch.matmul(q, k.transpose(-1, -2)) * (self.dim_head ** -0.5)
        if key_bias is not None:
            scores = scores + key_bias.to(scores.dtype).view(B, 1, 1, T)
        o = torch.matmul(torch.softmax(scores, dim=-1), v)
        o = o.transpose(1, 2).reshape(B, T, self.heads * self.dim_head)
        return self.to_out[1](self.to_out[0](o))


class TransformerBlock(nn.Module):
    """ref: transformer.py:123-303 with norm_type="layer_norm", no cross attention:
    x += attn1(norm1(x), mask);  x += ff(norm3(x))."""

    def __init__(self, dim: int, heads: int, dim_head: int, dropout: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = SelfAttention(dim, heads, dim_head, dropout)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = _CompiledNameShim(FeedForwardNet(dim, dropout))

    def forward(self, x, key_bias):
        x = self.attn1(self.norm1(x), key_bias) + x
        return self.ff(self.norm3(x)) + x


# --------------------------------------------------------------------------- the U-Net
class Decoder(nn.Module):
    """ref: decoder.py:202-426.  Same module tree / parameter names as the reference."""

    def __init__(self, in_channels, out_channels, channels=(256, 256), dropout=0.05, attention_head_dim=64,
                 n_blocks=1, num_mid_blocks=2, num_heads=4, down_block_type="transformer",
                 mid_block_type="transformer", up_block_type="transformer"):
        super().__init__()
        for kind in (down_block_type, mid_block_type, up_block_type):
            if kind != "transformer":  # decoder.py:312-339: "conformer" is unused by every config
                raise ValueError(f"Unknown or unsupported block type {kind}")
        channels = tuple(channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        temb_dim = channels[0] * 4
        self.time_mlp = TimeMLP(in_channels, temb_dim)

        def stack(c):
            return nn.ModuleList([TransformerBlock(c, num_heads, attention_head_dim, dropout) for _ in range(n_blocks)])

        self.down_blocks, self.mid_blocks, self.up_blocks = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        cout = in_channels
        for i, c in enumerate(channels):  # decoder.py:234-255
            cin, cout = cout, c
            tail = StridedDown(cout) if i < len(channels) - 1 else nn.Conv1d(cout, cout, 3, padding=1)
            self.down_blocks.append(nn.ModuleList([ResBlock(cin, cout, temb_dim), stack(cout), tail]))
        for _ in range(num_mid_blocks):  # decoder.py:257-276
            self.mid_blocks.append(nn.ModuleList([ResBlock(channels[-1], cout, temb_dim), stack(cout)]))
        ups = channels[::-1] + (channels[0],)
        for i in range(len(ups) - 1):  # decoder.py:278-307
            cin, cout = ups[i], ups[i + 1]
            tail = TransposedUp(cout) if i < len(ups) - 2 else nn.Conv1d(cout, cout, 3, padding=1)
            self.up_blocks.append(nn.ModuleList([ResBlock(2 * cin, cout, temb_dim), stack(cout), tail]))
        self.final_block = ConvNormAct(ups[-1], ups[-1])
        self.final_proj = nn.Conv1d(ups[-1], out_channels, 1)
        self.reset_parameters()

    def reset_parameters(self):
        """ref: decoder.py:341-357.  ConvTranspose1d is not a Conv1d: it keeps torch's default."""
        for m in self.modules():
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.GroupNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    @staticmethod
    def _run_stack(blocks, x, mask):
        x = x.transpose(1, 2)
        bias = mask[:, 0, :]
        for blk in blocks:
            x = blk(x, bias)
        return x.transpose(1, 2)

    def forward(self, x, mask, mu, t, spks=None):
        """ref: decoder.py:359-426.  ``spks`` (B, S) is NOT in the fork (removed, documentation/PROBLEMS.md:41-46): it restates
        upstream Matcha-TTS, where the speaker vector is broadcast over time and concatenated after [x, mu]
        (in_channels = 2 F + S) -- BASELINE config 5 names it."""
        temb = self.time_mlp(sinusoidal_embedding(t, self.in_channels).to(x.dtype))
        x = torch.cat([x, mu], dim=1)
        if spks is not None:
            x = torch.cat([x, spks.unsqueeze(-1).expand(-1, -1, x.shape[-1])], dim=1)
        skips, masks = [], [mask]
        for res, blocks, tail in self.down_blocks:
            m = masks[-1]
            x = self._run_stack(blocks, res(x, m, temb), m)
            skips.append(x)  # saved unmasked (decoder.py:388)
            x = tail(x * m)
            masks.append(m[:, :, ::2])
        masks = masks[:-1]
        m = masks[-1]
        for res, blocks in self.mid_blocks:
            x = self._run_stack(blocks, res(x, m, temb), m)
        for res, blocks, tail in self.up_blocks:
            m = masks.pop()
            x = self._run_stack(blocks, res(torch.cat([x, skips.pop()], dim=1), m, temb), m)
            x = tail(x * m)
        x = self.final_block(x, m)
        return self.final_proj(x * m) * mask


# --------------------------------------------------------------------------- ODE driver
def odeint_fixed_grid(func, y0, t_grid, method: str):
    """Restatement of torchdiffeq's FixedGridODESolver (call site ref: flow_matching.py:62):
    the grid is ``t_grid`` itself, dt = t1 - t0, and the value at a grid point is the state.
    Returns the final state only (the reference takes ``trajectory[-1]``)."""
    if method not in SOLVER_NFE:
        raise ValueError(f"Unknown fixed-grid solver {method!r}; expected one of {sorted(SOLVER_NFE)}")
    y = y0
    for t0, t1 in zip(t_grid[:-1], t_grid[1:]):
        dt = t1 - t0
        k1 = func(t0, y)
        if method == "euler":
            dy = dt * k1
        elif method == "midpoint":
            half = 0.5 * dt
            dy = dt * func(t0 + half, y + k1 * half)
        elif method == "heun3":
            k2 = func(t0 + dt / 3, y + dt * k1 / 3)
            k3 = func(t0 + dt * 2 / 3, y + dt * k2 * 2 / 3)
            dy = dt * (k1 + 3 * k3) / 4
        else:  # rk4 == torchdiffeq's 3/8-rule variant
            k2 = func(t0 + dt / 3, y + dt * k1 / 3)
            k3 = func(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
            k4 = func(t1, y + dt * (k1 - k2 + k3))
            dy = dt * (k1 + 3 * (k2 + k3) + k4) / 8
        y = y + dy
    return y


class CFM(nn.Module):
    """ref: flow_matching.py:11-117 (BASECFM + CFM)."""

    def __init__(self, in_channels, out_channel, cfm_params, decoder_params):
        super().__init__()
        self.n_feats = in_channels
        self.solver = cfm_params.solver
        self.sigma_min = getattr(cfm_params, "sigma_min", 1e-4)
        self.use_mu_prior = getattr(cfm_params, "use_mu_prior", False)
        self.estimator = Decoder(in_channels=in_channels, out_channels=out_channel, **decoder_params)

    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps):
        """ref: flow_matching.py:25-58 (seed-42 noise on mu's device, linspace grid)."""
        g = torch.Generator(device=mu.device)
        g.manual_seed(42)
        noise = torch.randn(mu.shape, generator=g, device=mu.device, dtype=mu.dtype)
        z = mu + noise if self.use_mu_prior else noise
        t_span = torch.linspace(0, 1, n_timesteps + 1, device=mu.device)
        return self.solve(z, t_span=t_span, mu=mu, mask=mask)

    @torch.inference_mode()
    def solve(self, x, t_span, mu, mask, spks=None):
        """ref: flow_matching.py:60-63 + ode_solver_wrapper.py:11-16."""
        return odeint_fixed_grid(lambda t, y: self.estimator(y, mask, mu, t, spks), x, t_span.to(x.dtype), self.solver)

    def compute_loss(self, x1, mask, mu):
        """ref: flow_matching.py:65-107."""
        b = mu.shape[0]
        t = torch.rand([b, 1, 1], device=mu.device, dtype=mu.dtype)
        x0 = mu + torch.randn_like(x1) if self.use_mu_prior else torch.randn_like(x1)
        y = (1 - (1 - self.sigma_min) * t) * x0 + t * x1
        u = x1 - (1 - self.sigma_min) * x0
        pred = self.estimator(y, mask, mu, t.squeeze())
        return F.mse_loss(pred * mask, u * mask, reduction="sum") / (torch.sum(mask) * u.shape[1])


def sequence_mask(lengths: torch.Tensor, max_length: int) -> torch.Tensor:
    """ref: matcha/utils/model.py:7-9."""
    return torch.arange(max_length, dtype=lengths.dtype, device=lengths.device).unsqueeze(0) < lengths.unsqueeze(1)
