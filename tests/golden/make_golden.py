"""Generate tests/golden/*.npz by running the REFERENCE's own CFM code.

Run here (build container) only:  python tests/golden/make_golden.py
It needs /root/reference, which does not exist on the GPU box; the outputs are committed.

The reference's decoder.py / transformer.py / flow_matching.py are imported unmodified from
/root/reference.  Their third-party imports are not installed here (no network), so this
script registers minimal stand-ins for exactly the symbols those files import:

  diffusers.models.attention_processor.Attention   <- restated (AttnProcessor2_0 semantics)
  diffusers.models.lora.LoRACompatibleLinear        <- nn.Linear (no LoRA layer attached)
  diffusers.models.attention.AdaLayerNorm[Zero]     <- never instantiated (norm_type="layer_norm")
  diffusers.models.activations.get_activation       <- "silu" -> nn.SiLU
  diffusers.utils.torch_utils.maybe_allow_in_graph  <- identity decorator
  conformer.ConformerBlock                          <- never instantiated
  torchdiffeq.odeint                                <- restated fixed-grid solvers

Everything else -- Decoder.forward, ResnetBlock1D, Block1D, SnakeBeta, FeedForward,
BasicTransformerBlock, BASECFM.solve / forward, OdeSolverWrapper -- is the reference's code.
``torch.compiler.set_stance("force_eager")`` makes the reference's ``torch.compile(self.ff)``
wrapper run eagerly (no Inductor build) while keeping its ``_orig_mod`` key infix.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def _install_third_party_shims():
    class Attention(nn.Module):
        def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False, cross_attention_dim=None,
                     upcast_attention=False):
            super().__init__()
            assert cross_attention_dim is None and not upcast_attention
            inner = heads * dim_head
            self.heads, self.scale = heads, dim_head ** -0.5
            self.to_q = nn.Linear(query_dim, inner, bias=bias)
            self.to_k = nn.Linear(query_dim, inner, bias=bias)
            self.to_v = nn.Linear(query_dim, inner, bias=bias)
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

        def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
            assert encoder_hidden_states is None
            b, t, _ = hidden_states.shape
            q, k, v = self.to_q(hidden_states), self.to_k(hidden_states), self.to_v(hidden_states)
            d = q.shape[-1] // self.heads
            q, k, v = (y.view(b, t, self.heads, d).transpose(1, 2) for y in (q, k, v))
            if attention_mask is not None:  # prepare_attention_mask + view(b, heads, -1, t)
                attention_mask = attention_mask.repeat_interleave(self.heads, dim=0).view(b, self.heads, -1, t)
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
            o = o.transpose(1, 2).reshape(b, t, self.heads * d).to(q.dtype)
            return self.to_out[1](self.to_out[0](o))

    def odeint(func, y0, t, method=None, **kw):
        ys = [y0]
        y = y0
        for t0, t1 in zip(t[:-1], t[1:]):
            dt = t1 - t0
            f0 = func(t0, y)
            if method == "euler":
                dy = dt * f0
            elif method == "midpoint":
                half_dt = 0.5 * dt
                dy = dt * func(t0 + half_dt, y + f0 * half_dt)
            elif method == "heun3":  # torchdiffeq fixed_grid.Heun3: tableau c = (0, 1/3, 2/3), b = (1/4, 0, 3/4)
                k2 = func(t0 + dt / 3, y + dt * f0 / 3)
                k3 = func(t0 + dt * 2 / 3, y + dt * k2 * 2 / 3)
                dy = dt * (f0 + 3 * k3) * 0.25
            elif method == "rk4":
                k2 = func(t0 + dt / 3, y + dt * f0 / 3)
                k3 = func(t0 + dt * 2 / 3, y + dt * (k2 - f0 / 3))
                k4 = func(t1, y + dt * (f0 - k2 + k3))
                dy = (f0 + 3 * (k2 + k3) + k4) * dt * 0.125
            else:
                raise ValueError(method)
            y = y + dy
            ys.append(y)
        return torch.stack(ys)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Never(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("not on the hot path")

    mod("diffusers"); mod("diffusers.models"); mod("diffusers.utils")
    mod("diffusers.models.attention", AdaLayerNorm=_Never, AdaLayerNormZero=_Never)
    mod("diffusers.models.attention_processor", Attention=Attention)
    mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear)
    mod("diffusers.models.activations", get_activation=lambda name: {"silu": nn.SiLU(), "mish": nn.Mish()}[name])
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)
    mod("conformer", ConformerBlock=_Never)
    mod("torchdiffeq", odeint=odeint)


CASES = {
    # name: (decoder_params, lengths, T, solver, n_steps)
    "tiny_euler": (dict(channels=(64, 64), dropout=0.05, attention_head_dim=32, n_blocks=1, num_mid_blocks=1,
                        num_heads=2), [40, 33, 17], 40, "euler", 3),
    "tiny_midpoint": (dict(channels=(64, 64), dropout=0.05, attention_head_dim=32, n_blocks=1, num_mid_blocks=1,
                           num_heads=2), [21, 38], 38, "midpoint", 2),
    "tiny_rk4_padded": (dict(channels=(64, 64), dropout=0.05, attention_head_dim=32, n_blocks=2, num_mid_blocks=1,
                             num_heads=2), [9, 14], 20, "rk4", 1),
    "tiny_heun3": (dict(channels=(64, 64), dropout=0.05, attention_head_dim=32, n_blocks=1, num_mid_blocks=1,
                        num_heads=2), [19, 26], 26, "heun3", 2),
    "prod_euler": (dict(channels=(384, 384), dropout=0.05, attention_head_dim=64, n_blocks=2, num_mid_blocks=2,
                        num_heads=6), [30, 21], 30, "euler", 2),
    "default_euler": (dict(channels=(320, 320), dropout=0.05, attention_head_dim=64, n_blocks=2, num_mid_blocks=2,
                           num_heads=5), [26], 26, "euler", 2),
}


def main():
    torch.compiler.set_stance("force_eager")
    _install_third_party_shims()
    sys.path.insert(0, "/root/reference")
    from matcha.models.components.flow_matching import CFM as RefCFM  # the reference's own class

    import matcha_tts_24k_b200.synthetic as syn

    only = set(sys.argv[1:])
    for name, (dec, lengths, T, solver, n_steps) in CASES.items():
        if only and name not in only:
            continue
        cfm_params = types.SimpleNamespace(solver=solver, sigma_min=1e-4, use_mu_prior=True)
        ref = RefCFM(in_channels=200, out_channel=100, cfm_params=cfm_params, decoder_params=dec).eval()
        syn.fill_named_seed(ref.estimator, seed=1234)
        mu, mask, z, _ = syn.make_inputs(lengths, n_feats=100, seed=7, T=T)
        with torch.inference_mode():
            t_span = torch.linspace(0, 1, n_steps + 1)
            out = ref.solve(z, t_span=t_span, mu=mu, mask=mask)
            v = ref.estimator(z, mask, mu, torch.tensor(0.3))
            fwd = ref(mu, mask, n_steps)  # seed-42 noise path of BASECFM.forward
        keys = sorted(ref.estimator.state_dict().keys())
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            lengths=np.asarray(lengths), T=T, n_steps=n_steps, solver=solver, input_seed=7, weight_seed=1234,
            decoder_params=repr(dec), solve_out=out.numpy(), estimator_v_t03=v.numpy(), forward_out=fwd.numpy(),
            state_dict_keys=np.asarray(keys), n_params=sum(p.numel() for p in ref.estimator.parameters()),
        )
        print(name, tuple(out.shape), float(out.abs().mean()), len(keys))


def front_back_golden():
    """Front / back of the decode with the reference's OWN functions (matcha/utils/model.py needs only torch): the statements of
    matcha/inference.py:146-172 executed on seeded inputs."""
    sys.path.insert(0, "/root/reference")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_utils_model", "/root/reference/matcha/utils/model.py")
    M = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(M)
    g = torch.Generator().manual_seed(11)
    B, F_, Tx = 3, 100, 23
    x_lengths = torch.tensor([23, 17, 9])
    x_mask = M.sequence_mask(x_lengths, Tx).unsqueeze(1).float()
    mu_x = torch.randn(B, F_, Tx, generator=g) * x_mask
    phoneme_durations = (torch.rand(B, Tx, generator=g) * 6.0 - 0.3)
    phoneme_durations = phoneme_durations.round().clamp(min=1) * x_mask.squeeze(1)  # inference.py:143
    # inference.py:146-167, verbatim
    y_fine_lengths = torch.clamp_min(phoneme_durations.sum(dim=1).long(), 1)
    y_fine_max_length = y_fine_lengths.max()
    y_fine_max_length_ = M.fix_len_compatibility(y_fine_max_length) * 2
    y_fine_mask = M.sequence_mask(y_fine_lengths, y_fine_max_length_).unsqueeze(1).to(x_mask.dtype)
    attn_mask_fine = x_mask.unsqueeze(-1) * y_fine_mask.unsqueeze(2)
    attn_fine = M.generate_path(phoneme_durations, attn_mask_fine.squeeze(1)).unsqueeze(1)
    mu_y_fine = torch.matmul(mu_x.float(), attn_fine.float().squeeze(1))
    mu_y = M.downsample(mu_y_fine)
    y_max_length_ = y_fine_max_length_ // 2
    y_lengths = torch.clamp_min((y_fine_lengths + 1) // 2, 1)
    y_max_length = y_lengths.max()
    y_mask = M.sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)
    decoder_outputs = torch.randn(B, F_, y_max_length_, generator=g)
    mel = M.denormalize(decoder_outputs[:, :, :y_max_length], torch.tensor(-5.52), torch.tensor(2.07))  # inference.py:170-172
    np.savez_compressed(os.path.join(HERE, "front_back.npz"), mu_x=mu_x.numpy(), phoneme_durations=phoneme_durations.numpy(),
                        x_mask=x_mask.numpy(), mu_y=mu_y.numpy(), y_mask=y_mask.numpy(), y_lengths=y_lengths.numpy(),
                        y_max_length=int(y_max_length), decoder_outputs=decoder_outputs.numpy(), mel=mel.numpy(), mel_mean=-5.52, mel_std=2.07)
    print("front_back", tuple(mu_y.shape), y_lengths.tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "front_back":
        front_back_golden()
    else:
        main()
        front_back_golden()
