"""Executable specification of the pad-aware packed schedule the CUDA library implements.

Test infrastructure (torch, CPU).  One utterance of L valid frames inside a batch padded to T is
evaluated on L+1 token-major rows (row L = conv halo / virtual pad token); the (T-L) padded
frames of the reference enter only through closed forms (DESIGN.md "pad-aware packing"):

  * GroupNorm: stats over rows 0..L-1, plus the halo row L when P>=1, plus (P-1)*(b_c, b_c^2)
    for the pure-bias rows when P>=2, divided by (C/8)*T;
  * attention: one pad token per utterance whose key carries bias log(P) relative to the valid
    keys' +1 (the reference's float mask is additive: reference decoder.py:379-385 ->
    diffusers Attention); no pad key when P == 0.

``tests/test_packed_identity.py`` checks this against the dense padded oracle in fp64.
The functions read weights from an ``oracle.cfm_oracle.Decoder`` state_dict.
"""
import math

import torch
import torch.nn.functional as F


def _mish(x):
    return x * torch.tanh(F.softplus(x))


class PackedEstimator:
    def __init__(self, sd, n_heads, head_dim, n_blocks, n_mid):
        self.sd = {k: v for k, v in sd.items()}
        self.H, self.d, self.nb, self.nm = n_heads, head_dim, n_blocks, n_mid

    # ---- primitive ops on [rows, C] token-major tensors -------------------------------------
    def conv3(self, x, name):
        w, b = self.sd[name + ".weight"], self.sd[name + ".bias"]  # [Cout, Cin, 3]
        z = torch.zeros_like(x[:1])
        xm, xp = torch.cat([z, x[:-1]]), torch.cat([x[1:], z])
        return b + xm @ w[:, :, 0].T + x @ w[:, :, 1].T + xp @ w[:, :, 2].T

    def gn_mish(self, h, name_conv, name_gn, L, P, T):
        """h: raw conv output rows 0..L (L+1 rows).  Returns Mish(GN(h)) for all rows (caller masks)."""
        b = self.sd[name_conv + ".bias"]
        g, beta = self.sd[name_gn + ".weight"], self.sd[name_gn + ".bias"]
        C = h.shape[1]
        gs = C // 8
        rows = L + (1 if P >= 1 else 0)
        s1 = h[:rows].reshape(rows, 8, gs).sum(dim=(0, 2))
        s2 = (h[:rows] ** 2).reshape(rows, 8, gs).sum(dim=(0, 2))
        if P >= 2:
            s1 = s1 + (P - 1) * b.reshape(8, gs).sum(1)
            s2 = s2 + (P - 1) * (b ** 2).reshape(8, gs).sum(1)
        n = gs * T
        mean = s1 / n
        var = s2 / n - mean ** 2
        rstd = torch.rsqrt(var + 1e-5)
        hn = (h.reshape(-1, 8, gs) - mean[None, :, None]) * rstd[None, :, None]
        return _mish(hn.reshape(-1, C) * g + beta)

    def resnet(self, xin, prefix, temb, L, P, T):
        """xin: [L+1, Cin] already masked (rows >= L zero)."""
        valid = (torch.arange(xin.shape[0]) < L).to(xin.dtype)[:, None]
        tp = _mish(temb) @ self.sd[prefix + ".mlp.1.weight"].T + self.sd[prefix + ".mlp.1.bias"]
        h = self.gn_mish(self.conv3(xin, prefix + ".block1.block.0"), prefix + ".block1.block.0",
                         prefix + ".block1.block.1", L, P, T)
        a = (h + tp) * valid
        h = self.gn_mish(self.conv3(a, prefix + ".block2.block.0"), prefix + ".block2.block.0",
                         prefix + ".block2.block.1", L, P, T) * valid
        r = xin @ self.sd[prefix + ".res_conv.weight"][:, :, 0].T + self.sd[prefix + ".res_conv.bias"]
        return h + r

    def transformer(self, x, prefix, L, P):
        """x: [L+1, C]; keys 0..L, the last one being the pad token with bias log P - 1 (vs 0)."""
        sd, H, d = self.sd, self.H, self.d
        R = x.shape[0]
        xn = F.layer_norm(x, (x.shape[1],), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"])
        q = (xn @ sd[prefix + ".attn1.to_q.weight"].T).view(R, H, d).transpose(0, 1)
        k = (xn @ sd[prefix + ".attn1.to_k.weight"].T).view(R, H, d).transpose(0, 1)
        v = (xn @ sd[prefix + ".attn1.to_v.weight"].T).view(R, H, d).transpose(0, 1)
        bias = torch.zeros(R, dtype=x.dtype)
        bias[L] = (math.log(P) - 1.0) if P >= 1 else -math.inf
        s = q @ k.transpose(1, 2) * d ** -0.5 + bias
        o = (torch.softmax(s, dim=-1) @ v).transpose(0, 1).reshape(R, H * d)
        x = x + o @ sd[prefix + ".attn1.to_out.0.weight"].T + sd[prefix + ".attn1.to_out.0.bias"]
        xn = F.layer_norm(x, (x.shape[1],), sd[prefix + ".norm3.weight"], sd[prefix + ".norm3.bias"])
        ff = prefix + ".ff._orig_mod.net"
        h = xn @ sd[ff + ".0.proj.weight"].T + sd[ff + ".0.proj.bias"]
        h = h + torch.sin(h * torch.exp(sd[ff + ".0.alpha"])) ** 2 / (torch.exp(sd[ff + ".0.beta"]) + 1e-9)
        return x + h @ sd[ff + ".2.weight"].T + sd[ff + ".2.bias"]

    def stage(self, xin, prefix, temb, L, P, T):
        x = self.resnet(xin, prefix + ".0", temb, L, P, T)
        for j in range(self.nb):
            x = self.transformer(x, f"{prefix}.1.{j}", L, P)
        return x

    # ---- one NFE for one utterance -------------------------------------------------------------
    def __call__(self, x, mu, t_emb_sin, L, T):
        """x, mu: [L, F] valid frames (token-major); t_emb_sin: [in_channels] sinusoidal features.
        Returns v: [L, F]."""
        sd = self.sd
        dt = x.dtype
        P, L2, T2 = T - L, (L + 1) // 2, T // 2
        P2 = T2 - L2
        temb = F.silu(t_emb_sin @ sd["time_mlp.linear_1.weight"].T + sd["time_mlp.linear_1.bias"])
        temb = temb @ sd["time_mlp.linear_2.weight"].T + sd["time_mlp.linear_2.bias"]
        C = sd["final_proj.weight"].shape[1]
        m1 = (torch.arange(L + 1) < L).to(dt)[:, None]
        m2 = (torch.arange(L2 + 1) < L2).to(dt)[:, None]

        xin = torch.cat([torch.cat([x, mu], 1), torch.zeros(1, x.shape[1] + mu.shape[1], dtype=dt)])
        h0 = self.stage(xin, "down_blocks.0", temb, L, P, T)          # [L+1, C], row L = pad token
        a = h0 * m1
        # strided conv k3 s2 p1: out[j] = b + W0 a[2j-1] + W1 a[2j] + W2 a[2j+1]   (rows >= L are zero)
        ap = torch.cat([torch.zeros(1, C, dtype=dt), a, torch.zeros(3, C, dtype=dt)])
        w, b = sd["down_blocks.0.2.conv.weight"], sd["down_blocks.0.2.conv.bias"]
        j = torch.arange(L2 + 1)
        d = b + ap[2 * j] @ w[:, :, 0].T + ap[2 * j + 1] @ w[:, :, 1].T + ap[2 * j + 2] @ w[:, :, 2].T
        h1 = self.stage(d * m2, "down_blocks.1", temb, L2, P2, T2)
        y = self.conv3(h1 * m2, "down_blocks.1.2") * m2
        for i in range(self.nm):
            y = self.stage(y * m2, f"mid_blocks.{i}", temb, L2, P2, T2)
        y = self.stage(torch.cat([y, h1], 1) * m2, "up_blocks.0", temb, L2, P2, T2)
        # transposed conv k4 s2 p1: out[2i] = b + W1^T u[i] + W3^T u[i-1];  out[2i+1] = b + W2^T u[i] + W0^T u[i+1]
        u = y * m2
        w, b = sd["up_blocks.0.2.conv.weight"], sd["up_blocks.0.2.conv.bias"]  # [Cin, Cout, 4]
        um = torch.cat([torch.zeros(1, C, dtype=dt), u[:-1]])
        up = torch.cat([u[1:], torch.zeros(1, C, dtype=dt)])
        even = b + u @ w[:, :, 1] + um @ w[:, :, 3]
        odd = b + u @ w[:, :, 2] + up @ w[:, :, 0]
        full = torch.stack([even, odd], 1).reshape(-1, C)[: L + 1]
        y = self.stage(torch.cat([full, h0], 1) * m1, "up_blocks.1", temb, L, P, T)
        y = self.conv3(y * m1, "up_blocks.1.2") * m1
        y = self.gn_mish(self.conv3(y, "final_block.block.0"), "final_block.block.0", "final_block.block.1",
                         L, P, T) * m1
        v = y @ sd["final_proj.weight"][:, :, 0].T + sd["final_proj.bias"]
        return v[:L]
