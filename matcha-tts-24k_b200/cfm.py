"""Drop-in replacement for ``matcha.models.components.flow_matching.CFM`` (reference flow_matching.py:110-117).

Same constructor, same ``forward(mu, mask, n_timesteps)`` call (reference matcha/inference.py:169), same mutable
``.solver`` / ``.estimator`` attributes (reference cli.py:94, server.py:43,47,109) and the same state-dict keys below
``estimator.``; the arithmetic runs in libcfm_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/cfm_b200.h).  There is no PyTorch or CPU fallback: CPU tensors, a missing library or a non-B200 device raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _native as N
from .estimator import EstimatorWeights, config_from_decoder_params


def lengths_from_mask(mask: torch.Tensor):
    """mask: (B, 1, T) or (B, T), 1.0 = valid.  Returns python ints; rejects anything that is not a prefix mask
    (the reference only ever builds ``sequence_mask(lengths, T)``: matcha/utils/model.py:7-9, inference.py:167)."""
    if mask.dtype == torch.bool:
        raise ValueError("boolean masks select diffusers' masking semantics in the reference; this path implements "
                         "the float (additive) mask every reference caller passes")
    m = mask[:, 0, :] if mask.ndim == 3 else mask
    T = m.shape[-1]
    nz = m != 0  # counted as integers: a bf16 / fp16 mask would round its own sum above 256 / 2048 frames
    lengths = nz.sum(-1, dtype=torch.int64)
    is_prefix = (nz == (torch.arange(T, device=m.device).unsqueeze(0) < lengths.unsqueeze(1))).all()
    is_01 = ((m == 0) | (m == 1)).all()
    packed = torch.cat([lengths, (is_prefix & is_01).to(torch.int64).reshape(1)]).tolist()  # one device-to-host copy
    if not packed[-1]:
        raise ValueError("mask is not a 0/1 prefix (sequence) mask")
    return packed[:-1]


class CFM(torch.nn.Module):
    def __init__(self, in_channels, out_channel, cfm_params, decoder_params, precision: Optional[str] = None,
                 flags: int = 0, lanes: Optional[int] = None, lane_min_rows: int = 0):
        super().__init__()
        self.n_feats = in_channels  # sic: the reference passes in_channels as n_feats (flow_matching.py:112-115)
        self.solver = cfm_params.solver
        self.sigma_min = getattr(cfm_params, "sigma_min", 1e-4)
        self.use_mu_prior = getattr(cfm_params, "use_mu_prior", False)
        self.precision = precision or os.environ.get("CFM_B200_PRECISION", "bf16")
        if self.precision not in N.PREC:
            raise ValueError(f"precision must be one of {sorted(N.PREC)}, got {self.precision!r}")
        self.flags = int(flags) | int(os.environ.get("CFM_B200_FLAGS", "0"))
        # graph branches per decode (cfm_set_lanes); None keeps the library default / CFM_B200_LANES
        self.lanes, self.lane_min_rows = lanes, int(lane_min_rows)
        cfg = config_from_decoder_params(in_channels, out_channel, **dict(decoder_params))
        est = EstimatorWeights(cfg)
        est._owner = [self]
        self.estimator = est
        # survives ``model.decoder.estimator = torch.compile(...)`` (reference server.py:47): same Parameter objects
        object.__setattr__(self, "_weights", est)
        object.__setattr__(self, "_handle", None)
        object.__setattr__(self, "_weights_sig", None)
        object.__setattr__(self, "_plan_key", None)
        object.__setattr__(self, "_lib", None)

    # ------------------------------------------------------------------ native plumbing
    def _native(self, device: torch.device):
        if self._handle is not None and self._device == device:
            return self._lib, self._handle
        self.close()
        lib = N.load_library()
        c = self._weights.cfg
        cfg = N.Config(c.in_channels, c.out_channels, c.channels, c.n_heads, c.head_dim, c.n_blocks, c.n_mid_blocks,
                       N.PREC[self.precision], device.index if device.index is not None else torch.cuda.current_device(),
                       self.flags)
        handle = C.c_void_p()
        N.check(lib, None, lib.cfm_create(C.byref(cfg), C.byref(handle)))
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_handle", handle)
        object.__setattr__(self, "_device", device)
        object.__setattr__(self, "_weights_sig", None)
        object.__setattr__(self, "_plan_key", None)
        if self.lanes is not None:
            N.check(lib, handle, lib.cfm_set_lanes(handle, int(self.lanes), self.lane_min_rows))
        return lib, handle

    def set_option(self, key: str, value: int):
        """cfm_set_option: kernel-selection switches for A/B measurements; drops the current plan."""
        if self._handle is None:
            raise RuntimeError("set_option needs a live handle: call refresh(device) or run a decode first")
        N.check(self._lib, self._handle, self._lib.cfm_set_option(self._handle, key.encode(), int(value)))
        object.__setattr__(self, "_plan_key", None)

    def set_lanes(self, lanes: int, min_rows: int = 0):
        """Number of independent utterance groups the next plan runs as parallel graph branches (results unchanged)."""
        self.lanes, self.lane_min_rows = int(lanes), int(min_rows)
        if self._handle is not None:
            N.check(self._lib, self._handle, self._lib.cfm_set_lanes(self._handle, self.lanes, self.lane_min_rows))
            object.__setattr__(self, "_plan_key", None)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._lib is not None:
            self._lib.cfm_destroy(self._handle)
        object.__setattr__(self, "_handle", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh(self, device: Optional[torch.device] = None):
        """(Re)packs the estimator parameters into the library.  Called automatically when a parameter's storage or
        version counter changed (``load_state_dict`` and ``.to(device)`` happen after construction: inference.py:193-194)."""
        params = list(self._weights.state_dict(keep_vars=True).items())
        device = device or params[0][1].device
        lib, handle = self._native(device)
        keep, descs = [], (N.WeightDesc * len(params))()
        for i, (name, p) in enumerate(params):
            t = p.detach()
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            descs[i].name = name.encode()
            descs[i].data = t.data_ptr()
            descs[i].ndim = t.ndim
            for j, s in enumerate(t.shape):
                descs[i].shape[j] = s
        torch.cuda.synchronize(device)
        N.check(lib, handle, lib.cfm_load_weights(handle, descs, len(params)))
        object.__setattr__(self, "_weights_sig", self._signature())
        object.__setattr__(self, "_plan_key", None)

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self._weights.parameters())

    def _ensure(self, device, lengths: Sequence[int], T: int, t_span: Sequence[float], solver: str):
        if self.training:
            raise RuntimeError("the CUDA decode path is inference-only (dropout sites are identity only in eval mode); "
                               "call .eval() as the reference does at load time (matcha/inference.py:195)")
        if solver not in N.SOLVERS:
            raise ValueError(f"unknown ODE solver {solver!r}; supported fixed-grid solvers: {sorted(N.SOLVERS)}")
        if self._handle is None or self._device != device or self._weights_sig != self._signature():
            self.refresh(device)
        lib, handle = self._lib, self._handle
        key = (tuple(lengths), T, tuple(t_span), solver)
        if key != self._plan_key:
            arr = (C.c_int32 * len(lengths))(*lengths)
            ts = (C.c_float * len(t_span))(*t_span)
            N.check(lib, handle, lib.cfm_plan(handle, arr, len(lengths), T, ts, len(t_span), N.SOLVERS[solver]))
            object.__setattr__(self, "_plan_key", key)
        return lib, handle

    @staticmethod
    def _prep(x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("CFM (B200) needs CUDA tensors: there is no CPU path")
        return x.detach().to(torch.float32).contiguous()

    # ------------------------------------------------------------------ reference surface
    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps, temperature: float = 1.0, spks=None, cond=None, lengths=None):
        """reference flow_matching.py:25-58.  ``temperature``/``spks``/``cond`` are the upstream Matcha-TTS superset;
        with their defaults this is exactly the fork's ``forward(mu, mask, n_timesteps)``."""
        g = torch.Generator(device=mu.device)
        g.manual_seed(42)
        noise = torch.randn(mu.shape, generator=g, device=mu.device, dtype=mu.dtype)
        if temperature != 1.0:
            noise = noise * temperature
        z = mu + noise if self.use_mu_prior else noise
        return self.solve(z, t_span=self._t_span(int(n_timesteps), mu.device), mu=mu, mask=mask, lengths=lengths, spks=spks)

    def _t_span(self, n_timesteps: int, device):
        """linspace(0, 1, n + 1) evaluated on the caller's device as the reference does (flow_matching.py:57) and cached with
        its host copy, so that a repeated forward() does not pay a device-to-host sync for the time grid."""
        cache = self.__dict__.setdefault("_tspan_cache", {})
        key = (n_timesteps, str(device))
        if key not in cache:
            t = torch.linspace(0, 1, n_timesteps + 1, device=device)
            cache[key] = (t, [float(v) for v in t.to(torch.float32).cpu().tolist()])
        return cache[key][0]

    def _t_list(self, t_span):
        for t, lst in self.__dict__.get("_tspan_cache", {}).values():
            if t is t_span:
                return lst
        return [float(v) for v in t_span.detach().to(torch.float32).cpu().tolist()]

    @torch.inference_mode()
    def solve(self, x, t_span, mu, mask, lengths=None, spks=None, out=None):
        """reference flow_matching.py:60-63; ``x`` is the injected initial state z.  ``spks`` (B, S): upstream-style
        speaker conditioning, only for an estimator built with in_channels = 2*n_feats + S.  ``out``: optional result tensor
        (same shape, fp32, contiguous, same device): a caller that queues decodes back to back avoids a fresh allocation per
        call (with torch's caching allocator that can be a blocking cudaMalloc, see bench.py)."""
        mu_, x_ = self._prep(mu), self._prep(x)
        B, F, T = self._check_shapes(mu_, x_)
        spks_ = self._check_spks(spks, B, prep=True)
        if mask is not None and (mask.shape[0] != B or mask.shape[-1] != T):
            raise ValueError("x, mu and mask disagree on (B, F, T)")
        lengths = self._check_lengths(lengths if lengths is not None else lengths_from_mask(mask), B, T)
        ts = self._t_list(t_span)
        lib, handle = self._ensure(mu_.device, lengths, T, ts, self.solver)
        if out is None:
            out = torch.empty_like(mu_)
        elif (out.device != mu_.device or out.dtype != torch.float32 or tuple(out.shape) != tuple(mu_.shape) or not out.is_contiguous()):
            raise ValueError("out must be a contiguous fp32 tensor of mu's shape on mu's device")
        stream = torch.cuda.current_stream(mu_.device).cuda_stream
        N.check(lib, handle, lib.cfm_set_speakers(handle, spks_.data_ptr() if spks_ is not None else None))
        N.check(lib, handle, lib.cfm_solve(handle, mu_.data_ptr(), x_.data_ptr(), out.data_ptr(), stream))
        return out

    # ------------------------------------------------------------------ argument checks (the library trusts its pointers)
    def _check_shapes(self, mu_, x_):
        """The library packs with its own F / B / T: a tensor of another shape would be read out of bounds, where the
        reference raises a conv shape error."""
        if mu_.ndim != 3:
            raise ValueError(f"mu must have shape (B, n_feats, T), got {tuple(mu_.shape)}")
        B, F, T = mu_.shape
        if F != self._weights.cfg.out_channels:
            raise ValueError(f"mu has {F} mel channels, this estimator was built for {self._weights.cfg.out_channels}")
        if tuple(x_.shape) != tuple(mu_.shape):
            raise ValueError(f"x {tuple(x_.shape)} and mu {tuple(mu_.shape)} disagree on (B, F, T)")
        return B, F, T

    def _check_spks(self, spks, B, prep):
        spk_dim = self._weights.cfg.in_channels - 2 * self._weights.cfg.out_channels
        if (spks is None) != (spk_dim == 0):
            raise ValueError(f"estimator has {spk_dim} speaker channels but spks is {'missing' if spks is None else 'given'}")
        if spks is None:
            return None
        spks_ = self._prep(spks) if prep else spks.detach().float().contiguous()
        if tuple(spks_.shape) != (B, spk_dim):
            raise ValueError(f"spks must have shape ({B}, {spk_dim}), got {tuple(spks_.shape)}")
        return spks_

    @staticmethod
    def _check_lengths(lengths, B, T):
        lengths = [int(v) for v in lengths]
        if len(lengths) != B:
            raise ValueError(f"lengths has {len(lengths)} entries for a batch of {B}")
        if any(v < 1 or v > T for v in lengths):
            raise ValueError(f"lengths must lie in [1, T={T}]")
        return lengths

    @torch.inference_mode()
    def solve_host(self, x, t_span, mu, lengths, device=None, spks=None, out=None):
        """Host-buffer entry (cfm_solve_host): CPU tensors in, CPU tensor out, copies inside the library call.  Returns a new
        tensor unless ``out=`` (same shape, fp32, contiguous; pinned memory makes the device-to-host copy asynchronous) is given."""
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if x.is_cuda or mu.is_cuda:
            raise ValueError("solve_host takes host tensors; use solve() for device tensors")
        mu_, x_ = mu.detach().float().contiguous(), x.detach().float().contiguous()
        B, F, T = self._check_shapes(mu_, x_)
        spks_ = self._check_spks(spks, B, prep=False)
        lengths = self._check_lengths(lengths, B, T)
        ts = [float(v) for v in torch.as_tensor(t_span, dtype=torch.float32).tolist()]
        lib, handle = self._ensure(device, lengths, T, ts, self.solver)
        if out is None:
            out = torch.empty_like(mu_)
        elif out.is_cuda or out.dtype != torch.float32 or tuple(out.shape) != tuple(mu_.shape) or not out.is_contiguous():
            raise ValueError("out must be a contiguous fp32 host tensor of mu's shape")
        if spks_ is None:
            N.check(lib, handle, lib.cfm_solve_host(handle, mu_.data_ptr(), x_.data_ptr(), out.data_ptr()))
        else:
            N.check(lib, handle, lib.cfm_solve_host_spks(handle, mu_.data_ptr(), x_.data_ptr(), spks_.data_ptr(), out.data_ptr()))
        return out

    def _estimator_call(self, x, mask, mu, t, spks=None, lengths=None):
        """One call of Decoder.forward(x, mask, mu, t) (reference decoder.py:359-426).  ``t``: 0-dim (the ODE solver's call,
        ode_solver_wrapper.py:11-16) or shape (B,) (the training forward, flow_matching.py:84-97)."""
        mu_, x_ = self._prep(mu), self._prep(x)
        B, F, T = self._check_shapes(mu_, x_)
        spks_ = self._check_spks(spks, B, prep=True)
        tt = torch.as_tensor(t).detach().to(torch.float32).reshape(-1).cpu()
        if tt.numel() not in (1, B):
            raise ValueError(f"t must be a scalar or have one entry per utterance ({B}), got {tt.numel()}")
        lengths = self._check_lengths(lengths if lengths is not None else lengths_from_mask(mask), B, T)
        lib, handle = self._ensure(mu_.device, lengths, T, [0.0, 1.0], "euler")
        v = torch.empty_like(mu_)
        stream = torch.cuda.current_stream(mu_.device).cuda_stream
        arr = (C.c_float * tt.numel())(*tt.tolist())
        N.check(lib, handle, lib.cfm_set_speakers(handle, spks_.data_ptr() if spks_ is not None else None))
        N.check(lib, handle, lib.cfm_estimator_t(handle, x_.data_ptr(), mu_.data_ptr(), arr, tt.numel(), v.data_ptr(), stream))
        return v

    @torch.inference_mode()
    def compute_loss(self, x1, mask, mu, t=None, x0=None):
        """Forward value of the conditional-flow-matching loss (reference flow_matching.py:65-107): same draws (``t`` then the
        prior noise, from the global generator of ``mu``'s device), same masked MSE; the estimator call with t of shape (B,)
        runs on the CUDA path.  No autograd graph is built - this library has no backward, so the value
        serves validation / monitoring, not optimisation.  ``t`` (B,) and ``x0`` can be injected for parity tests."""
        b = mu.shape[0]
        if t is None:
            t = torch.rand([b, 1, 1], device=mu.device, dtype=mu.dtype)
        t = t.reshape(b, 1, 1).to(device=mu.device, dtype=mu.dtype)
        if x0 is None:
            x0 = mu + torch.randn_like(x1) if self.use_mu_prior else torch.randn_like(x1)
        y = (1 - (1 - self.sigma_min) * t) * x0 + t * x1
        u = x1 - (1 - self.sigma_min) * x0
        was_training = self.training
        self.eval()  # dropout (p = 0.05 at two sites) is NOT applied: the value is the eval-mode loss
        try:
            pred = self._estimator_call(y, mask, mu, t.reshape(b))
        finally:
            self.train(was_training)
        return torch.nn.functional.mse_loss(pred * mask, u * mask, reduction="sum") / (torch.sum(mask) * u.shape[1])

    # ------------------------------------------------------------------ front / back of the decode (reference inference.py:146-172)
    @torch.inference_mode()
    def expand_encoder_output(self, mu_x, phoneme_durations):
        """Front of the decode (reference matcha/inference.py:146-167): ``mu_x`` (B, n_feats, Tx) and the integer-valued
        ``phoneme_durations`` (B, Tx) (rounded, clamped and masked as at inference.py:143) -> ``(mu_y, y_mask, y_lengths)`` with
        mu_y (B, n_feats, T), y_mask (B, 1, T) float, y_lengths a python list.  The alignment path is never materialised."""
        mu_x = self._prep(mu_x)
        dur = self._prep(phoneme_durations)
        B, F, Tx = mu_x.shape
        if F != self._weights.cfg.out_channels or tuple(dur.shape) != (B, Tx):
            raise ValueError("mu_x must be (B, n_feats, Tx) and phoneme_durations (B, Tx)")
        if self._handle is None or self._device != mu_x.device:
            self.refresh(mu_x.device)
        lib, handle = self._lib, self._handle
        stream = torch.cuda.current_stream(mu_x.device).cuda_stream
        cum = torch.empty(B, Tx, dtype=torch.int32, device=mu_x.device)
        fine = torch.empty(B, dtype=torch.int32, device=mu_x.device)
        N.check(lib, handle, lib.cfm_front_durations(handle, dur.data_ptr(), B, Tx, cum.data_ptr(), fine.data_ptr(), stream))
        fine_h = fine.tolist()  # the one host round trip of the front: the output shape depends on max() (reference :148)
        # reference: y_fine_max_length_ = fix_len_compatibility(max) * 2 = 4 ceil(max / 2); mel length T = that // 2 (always even)
        t_pad = 2 * ((max(fine_h) + 1) // 2)
        y_lengths = [max((v + 1) // 2, 1) for v in fine_h]
        mu_y = torch.empty(B, F, t_pad, dtype=torch.float32, device=mu_x.device)
        y_mask = torch.empty(B, 1, t_pad, dtype=torch.float32, device=mu_x.device)
        N.check(lib, handle, lib.cfm_front_expand(handle, mu_x.data_ptr(), cum.data_ptr(), fine.data_ptr(), B, Tx, t_pad, mu_y.data_ptr(),
                                                  y_mask.data_ptr(), stream))
        return mu_y, y_mask, y_lengths

    @torch.inference_mode()
    def denormalize(self, decoder_outputs, mel_mean, mel_std, t_out=None):
        """Back of the decode (reference inference.py:170-172): ``decoder_outputs[:, :, :t_out] * mel_std + mel_mean``."""
        x = self._prep(decoder_outputs)
        B, F, T = x.shape
        t_out = T if t_out is None else int(t_out)
        if self._handle is None or self._device != x.device:
            self.refresh(x.device)
        out = torch.empty(B, F, t_out, dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        N.check(self._lib, self._handle, self._lib.cfm_denormalize(self._handle, x.data_ptr(), B, T, t_out, float(mel_mean), float(mel_std),
                                                                   out.data_ptr(), stream))
        return out

    @torch.inference_mode()
    def synthesise_mel(self, mu_x, phoneme_durations, n_timesteps, mel_mean=0.0, mel_std=1.0, temperature: float = 1.0):
        """inference.py:146-172 in one call: front -> CFM.forward (lengths handed over directly: no mask -> lengths round trip) ->
        slice + denormalize.  Returns (mel (B, n_feats, max y_length), y_lengths)."""
        mu_y, y_mask, y_lengths = self.expand_encoder_output(mu_x, phoneme_durations)
        dec = self.forward(mu_y, y_mask, n_timesteps, temperature=temperature, lengths=y_lengths)
        return self.denormalize(dec, mel_mean, mel_std, max(y_lengths)), y_lengths

    def timeline(self, x, t_span, mu, lengths):
        """cfm_debug_timeline: in-situ time of every launch of one direct-launch decode -> list of (tag, M, N, K, flops, us)."""
        mu_, x_ = self._prep(mu), self._prep(x)
        B, F, T = self._check_shapes(mu_, x_)
        lengths = self._check_lengths(lengths, B, T)
        lib, handle = self._ensure(mu_.device, lengths, T, self._t_list(t_span), self.solver)
        out = torch.empty_like(mu_)
        buf = C.create_string_buffer(1 << 20)
        stream = torch.cuda.current_stream(mu_.device).cuda_stream
        N.check(lib, handle, lib.cfm_debug_timeline(handle, mu_.data_ptr(), x_.data_ptr(), out.data_ptr(), buf, len(buf), stream))
        rows = []
        for line in buf.value.decode().splitlines():
            f = line.split(",")
            rows.append((f[0], int(f[1]), int(f[2]), int(f[3]), float(f[4]), float(f[5])))
        return rows

    # ------------------------------------------------------------------ introspection (tests / bench)
    def plan_info(self):
        vals = [C.c_int64() for _ in range(5)]
        N.check(self._lib, self._handle, self._lib.cfm_plan_info(self._handle, *[C.byref(v) for v in vals]))
        keys = ("rows_full", "rows_half", "n_nfe", "kernels_per_solve", "workspace_bytes")
        return dict(zip(keys, (v.value for v in vals)))

    def debug_read(self, name: str) -> torch.Tensor:
        r, c = C.c_int64(), C.c_int64()
        N.check(self._lib, self._handle, self._lib.cfm_debug_read(self._handle, name.encode(), None, 0, C.byref(r), C.byref(c)))
        out = torch.empty(r.value, c.value, dtype=torch.float32)
        N.check(self._lib, self._handle,
                self._lib.cfm_debug_read(self._handle, name.encode(), out.data_ptr(), out.numel(), C.byref(r), C.byref(c)))
        return out


def install():
    """Makes ``from matcha.models.components.flow_matching import CFM`` resolve to this class, so that
    matcha.inference / cli.py / server.py run unchanged (reference import sites: inference.py:6, matcha_tts.py:7)."""
    import importlib

    mod = importlib.import_module("matcha.models.components.flow_matching")
    mod.CFM = CFM
    for name in ("matcha.inference", "matcha.models.matcha_tts"):
        import sys
        if name in sys.modules and hasattr(sys.modules[name], "CFM"):
            sys.modules[name].CFM = CFM
    return CFM
