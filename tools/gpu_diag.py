"""GPU bring-up diagnostics: runs each check in its own subprocess (timeout-guarded) so one fault or hang does not
hide the rest.  Usage on the GPU box:  python tools/gpu_diag.py [check ...]   -> gpurun_out/diag.log"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _models(dec, solver, precision, flags, seed=1234):
    import types
    import torch
    import matcha_tts_24k_b200 as P
    from oracle import cfm_oracle as O
    cp = types.SimpleNamespace(solver=solver, sigma_min=1e-4, use_mu_prior=True)
    ora = O.CFM(200, 100, cp, dec).eval()
    P.synthetic.fill_named_seed(ora.estimator, seed)
    m = P.CFM(200, 100, cp, dec, precision=precision, flags=flags).eval()
    m.estimator.load_state_dict(ora.estimator.state_dict())
    return ora, m.cuda()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


TINY = dict(channels=(64, 64), dropout=0.05, attention_head_dim=32, n_blocks=2, num_mid_blocks=2, num_heads=2)
TINY64 = dict(channels=(128, 128), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=1, num_heads=2)


def check_estimator(precision, flags, dec=TINY, lengths=(40, 33, 17), T=40, label=""):
    import torch
    import matcha_tts_24k_b200 as P
    ora, m = _models(dec, "euler", precision, flags)
    mu, mask, z, _ = P.synthetic.make_inputs(list(lengths), seed=7, T=T)
    with torch.inference_mode():
        v_ref = ora.estimator(z, mask, mu, torch.tensor(0.3))
    v = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), torch.tensor(0.3))
    torch.cuda.synchronize()
    print(f"[estimator {label} prec={precision} flags={flags}] rel_l2={rel(v, v_ref):.3e} max_abs={float((v.cpu()-v_ref).abs().max()):.3e}"
          f" finite={bool(torch.isfinite(v).all())}")
    return m, ora


def check_solve(precision, flags, dec=TINY, lengths=(40, 33, 17), T=40, solver="euler", n=3, label=""):
    import torch
    import matcha_tts_24k_b200 as P
    ora, m = _models(dec, solver, precision, flags)
    mu, mask, z, _ = P.synthetic.make_inputs(list(lengths), seed=7, T=T)
    ts = torch.linspace(0, 1, n + 1)
    ref = ora.solve(z, ts, mu, mask)
    out = m.solve(z.cuda(), ts.cuda(), mu.cuda(), mask.cuda())
    torch.cuda.synchronize()
    print(f"[solve {label} {solver}/{n} prec={precision} flags={flags}] rel_l2={rel(out, ref):.3e} "
          f"max_abs={float((out.cpu()-ref).abs().max()):.3e} info={m.plan_info()}")


def c_fp32_simt():
    check_estimator("fp32", 1, label="tiny")
    check_estimator("fp32", 1, lengths=(9, 30, 1, 2), T=64, label="tiny-padded")
    check_solve("fp32", 1, label="tiny")
    check_solve("fp32", 0, label="tiny-graph")
    check_solve("fp32", 0, solver="midpoint", n=2, lengths=(21, 38), T=38, label="tiny")
    check_solve("fp32", 0, solver="rk4", n=1, lengths=(9, 14), T=20, label="tiny")
    check_solve("fp32", 0, solver="heun3", n=2, lengths=(9, 14), T=20, label="tiny")


def c_bf16_simt():
    check_estimator("bf16", 1 | 2 | 8, label="tiny simt-gemm simt-attn")
    check_solve("bf16", 1 | 2 | 8, label="tiny simt-gemm simt-attn")


def c_debug_gemm():
    import torch
    import matcha_tts_24k_b200 as P
    ora, m = _models(TINY, "euler", "bf16", 1)
    m.refresh(torch.device("cuda", 0))
    lib, h = m._lib, m._handle
    g = torch.Generator().manual_seed(0)
    for (M, N, K, taps) in [(128, 64, 64, 1), (256, 128, 128, 1), (300, 192, 256, 1), (1000, 384, 384, 3), (517, 160, 320, 3),
                            (4096, 1536, 384, 1), (333, 100, 384, 1), (2048, 384, 1536, 1)]:
        a = (torch.randn(M, K, generator=g)).bfloat16().cuda()
        w = (torch.randn(taps * N, K, generator=g) / K ** 0.5).bfloat16().cuda()
        shifts = [0] if taps == 1 else [-1, 0, 1]
        sh = (C.c_int32 * taps)(*shifts)
        d_tc = torch.full((M, N), float("nan"), device="cuda")
        d_si = torch.full((M, N), float("nan"), device="cuda")
        P.native.check(lib, h, lib.cfm_debug_gemm(h, a.data_ptr(), w.data_ptr(), d_si.data_ptr(), M, N, K, taps, sh, 0, None))
        P.native.check(lib, h, lib.cfm_debug_gemm(h, a.data_ptr(), w.data_ptr(), d_tc.data_ptr(), M, N, K, taps, sh, 1, None))
        torch.cuda.synchronize()
        ref = torch.zeros(M, N, dtype=torch.float64)
        ad, wd = a.double().cpu(), w.double().cpu()
        for t, s in enumerate(shifts):
            sa = torch.zeros_like(ad)
            if s == 0: sa = ad
            elif s < 0: sa[-s:] = ad[:s]
            else: sa[:-s] = ad[s:]
            ref += sa @ wd[t * N:(t + 1) * N].T
        print(f"[gemm M={M} N={N} K={K} taps={taps}] simt_vs_ref={rel(d_si, ref):.3e} tc_vs_ref={rel(d_tc, ref):.3e} "
              f"tc_nan={int(torch.isnan(d_tc).sum())}")


def c_bf16_tc_gemm():
    check_estimator("bf16", 1 | 8 | 4, label="tiny tc-gemm unfused-stats simt-attn")
    check_estimator("bf16", 1 | 8, label="tiny tc-gemm fused-stats simt-attn")
    check_estimator("bf16", 1 | 8, lengths=(9, 30, 1, 2), T=64, label="tiny-padded tc-gemm simt-attn")
    check_solve("bf16", 8, label="tiny tc-gemm graph simt-attn")


def c_bf16_tc_attn():
    check_estimator("bf16", 1 | 2, dec=TINY64, label="tiny64 simt-gemm TC-attn")
    check_estimator("bf16", 1 | 2 | 8, dec=TINY64, label="tiny64 simt-gemm simt-attn (baseline)")
    check_estimator("bf16", 1 | 2, dec=TINY64, lengths=(300, 129, 5), T=300, label="tiny64-long simt-gemm TC-attn")
    check_estimator("bf16", 1 | 2 | 8, dec=TINY64, lengths=(300, 129, 5), T=300, label="tiny64-long simt-gemm simt-attn")
    check_estimator("bf16", 1 | 2, dec=TINY64, lengths=(300, 129, 5), T=400, label="tiny64-long-padded simt-gemm TC-attn")
    check_estimator("bf16", 1 | 2 | 8, dec=TINY64, lengths=(300, 129, 5), T=400, label="tiny64-long-padded simt-gemm simt-attn")
    check_solve("bf16", 0, dec=TINY64, lengths=(300, 129, 5), T=300, label="tiny64 all-tc graph")


def c_prod():
    import torch
    import matcha_tts_24k_b200 as P
    for flags, label in ((8, "tc-gemm simt-attn"), (0, "all-tc")):
        check_solve("bf16", flags, dec=P.synthetic.PROD, lengths=(120, 77), T=120, n=2, label="prod " + label)
    check_solve("fp32", 0, dec=P.synthetic.PROD, lengths=(60, 41), T=60, n=2, label="prod fp32")


def c_time():
    import types
    import torch
    import matcha_tts_24k_b200 as P
    for flags, label in ((8, "tc-gemm simt-attn"), (0, "all-tc")):
        cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
        m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=flags).eval().cuda()
        P.synthetic.fill_named_seed(m.estimator, 1234)
        lengths = [938] * 32
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
        ts = torch.linspace(0, 1, 11, device="cuda")
        for _ in range(2):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fl = P.synthetic.algorithmic_flops(lengths, 384, 10)
        print(f"[time cfg2 {label}] {ms:.2f} ms/solve  {sum(lengths)/ms*1e3:.3e} frames/s  {fl/ms/1e9:.1f} TFLOP/s "
              f"finite={bool(torch.isfinite(out).all())} info={m.plan_info()}")


def c_epimodes():
    """cfg2 decode time against which epilogue modes use the TMA-store epilogue (bit m = EpiMode m) and the pair policy."""
    import types
    import torch
    import matcha_tts_24k_b200 as P
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16").eval().cuda()
    P.synthetic.fill_named_seed(m.estimator, 1234)
    m.refresh(torch.device("cuda", 0))
    lengths = P.synthetic.config_lengths(os.environ.get("DIAG_WORKLOAD", "cfg2"))
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
    ts = torch.linspace(0, 1, 11, device="cuda")
    combos = os.environ.get("DIAG_COMBOS", "0:1,1:1,4:1,8:1,16:1,63:1,63:2,0:2,0:0")
    for item in combos.split(","):
        f = [int(v) for v in item.split(":")]
        tma, pair = f[0], f[1]
        m.set_option("tma_epi", tma)
        m.set_option("pair_mode", pair)
        m.set_option("pdl", f[2] if len(f) > 2 else 0)
        m.set_option("small_tiles", f[3] if len(f) > 3 else 1024)
        m.set_option("cluster", f[4] if len(f) > 4 else 1)
        m.set_option("direct_epi", f[5] if len(f) > 5 else 0)
        m.set_option("pair_n256", f[6] if len(f) > 6 else 0)
        m.set_option("l2_persist_mb", f[7] if len(f) > 7 else 0)
        m.set_option("snake_warps", f[8] if len(f) > 8 else 12)
        reps = int(os.environ.get("DIAG_REPS", "5"))
        warm_s = float(os.environ.get("DIAG_WARM_S", "0"))  # > 0: sustained measurement (the decode is power-capped after ~1 s)
        import time as _t
        t_w = _t.perf_counter()
        n_w = 0
        while n_w < 2 or _t.perf_counter() - t_w < warm_s:
            out = m.solve(z, ts, mu, mask, lengths=lengths)
            torch.cuda.synchronize()
            n_w += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"[epimodes tma_mask={tma} pair_mode={pair} pdl={f[2] if len(f) > 2 else 0} small_tiles={f[3] if len(f) > 3 else 1024} cluster={f[4] if len(f) > 4 else 1} direct_epi={f[5] if len(f) > 5 else 0} pair_n256={f[6] if len(f) > 6 else 0} l2_persist_mb={f[7] if len(f) > 7 else 0} snake_warps={f[8] if len(f) > 8 else 12}] {ms:.2f} ms/solve {P.synthetic.algorithmic_flops(lengths, 384, 10)/ms/1e9:.1f} TFLOP/s "
              f"finite={bool(torch.isfinite(out).all())}", flush=True)


def c_steptime():
    """Time of individual launches of the first estimator stage on cfg2 (direct launches): t(stop_after=k) - t(stop_after=k-1).
    Schedule: 1 memset, 2 conv1+stats, 3 res_conv, 4 gn_apply, 5 conv2+stats, 6 gn_apply+LN, 7 QKV, 8 attention, 9 out-proj,
    10 LN, 11 FF1+snake, 12 FF2."""
    import types
    import torch
    import matcha_tts_24k_b200 as P
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=1).eval().cuda()
    P.synthetic.fill_named_seed(m.estimator, 1234)
    lengths = P.synthetic.config_lengths(os.environ.get("DIAG_WORKLOAD", "cfg2"))
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
    t = torch.tensor(0.3)
    m.estimator(z, mask, mu, t)
    torch.cuda.synchronize()
    names = {1: "memset", 2: "conv1+stats", 3: "res_conv", 4: "gn_apply", 5: "conv2+stats", 6: "gn_apply+LN", 7: "QKV", 8: "attention",
             9: "out-proj", 10: "LN", 11: "FF1+snake", 12: "FF2"}
    prev = None
    for k in range(0, 13):
        m._lib.cfm_debug_stop_after(m._handle, k)
        for _ in range(3):
            m.estimator(z, mask, mu, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m.estimator(z, mask, mu, t)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        if prev is not None:
            print(f"[steptime {k:2d} {names.get(k, '?'):12s}] {us - prev:7.1f} us", flush=True)
        prev = us
    m._lib.cfm_debug_stop_after(m._handle, -1)


def c_lanes():
    """Decode time of cfg2 / cfg3 / cfg4 against the number of lanes (parallel graph branches)."""
    import types
    import torch
    import matcha_tts_24k_b200 as P
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16").eval().cuda()
    P.synthetic.fill_named_seed(m.estimator, 1234)
    names = os.environ.get("DIAG_WORKLOADS", "cfg2,cfg4,cfg3").split(",")
    for name in names:
        lengths = P.synthetic.config_lengths(name)
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
        ts = torch.linspace(0, 1, 11, device="cuda")
        ref = None
        for lanes in [int(v) for v in os.environ.get("DIAG_LANES", "1,2,3,4,6").split(",")]:
            m.set_lanes(lanes, 2048)
            for _ in range(2):
                out = m.solve(z, ts, mu, mask, lengths=lengths)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if name != "cfg3" else 2
            e0.record()
            for _ in range(reps):
                out = m.solve(z, ts, mu, mask, lengths=lengths)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            fl = P.synthetic.algorithmic_flops(lengths, 384, 10)
            if ref is None:
                ref = out.clone()
            print(f"[lanes {name} lanes={lanes}] {ms:.2f} ms/solve  {sum(lengths)/ms*1e3:.3e} frames/s  {fl/ms/1e9:.1f} TFLOP/s "
                  f"rel_vs_1lane={rel(out, ref):.2e} finite={bool(torch.isfinite(out).all())} kernels={m.plan_info()['kernels_per_solve']}", flush=True)


def c_trace():
    """Bisect the schedule: compare intermediate buffers (fp32 mode, SIMT kernels; the stop indices below were derived for the
    round-1a schedule of 7 launches per resnet / block and must be re-derived from emit_nfe when the schedule changes)
    against hooks on the dense oracle.  Valid rows per utterance (+ the pad-token row for the residual stream)."""
    import torch
    import matcha_tts_24k_b200 as P
    lengths, T = [40, 33, 17], 40
    ora, m = _models(TINY, "euler", "fp32", 1)
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=7, T=T)
    caps = {}
    est = ora.estimator
    def hook(name, tr=False):
        def f(mod, inp, out):
            caps[name] = (out.transpose(1, 2) if not tr else out).detach()
        return f
    est.down_blocks[0][0].register_forward_hook(hook("d0.res"))
    est.down_blocks[0][1][0].register_forward_hook(hook("d0.tb0", True))
    est.down_blocks[0][1][1].register_forward_hook(hook("d0.tb1", True))
    est.down_blocks[0][2].register_forward_hook(hook("d0.tail"))
    est.down_blocks[1][0].register_forward_hook(hook("d1.res"))
    est.down_blocks[1][1][1].register_forward_hook(hook("d1.tb1", True))
    est.down_blocks[1][2].register_forward_hook(hook("d1.tail"))
    est.mid_blocks[1][1][1].register_forward_hook(hook("m1.tb1", True))
    est.up_blocks[0][1][1].register_forward_hook(hook("u0.tb1", True))
    est.up_blocks[0][2].register_forward_hook(hook("u0.tail"))
    est.up_blocks[1][1][1].register_forward_hook(hook("u1.tb1", True))
    est.up_blocks[1][2].register_forward_hook(hook("u1.tail"))
    est.final_block.register_forward_hook(hook("final_block"))
    with torch.inference_mode():
        v_ref = est(z, mask, mu, torch.tensor(0.3))
    lib, hd = None, None
    pts = [(8, "X1", "d0.res", 0, True), (15, "X1", "d0.tb0", 0, True), (22, "X1", "d0.tb1", 0, True),
           (23, "sin2", "d0.tail", 1, False), (30, "X2", "d1.res", 1, True), (44, "X2", "d1.tb1", 1, True),
           (45, "sin2", "d1.tail", 1, False), (87, "X2", "m1.tb1", 1, True), (108, "X2", "u0.tb1", 1, True),
           (110, "cat1", "u0.tail", 0, False), (131, "X1", "u1.tb1", 0, True), (132, "hact1", "u1.tail", 0, False),
           (135, "sin1", "final_block", 0, False)]
    zc, mc, muc = z.cuda(), mask.cuda(), mu.cuda()
    m.estimator(zc, mc, muc, torch.tensor(0.3))
    lib, hd = m._lib, m._handle
    for stop, buf, name, res, with_pad in pts:
        lib.cfm_debug_stop_after(hd, stop)
        m.estimator(zc, mc, muc, torch.tensor(0.3))
        got = m.debug_read(buf)
        ref = caps[name]  # (B, T_res, C)
        C_ = ref.shape[2]
        msgs = []
        start = 0
        for b, L in enumerate(lengths):
            L2 = (L + 1) // 2
            Lr = L if res == 0 else L2
            s = 2 * start if res == 0 else start
            g = got[s:s + Lr, :C_]
            r = ref[b, :Lr]
            e = rel(g, r)
            ep = float("nan")
            Tr = ref.shape[1]
            if with_pad and Lr < Tr:
                ep = rel(got[s + Lr, :C_], ref[b, Lr])
            msgs.append(f"utt{b}(L={Lr}) valid={e:.2e} pad={ep:.2e}")
            start += L2 + 2
        print(f"[trace stop={stop} {buf} vs {name}] " + "  ".join(msgs))
    lib.cfm_debug_stop_after(hd, -1)
    v = m.estimator(zc, mc, muc, torch.tensor(0.3))
    for b, L in enumerate(lengths):
        print(f"[trace final v utt{b}] rel={rel(v[b, :, :L], v_ref[b, :, :L]):.3e}")
    print("sinemb", rel(m.debug_read("sinemb")[0], __import__("oracle.cfm_oracle", fromlist=["x"]).sinusoidal_embedding(torch.tensor(0.3), 200)[0]))


def c_perm():
    """Which component makes results depend on utterance order?"""
    import torch
    import matcha_tts_24k_b200 as P
    lengths = [400, 123, 398, 57, 256, 311]
    perm = [4, 0, 5, 2, 1, 3]
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=9, T=400)
    for prec, flags, label in (("fp32", 0, "fp32 simt"), ("bf16", 2 | 8, "bf16 simt-gemm simt-attn"), ("bf16", 2, "bf16 simt-gemm TC-attn"),
                               ("bf16", 8 | 4, "bf16 TC-gemm unfused simt-attn"), ("bf16", 8, "bf16 TC-gemm fused simt-attn"),
                               ("bf16", 0, "bf16 all-tc")):
        for dec, dl in ((TINY64, "tiny64"), (P.synthetic.PROD, "prod")):
            if prec == "fp32" and dl == "prod":
                continue
            ora, m = _models(dec, "euler", prec, flags | 1)
            t = torch.tensor(0.4)
            full = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t).cpu()
            again = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t).cpu()
            sh = m.estimator(z[perm].cuda(), mask[perm].cuda(), mu[perm].cuda(), t).cpu()
            errs = [rel(sh[j], full[i]) for j, i in enumerate(perm)]
            print(f"[perm {label} {dl}] repeat={rel(again, full):.2e} perm_max={max(errs):.2e} per-utt=" + " ".join(f"{e:.1e}" for e in errs))


def c_stats():
    """Compare the fused-epilogue GroupNorm sums against the stand-alone statistics kernel, site 0 (first conv)."""
    import torch
    import matcha_tts_24k_b200 as P
    lengths = [400, 123, 398, 57, 256, 311]
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=9, T=400)
    res = {}
    for flags, stop, label in ((1 | 8 | 4, 3, "unfused"), (1 | 8, 2, "fused")):
        ora, m = _models(TINY64, "euler", "bf16", flags)
        t = torch.tensor(0.4)
        m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t)
        m._lib.cfm_debug_stop_after(m._handle, stop)
        m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t)
        res[label] = m.debug_read("stats")[:len(lengths)].clone()
        res[label + "_h"] = m.debug_read("hraw1").clone()
    torch.set_printoptions(linewidth=200, precision=6)
    print("hraw equal:", rel(res["fused_h"], res["unfused_h"]))
    d = (res["fused"] - res["unfused"]).abs() / res["unfused"].abs().clamp_min(1e-6)
    for b in range(len(lengths)):
        print(f"utt{b} L={lengths[b]} max rel diff={float(d[b].max()):.3e}\n  fused  ={res['fused'][b][:6].tolist()}\n  unfused={res['unfused'][b][:6].tolist()}")


def c_stats_all():
    import torch
    import matcha_tts_24k_b200 as P
    lengths = [400, 123, 398, 57, 256, 311]
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=9, T=400)
    res = {}
    for flags, label in ((1 | 8 | 4, "unfused"), (1 | 8, "fused")):
        ora, m = _models(TINY64, "euler", "bf16", flags)
        m.estimator(z.cuda(), mask.cuda(), mu.cuda(), torch.tensor(0.4))
        res[label] = m.debug_read("stats").clone().view(-1, len(lengths), 16)
    d = (res["fused"] - res["unfused"]).abs() / res["unfused"].abs().clamp_min(1e-3)
    for site in range(d.shape[0]):
        print(f"site {site}: " + " ".join(f"utt{b}={float(d[site, b].max()):.1e}" for b in range(len(lengths))))
        bad = (d[site] > 1e-3).nonzero()
        for b, k in bad[:4].tolist():
            print(f"    utt{b} k={k} fused={float(res['fused'][site, b, k]):.6f} unfused={float(res['unfused'][site, b, k]):.6f}")


def c_cluster():
    """TMA-multicast cluster sizes 1/2/4: GEMM correctness vs the fp32-FMA kernel, then cfg2 decode time."""
    import types
    import torch
    import matcha_tts_24k_b200 as P
    for cl, pair in ((1, 0), (1, 1), (1, 2)):
        os.environ["CFM_B200_CLUSTER"] = str(cl)
        os.environ["CFM_B200_PAIR"] = str(pair)
        cl = f"{cl} pair={pair}"
        ora, m = _models(TINY, "euler", "bf16", 1)
        m.refresh(torch.device("cuda", 0))
        lib, h = m._lib, m._handle
        g = torch.Generator().manual_seed(0)
        worst = 0.0
        for (M, N, K, taps) in [(128, 64, 64, 1), (300, 192, 256, 1), (1000, 384, 384, 3), (517, 160, 320, 3), (4096, 1536, 384, 1),
                                (333, 100, 384, 1), (2048, 384, 1536, 1), (30144, 384, 384, 3), (30144, 1152, 384, 1)]:
            a = torch.randn(M, K, generator=g).bfloat16().cuda()
            w = (torch.randn(taps * N, K, generator=g) / K ** 0.5).bfloat16().cuda()
            shifts = [0] if taps == 1 else [-1, 0, 1]
            sh = (C.c_int32 * taps)(*shifts)
            d = [torch.full((M, N), float("nan"), device="cuda") for _ in range(2)]
            for use_tc in (0, 1):
                P.native.check(lib, h, lib.cfm_debug_gemm(h, a.data_ptr(), w.data_ptr(), d[use_tc].data_ptr(), M, N, K, taps, sh, use_tc, None))
            torch.cuda.synchronize()
            worst = max(worst, rel(d[1], d[0]))
        print(f"[cluster {cl}] gemm tc-vs-simt worst rel = {worst:.3e}")
        m.close()
        cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
        m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=0).eval().cuda()
        P.synthetic.fill_named_seed(m.estimator, 1234)
        lengths = [938] * 32
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
        ts = torch.linspace(0, 1, 11, device="cuda")
        for _ in range(2):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"[cluster {cl}] cfg2 {ms:.2f} ms/solve {P.synthetic.algorithmic_flops(lengths, 384, 10)/ms/1e9:.1f} TFLOP/s finite={bool(torch.isfinite(out).all())}")
        m.close()


def c_gemmprof():
    """Stand-alone GEMM timings (CUDA events, 20 reps) + CTA-0 role cycle counters for the cfg2 shapes."""
    import torch
    import matcha_tts_24k_b200 as P
    ora, m = _models(TINY, "euler", "bf16", 1)
    m.refresh(torch.device("cuda", 0))
    lib, h = m._lib, m._handle
    g = torch.Generator().manual_seed(0)
    M = 30144
    for key in ("tma_epi", "pair_mode"):
        if os.environ.get("DIAG_" + key.upper()) is not None:
            m.set_option(key, int(os.environ["DIAG_" + key.upper()]))
    shapes = [("QKV bf16", 1152, 384, 1, 0), ("QKV f32out", 1152, 384, 1, 1), ("out-proj resid", 384, 384, 1, 2), ("out-proj f32", 384, 384, 1, 1),
              ("out-proj bf16", 384, 384, 1, 0), ("conv2 f32", 384, 384, 3, 1), ("conv2 f32+stats", 384, 384, 3, 3), ("conv2 bf16", 384, 384, 3, 0), ("FF2 resid", 384, 1536, 1, 2),
              ("FF2 bf16", 384, 1536, 1, 0), ("FF1 bf16", 1536, 384, 1, 0)]
    for name, N, K, taps, mode in shapes:
        a = torch.randn(M, K, generator=g).bfloat16().cuda()
        w = (torch.randn(taps * N, K, generator=g) / K ** 0.5).bfloat16().cuda()
        shifts = [0] if taps == 1 else [-1, 0, 1]
        sh = (C.c_int32 * taps)(*shifts)
        d32 = torch.zeros(M, N, device="cuda")
        d16 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        call = lambda pr: P.native.check(lib, h, lib.cfm_debug_gemm_profile(h, a.data_ptr(), w.data_ptr(), d32.data_ptr(), d16.data_ptr(),
                                                                        M, N, K, taps, sh, mode, pr, None))
        for _ in range(3):
            call(None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call(None)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        call(prof.data_ptr())
        torch.cuda.synchronize()
        pr = prof.tolist()
        fl = 2.0 * M * N * K * taps
        print(f"[gemmprof {name:16s} N={N} K={K}x{taps}] {us:7.1f} us {fl/us/1e6:7.1f} TFLOP/s | tiles(cta0)={pr[8]} prod tot={pr[0]} wait_empty={pr[1]} | "
              f"mma tot={pr[2]} wait_full={pr[3]} wait_tempty={pr[4]} | epi tot={pr[5]} wait_tfull={pr[6]}")


def c_tmaepi():
    """TMA-store epilogue vs the transposing epilogue: results (bf16 store, in-place fp32 residual add) and stand-alone
    timings for the cfg2 GEMM shapes, 1-CTA and CTA-pair kernels."""
    import torch
    import matcha_tts_24k_b200 as P
    ora, m = _models(TINY, "euler", "bf16", 1)
    m.refresh(torch.device("cuda", 0))
    lib, h = m._lib, m._handle
    g = torch.Generator().manual_seed(0)
    shapes = [("QKV", 1152, 384, 1, 0), ("FF1-store", 1536, 384, 1, 0), ("out-proj resid", 384, 384, 1, 2), ("FF2 resid", 384, 1536, 1, 2),
              ("conv bf16", 384, 384, 3, 0), ("N=320 resid", 320, 320, 1, 2), ("N=1280 store", 1280, 320, 1, 0), ("N=960 store", 960, 320, 1, 0)]
    for M in (30144, 1000):
        for name, N, K, taps, mode in shapes:
            a = torch.randn(M, K, generator=g).bfloat16().cuda()
            w = (torch.randn(taps * N, K, generator=g) / K ** 0.5).bfloat16().cuda()
            shifts = [0] if taps == 1 else [-1, 0, 1]
            sh = (C.c_int32 * taps)(*shifts)
            base = torch.randn(M, N, generator=g).cuda()
            outs = {}
            line = f"[tmaepi M={M} {name:14s} N={N} K={K}x{taps}]"
            for tma in (0, 1):
                for pair in (0, 2):
                    m.set_option("tma_epi", tma)
                    m.set_option("pair_mode", pair)
                    d32 = base.clone()
                    d16 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
                    call = lambda: P.native.check(lib, h, lib.cfm_debug_gemm_profile(h, a.data_ptr(), w.data_ptr(), d32.data_ptr(), d16.data_ptr(),
                                                                                   M, N, K, taps, sh, mode, None, None))
                    call()
                    torch.cuda.synchronize()
                    outs[(tma, pair)] = (d16.float() if mode == 0 else d32).clone()
                    for _ in range(3):
                        call()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        call()
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) / 20 * 1e3
                    line += f" | tma={tma} pair={pair}: {us:6.1f} us {2.0 * M * N * K * taps / us / 1e6:6.1f} TF"
            ref = outs[(0, 0)]
            worst = max(rel(v, ref) for v in outs.values())
            exact = all(torch.equal(v, ref) for v in outs.values())
            print(line + f" | worst rel vs old={worst:.2e} bitwise={exact} finite={bool(torch.isfinite(ref).all())}", flush=True)


def c_attnprof():
    """CTA-0 cycle counters of attn_tc_kernel on cfg2 (first full-resolution attention of an estimator call)."""
    import types
    import torch
    import matcha_tts_24k_b200 as P
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=1).eval().cuda()
    P.synthetic.fill_named_seed(m.estimator, 1234)
    lengths = [938] * 32
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
    m.estimator(z, mask, mu, torch.tensor(0.3))
    prof = torch.zeros(32 + 4 * 4096, dtype=torch.int64, device="cuda")
    m._lib.cfm_debug_attn_profile(m._handle, prof.data_ptr())
    m._lib.cfm_debug_stop_after(m._handle, 12)  # memset, conv1, res, gn, conv2, gn, LN, QKV, attention
    m.estimator(z, mask, mu, torch.tensor(0.3))
    torch.cuda.synchronize()
    p = prof.tolist()
    print(f"[attnprof] tiles={p[5]} | mma thread: total={p[0]} wait_q={p[1]} wait_k={p[2]} wait_p={p[3]} wait_v={p[4]} | "
          f"softmax warp: total={p[8]} wait_s={p[9]} wait_rowmax_barrier={p[10]} wait_pv={p[11]}")
    # per-CTA lifetimes: the grid is 32 utterances x 6 heads x 8 query tiles = 1536 CTAs, two resident per SM
    n = 1536
    cyc = [p[32 + 4 * i] for i in range(n)]
    t0 = [p[32 + 4 * i + 1] for i in range(n)]
    t1 = [p[32 + 4 * i + 2] for i in range(n)]
    base = min(t0)
    import statistics as st
    print(f"[attnprof] per-CTA MMA-warp lifetime [cycles]: min {min(cyc)} median {st.median(cyc)} mean {st.mean(cyc):.0f} max {max(cyc)}; "
          f"kernel span {(max(t1) - base) / 1e3:.1f} us")
    order = sorted(range(n), key=lambda i: t0[i])
    for lo in range(0, n, 296):
        grp = order[lo:lo + 296]
        print(f"   CTAs started {lo}-{lo + len(grp) - 1}: start {(min(t0[i] for i in grp) - base) / 1e3:6.1f}-{(max(t0[i] for i in grp) - base) / 1e3:6.1f} us, "
              f"lifetime {st.mean((t1[i] - t0[i]) for i in grp) / 1e3:5.1f} us mean, {st.mean(cyc[i] for i in grp):.0f} cycles mean")
    last = sorted(t1)
    print(f"   50 % of the CTAs done at {(last[n // 2] - base) / 1e3:.1f} us, 90 % at {(last[int(n * 0.9)] - base) / 1e3:.1f} us, all at {(last[-1] - base) / 1e3:.1f} us")


def c_plantime():
    """Host-side cost of cfm_load_weights and cfm_plan (row tables, workspace, graph capture) for cfg1 and cfg2."""
    import time as _t
    import types
    import torch
    import matcha_tts_24k_b200 as P
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16").eval().cuda()
    P.synthetic.fill_named_seed(m.estimator, 1234)
    t0 = _t.perf_counter(); m.refresh(torch.device("cuda", 0)); torch.cuda.synchronize(); t1 = _t.perf_counter()
    print(f"[plantime] cfm_load_weights (pack 37 M params): {(t1 - t0) * 1e3:.1f} ms")
    for name in ("cfg1", "cfg2", "cfg3"):
        lengths = P.synthetic.config_lengths(name)
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
        ts = torch.linspace(0, 1, 11, device="cuda")
        torch.cuda.synchronize(); t0 = _t.perf_counter()
        out = m.solve(z, ts, mu, mask, lengths=lengths); torch.cuda.synchronize(); t1 = _t.perf_counter()
        out = m.solve(z, ts, mu, mask, lengths=lengths); torch.cuda.synchronize(); t2 = _t.perf_counter()
        out = m.solve(z, ts, mu, mask, lengths=lengths); torch.cuda.synchronize(); t3 = _t.perf_counter()
        print(f"[plantime] {name}: first call (plan + direct-launch decode) {(t1 - t0) * 1e3:.1f} ms, second call (graph capture + decode) "
              f"{(t2 - t1) * 1e3:.1f} ms, third call (graph replay) {(t3 - t2) * 1e3:.1f} ms, {m.plan_info()}")
        for k, ga in enumerate((0, 1, 0, 1)):  # a NEW shape per call (what a server sees), both graph policies, warm library
            m.set_option("graph_after", ga)
            lengths2 = [max(1, v - 2 * (k + 1)) for v in lengths]
            mu2, mask2, z2, _ = P.synthetic.make_inputs(lengths2, seed=1, device="cuda")
            torch.cuda.synchronize(); t0 = _t.perf_counter()
            out = m.solve(z2, ts, mu2, mask2, lengths=lengths2); torch.cuda.synchronize(); t1 = _t.perf_counter()
            print(f"[plantime] {name} new shape, graph_after={ga}: plan + first decode {(t1 - t0) * 1e3:.1f} ms")
        m.set_option("graph_after", 1)


CHECKS = {k[2:]: v for k, v in list(globals().items()) if k.startswith("c_")}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        CHECKS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CHECKS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.log"), "a") as log:
        for n in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True, timeout=420)
                tail = "\n".join((p.stdout + "\n" + p.stderr[-3000:]).strip().splitlines()[-40:])
                msg = f"=== {n}: exit {p.returncode} in {time.time()-t0:.0f}s\n{tail}\n"
            except subprocess.TimeoutExpired as e:
                msg = f"=== {n}: TIMEOUT\n{(e.stdout or b'')[-2000:]}\n"
            print(msg, flush=True)
            log.write(msg)
            log.flush()
