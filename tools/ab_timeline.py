"""A/B of kernel-selection options on one workload: graph-replay decode time + in-situ per-launch-class times (cfm_debug_timeline).

    python tools/ab_timeline.py "name=value,name=value" "name=value" ...      (one argument per setting; "" = defaults)
    AB_WORKLOAD=cfg2|cfg3|cfg4  AB_REPS=5  AB_TAGS=qkv,out_proj (only these classes are printed)
"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402


def main():
    settings = sys.argv[1:] or [""]
    wl = os.environ.get("AB_WORKLOAD", "cfg2")
    reps = int(os.environ.get("AB_REPS", "5"))
    only = [t for t in os.environ.get("AB_TAGS", "").split(",") if t]
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, P.synthetic.PROD, precision=os.environ.get("AB_PRECISION", "bf16")).eval().cuda()
    P.synthetic.fill_named_seed(m.estimator, 1234)
    m.refresh(torch.device("cuda", 0))
    lengths = P.synthetic.config_lengths(wl)
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
    ts = torch.linspace(0, 1, 11)
    base = {}
    for setting in settings:
        opts = dict(kv.split("=") for kv in setting.split(",") if kv)
        for k, v in base.items():  # back to defaults first
            if k not in opts:
                m.set_option(k, v)
        for k, v in opts.items():
            m.set_option(k, int(v))
        for _ in range(3):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = m.solve(z, ts, mu, mask, lengths=lengths)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        agg = {}
        for _ in range(2):
            for tag, M, N, K, fl, us in m.timeline(z, ts, mu, lengths):
                key = f"{tag}@{'full' if M > 16384 else 'half'}" if os.environ.get("AB_SPLIT_RES") else tag
                a = agg.setdefault(key, [0.0, 0.0, 0])
                a[0] += us / 2
                a[1] += fl / 2
                a[2] += 1
        tl = sum(a[0] for a in agg.values()) / 1e3
        print(f"=== [{setting or 'defaults'}] {wl}: {ms:.2f} ms/decode (graph), timeline {tl:.2f} ms, finite={bool(torch.isfinite(out).all())}", flush=True)
        for tag, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            if only and tag.split("@")[0] not in only:
                continue
            tf = a[1] / (a[0] * 1e-6) / 1e12 if a[0] > 0 else 0
            print(f"    {tag:24s} {a[0] / 1e3:7.3f} ms  {a[2] // 2:4d} launches  {a[0] / max(1, a[2] // 2):7.1f} us/launch  {tf:7.1f} TFLOP/s", flush=True)
        for k in opts:
            base.setdefault(k, {"bn_full": 0, "bn_half": 0, "pair_min_k": 1024, "pair_mode": 1, "tma_epi": 4, "direct_epi": 17, "snake_warps": 12,
                                "pair_n256": 0, "small_tiles": 1024, "l2_persist_mb": 32, "pdl": -1, "cluster": 1, "attn_persist": 1, "bf16_mid": 1,
                                "rowln_ff2": 1}.get(k, 0))


if __name__ == "__main__":
    main()
