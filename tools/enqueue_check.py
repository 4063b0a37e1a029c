"""Host-side enqueue time of back-to-back decodes (graph replays) after a device synchronisation, per position in the burst.
Patterns: 'plain' (no events), 'marks' (a torch timing event recorded after every decode), for attn_persist 1 / 0."""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

wl = os.environ.get("AB_WORKLOAD", "cfg2")
reps, burst = int(os.environ.get("REPS", "12")), 6
cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16").eval().cuda()
P.synthetic.fill_named_seed(m.estimator, 1234)
m.refresh(torch.device("cuda", 0))
lengths = P.synthetic.config_lengths(wl)
mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
ts = torch.linspace(0, 1, 5 if wl == "cfg5" else 11)
for persist in (1, 0, 1, 0):
    m.set_option("attn_persist", persist)
    for _ in range(3):
        m.solve(z, ts, mu, mask, lengths=lengths)
        torch.cuda.synchronize()
    for pattern in ("plain", "marks"):
        worst = [0.0] * burst
        total = []
        for _ in range(reps):
            torch.cuda.synchronize()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(burst + 1)]
            t_all = time.perf_counter()
            if pattern == "marks":
                evs[0].record()
            for i in range(burst):
                t0 = time.perf_counter()
                m.solve(z, ts, mu, mask, lengths=lengths)
                worst[i] = max(worst[i], (time.perf_counter() - t0) * 1e3)
                if pattern == "marks":
                    evs[i + 1].record()
            torch.cuda.synchronize()
            total.append((time.perf_counter() - t_all) * 1e3)
        print(f"[{wl}] persist={persist} {pattern:5s}: worst enqueue ms per position {[round(v, 2) for v in worst]}  burst wall ms min {min(total):.1f} max {max(total):.1f}", flush=True)
