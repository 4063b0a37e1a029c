"""Per-decode device times of N graph replays for kernel-selection settings: median, max, number of outliers (> 1.15 x median).

    python tools/outlier_check.py N "attn_persist=1" "attn_persist=0" ...        AB_WORKLOAD=cfg2|cfg4|...
"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

n = int(sys.argv[1])
wl = os.environ.get("AB_WORKLOAD", "cfg2")
cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16").eval().cuda()
P.synthetic.fill_named_seed(m.estimator, 1234)
m.refresh(torch.device("cuda", 0))
lengths = P.synthetic.config_lengths(wl)
mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
ts = torch.linspace(0, 1, 11)
sampler = None
if os.environ.get("WITH_SAMPLER"):  # the benchmark's clock sampler: does an nvidia-smi poll every 100 ms disturb the decode?
    import subprocess
    sampler = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                                "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                "--format=csv,noheader,nounits", "-lms", os.environ["WITH_SAMPLER"]], stdout=subprocess.DEVNULL)
for rep in range(2):
    for setting in sys.argv[2:]:
        for kv in setting.split(","):
            k, v = kv.split("=")
            m.set_option(k, int(v))
        for _ in range(4):
            m.solve(z, ts, mu, mask, lengths=lengths)
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            m.solve(z, ts, mu, mask, lengths=lengths)
            evs[i + 1].record()
        torch.cuda.synchronize()
        t = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(n))
        med = t[n // 2]
        out = [round(x, 1) for x in t if x > 1.15 * med]
        print(f"[{wl}] [{setting}] n={n} median {med:.2f} ms  min {t[0]:.2f}  max {t[-1]:.2f}  outliers(>1.15x): {len(out)} {out[-8:]}", flush=True)
if sampler:
    sampler.terminate()
