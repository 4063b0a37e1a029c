"""In-kernel cycle counters of ff_fused_kernel (CTA 0) on one estimator evaluation of a workload (direct launches)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=1).eval().cuda()
P.synthetic.fill_named_seed(m.estimator, 1234)
lengths = P.synthetic.config_lengths(os.environ.get("AB_WORKLOAD", "cfg2"))
mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
t = torch.tensor(0.3)
m.estimator(z, mask, mu, t)
prof = torch.zeros(512, dtype=torch.int64, device="cuda")
m._lib.cfm_debug_ff_profile(m._handle, prof.data_ptr())
for stop in (12, 10000):  # 12 launches = through the first full-resolution feed-forward; all = the last (full-resolution) one
    m._lib.cfm_debug_stop_after(m._handle, stop)
    m.estimator(z, mask, mu, t)
    torch.cuda.synchronize()
    v = prof.tolist()
    print(f"[stop_after={stop}] tiles(CTA0)={v[15]}")
    print(f"  producer total {v[0]}  wait a_empty {v[1]}  w1_empty {v[2]}  w2_empty {v[3]}")
    print(f"  mma1     total {v[4]}  wait a_full {v[5]}  w1_full {v[6]}  h_free {v[14]}")
    print(f"  mma2     total {v[16]}  wait p_full {v[7]}  w2_full {v[8]}  y_free {v[9]}")
    print(f"  epilogue total {v[10]}  wait h_full {v[11]}  y_full {v[12]}  tail {v[13]}  per-chunk: tmem ld {v[17]}  math {v[18]}  st+arrive {v[19]}")

base = min(x for x in v[32:96] if x > 0)
print("chunk:  mma1 start(after h_free)  mma1 issued  epi start(h_full seen)  epi done(p arrive)  mma2 start(p,w2 seen)  mma2 issued   [cycles from first event]")
for c in range(8):
    r = [x - base for x in v[32 + c * 8: 32 + c * 8 + 6]]
    print(f"  g={8 + c}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} {r[4]:7d} {r[5]:7d}   p_full seen {v[32 + c * 8 + 6] - base:7d}")

