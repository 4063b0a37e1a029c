"""In-kernel cycle counters of gemm_rowln_kernel (CTA 0) on one estimator evaluation of a workload (direct launches)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=1).eval().cuda()
P.synthetic.fill_named_seed(m.estimator, 1234)
m.refresh(torch.device("cuda", 0))
m.set_option("rowln", 2)
lengths = P.synthetic.config_lengths(os.environ.get("AB_WORKLOAD", "cfg2"))
mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
t = torch.tensor(0.3)
m.estimator(z, mask, mu, t)
prof = torch.zeros(64, dtype=torch.int64, device="cuda")
m._lib.cfm_debug_rowln_profile(m._handle, prof.data_ptr())
for stop in [int(x) for x in os.environ.get("STOPS", "9,13,10000").split(",")]:
    prof.zero_()
    m._lib.cfm_debug_stop_after(m._handle, stop)
    m.estimator(z, mask, mu, t)
    torch.cuda.synchronize()
    v = prof.tolist()
    print(f"[stop_after={stop}] tiles(CTA0)={v[10]}")
    print(f"  producer total {v[0]}  wait empty {v[1]}")
    print(f"  mma      total {v[2]}  wait full {v[3]}  wait tempty {v[4]}")
    print(f"  epilogue total {v[5]}  wait tfull {v[6]}  pass1 {v[7]}  exchange {v[8]}  pass2 {v[9]}")
