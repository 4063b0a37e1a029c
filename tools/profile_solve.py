"""Short profiling target for ncu: N decodes of a BASELINE workload (default cfg2, bf16), direct launches (no graph)
so that every kernel is a plain launch.  python tools/profile_solve.py [workload] [n_solves] [flags]"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lengths = P.synthetic.config_lengths(name)
cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16", flags=flags).eval()
P.synthetic.fill_named_seed(m.estimator, 1234)
m = m.cuda()
mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device="cuda")
ts = torch.linspace(0, 1, 11, device="cuda")
for _ in range(n):
    out = m.solve(z, ts, mu, mask, lengths=lengths)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(out).all()), m.plan_info())
