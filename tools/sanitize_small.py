"""Tiny decode for compute-sanitizer (memcheck): every kernel family once, bf16 tensor-core path and fp32 path,
ragged batch.   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

dec = dict(channels=(128, 128), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=1, num_heads=2)
for precision, solver in (("bf16", "euler"), ("bf16", "rk4"), ("fp32", "midpoint")):
    cp = types.SimpleNamespace(solver=solver, sigma_min=1e-4, use_mu_prior=True)
    m = P.CFM(200, 100, cp, dec, precision=precision).eval()
    P.synthetic.fill_named_seed(m.estimator, 3)
    m = m.cuda()
    mu, mask, z, lengths = P.synthetic.make_inputs([150, 97, 5, 131], seed=2, device="cuda", T=160)
    out = m.solve(z, torch.linspace(0, 1, 3, device="cuda"), mu, mask)
    v = m.estimator(z, mask, mu, torch.tensor(0.4))
    torch.cuda.synchronize()
    print(precision, solver, "finite", bool(torch.isfinite(out).all()), bool(torch.isfinite(v).all()), m.plan_info())
    m.close()
print("done")
