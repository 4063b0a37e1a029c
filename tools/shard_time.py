"""One cfg3 shard (as ShardedCFM deals it to 1 of N GPUs) on a single GPU: device-resident decode vs the host-buffer entry."""
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lengths_all = P.synthetic.config_lengths("cfg3")
T = 2 * ((max(lengths_all) + 1) // 2)
idx = P.shard_utterances(lengths_all, n)[0]
lengths = [lengths_all[i] for i in idx]
cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
m = P.CFM(200, 100, cp, P.synthetic.PROD, precision="bf16").eval().cuda()
P.synthetic.fill_named_seed(m.estimator, 1234)
mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, T=T)
ts = torch.linspace(0, 1, 11)
mu_d, mask_d, z_d = mu.cuda(), mask.cuda(), z.cuda()
for _ in range(3):
    m.solve(z_d, ts, mu_d, mask_d, lengths=lengths)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.solve(z_d, ts, mu_d, mask_d, lengths=lengths)
e1.record()
torch.cuda.synchronize()
print(f"shard 0 of {n}: {len(lengths)} utterances, {sum(lengths)} frames, T={T}, {m.plan_info()}")
print(f"  device-resident decode {e0.elapsed_time(e1) / 5:.2f} ms")
mu_h, z_h = mu.pin_memory(), z.pin_memory()
out = torch.empty_like(mu).pin_memory()
for _ in range(2):
    m.solve_host(z_h, ts, mu_h, lengths, out=out)
t = []
for _ in range(5):
    t0 = time.perf_counter()
    m.solve_host(z_h, ts, mu_h, lengths, out=out)
    t.append((time.perf_counter() - t0) * 1e3)
print(f"  solve_host (one H2D / D2H of the whole shard) {sorted(t)[2]:.2f} ms")
sh = P.ShardedCFM(200, 100, cp, P.synthetic.PROD, devices=[0], precision="bf16")
P.synthetic.fill_named_seed(sh.replicas[0].estimator, 1234)
for _ in range(3):
    sh.solve_host(z_h, ts, mu_h, lengths, out=out)
t = []
for _ in range(5):
    t0 = time.perf_counter()
    sh.solve_host(z_h, ts, mu_h, lengths, out=out)
    t.append((time.perf_counter() - t0) * 1e3)
print(f"  ShardedCFM on one GPU (per-utterance H2D / D2H) {sorted(t)[2]:.2f} ms")
