"""Fused feed-forward kernel vs the two-GEMM schedule vs the CPU oracle (estimator call), then cfg2 timing."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import matcha_tts_24k_b200 as P  # noqa: E402
from oracle import cfm_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    for name, dec, lengths in (("prod", P.synthetic.PROD, [300, 171, 64]), ("default", P.synthetic.DEFAULT, [257, 130]),
                               ("tiny128", dict(channels=(128, 128), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=1, num_heads=2), [150, 97])):
        ora = O.CFM(200, 100, cp, dec).eval()
        P.synthetic.fill_named_seed(ora.estimator, 1234)
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=3)
        with torch.inference_mode():
            ref = ora.estimator(z, mask, mu, torch.tensor(0.4))
        m = P.CFM(200, 100, cp, dec, precision="bf16").eval()
        m.estimator.load_state_dict(ora.estimator.state_dict())
        m = m.cuda()
        m.refresh(torch.device("cuda", 0))
        outs = {}
        for fused in (0, 1):
            m.set_option("small_tiles", 0)
            m.set_option("ff_fused", fused)
            outs[fused] = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), torch.tensor(0.4)).cpu()
            torch.cuda.synchronize()
        print(f"[{name}] unfused vs oracle {rel(outs[0], ref):.3e}  fused vs oracle {rel(outs[1], ref):.3e}  fused vs unfused {rel(outs[1], outs[0]):.3e} "
              f"finite={bool(torch.isfinite(outs[1]).all())}", flush=True)
        m.close()


if __name__ == "__main__":
    main()
