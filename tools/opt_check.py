"""Kernel-selection options vs the default schedule vs the CPU oracle on ragged batches (estimator call + 3-step solve).

    python tools/opt_check.py "attn_pf=1" "ff_fused=1,small_tiles=0" ...
Prints rel-L2 of every setting against the oracle and against the default schedule (0 = bitwise identical).
"""
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
    settings = [""] + sys.argv[1:]
    prec = os.environ.get("OC_PRECISION", "bf16")
    tol = 1.2e-2 if prec == "bf16" else 1e-3
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    cases = (("prod", P.synthetic.PROD, [300, 171, 64, 129, 1]), ("prod_long", P.synthetic.PROD, [938, 517]),
             ("default", P.synthetic.DEFAULT, [257, 130]))
    ok = True
    for name, dec, lengths in cases:
        ora = O.CFM(200, 100, cp, dec).eval()
        P.synthetic.fill_named_seed(ora.estimator, 1234)
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=3)
        ts = torch.linspace(0, 1, 4)
        with torch.inference_mode():
            ref_e = ora.estimator(z, mask, mu, torch.tensor(0.4))
            ref_s = ora.solve(z, ts, mu, mask)
        m = P.CFM(200, 100, cp, dec, precision=prec).eval()
        m.estimator.load_state_dict(ora.estimator.state_dict())
        m = m.cuda()
        m.refresh(torch.device("cuda", 0))
        base = None
        for setting in settings:
            opts = dict(kv.split("=") for kv in setting.split(",") if kv)
            for k, v in opts.items():
                m.set_option(k, int(v))
            e = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), torch.tensor(0.4)).cpu()
            s = m.solve(z.cuda(), ts, mu.cuda(), mask.cuda(), lengths=lengths).cpu()
            torch.cuda.synchronize()
            if base is None:
                base = (e, s)
            fin = bool(torch.isfinite(e).all() and torch.isfinite(s).all())
            good = fin and rel(e, ref_e) < tol and rel(s, ref_s) < tol
            ok = ok and good
            print(f"[{name}] [{setting or 'defaults'}] estimator vs oracle {rel(e, ref_e):.3e} vs default {rel(e, base[0]):.3e} | "
                  f"solve vs oracle {rel(s, ref_s):.3e} vs default {rel(s, base[1]):.3e} finite={fin} {'ok' if good else 'FAIL'}", flush=True)
            for k in opts:  # back to the defaults known to ab_timeline
                m.set_option(k, {"small_tiles": 1024, "attn_persist": 1, "bf16_mid": 1, "rowln_ff2": 1}.get(k, 0))
        m.close()
    print("ALL OK" if ok else "SOME FAILED")


if __name__ == "__main__":
    main()
