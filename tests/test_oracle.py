"""CPU tests: the oracle against the golden vectors produced by the reference's own code
(tests/golden/make_golden.py), plus the behavioural facts SURVEY.md section 8(c) lists."""
import pytest
import torch

from conftest import GOLDEN_CASES, cfm_params, load_golden, rel_l2
from oracle import cfm_oracle as O
import matcha_tts_24k_b200 as P
from matcha_tts_24k_b200 import synthetic as syn


def build_oracle(case, dtype=torch.float32):
    m = O.CFM(200, 100, cfm_params(case["solver"]), case["decoder_params"]).eval()
    syn.fill_named_seed(m.estimator, case["weight_seed"])
    return m.to(dtype)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    case = load_golden(name)
    m = build_oracle(case)
    assert sorted(m.estimator.state_dict().keys()) == [str(k) for k in case["state_dict_keys"]]
    assert sum(p.numel() for p in m.estimator.parameters()) == case["n_params"]
    mu, mask, z, _ = syn.make_inputs(case["lengths"], seed=case["input_seed"], T=case["T"])
    v = m.estimator(z, mask, mu, torch.tensor(0.3))
    assert rel_l2(v, case["estimator_v_t03"]) < 2e-6
    out = m.solve(z, torch.linspace(0, 1, case["n_steps"] + 1), mu, mask)
    assert rel_l2(out, case["solve_out"]) < 2e-6
    assert rel_l2(m(mu, mask, case["n_steps"]), case["forward_out"]) < 2e-6


@pytest.mark.parametrize("name", ["prod_euler", "default_euler", "tiny_euler"])
def test_product_weight_spec_equals_reference_keys(name):
    case = load_golden(name)
    cfg = P.config_from_decoder_params(200, 100, **case["decoder_params"])
    spec = P.weight_spec(cfg)
    assert sorted(n for n, _ in spec) == [str(k) for k in case["state_dict_keys"]]
    oracle_sd = build_oracle(case).estimator.state_dict()
    for n, shape in spec:
        assert tuple(oracle_sd[n].shape) == tuple(shape), n
    w = P.EstimatorWeights(cfg)
    missing, unexpected = w.load_state_dict(oracle_sd, strict=True)
    assert not missing and not unexpected


def test_param_counts_match_survey():
    for dec, n in ((syn.PROD, 37_034_212), (syn.DEFAULT, 25_823_780)):
        cfg = P.config_from_decoder_params(200, 100, **dec)
        assert sum(torch.Size(s).numel() for _, s in P.weight_spec(cfg)) == n


def test_solver_nfe_counts():
    calls = []
    f = lambda t, y: (calls.append(float(t)), -y)[1]
    grid = torch.linspace(0, 1, 5)
    for method, nfe in O.SOLVER_NFE.items():
        calls.clear()
        O.odeint_fixed_grid(f, torch.ones(3), grid, method)
        assert len(calls) == 4 * nfe
    with pytest.raises(ValueError):
        O.odeint_fixed_grid(f, torch.ones(3), grid, "dopri5")


def test_solvers_integrate_linear_ode():
    grid = torch.linspace(0, 1, 21, dtype=torch.float64)
    y0 = torch.ones(2, dtype=torch.float64)
    exact = torch.exp(torch.tensor(-1.0, dtype=torch.float64))
    err = {m: float((O.odeint_fixed_grid(lambda t, y: -y, y0, grid, m)[0] - exact).abs()) for m in O.SOLVER_NFE}
    assert err["euler"] > err["midpoint"] > err["heun3"] > err["rk4"]


def test_attention_mask_is_additive_and_padding_leaks():
    """padded keys are attended (+0 vs +1) and GroupNorm spans padded T,
    so the valid region depends on the amount of padding."""
    case = load_golden("tiny_euler")
    m = build_oracle(case, torch.float64)
    mu, mask, z, _ = syn.make_inputs([20], seed=3, T=20)
    mu, mask, z = mu.double(), mask.double(), z.double()
    pad = lambda x: torch.nn.functional.pad(x, (0, 12))
    t = torch.tensor(0.5, dtype=torch.float64)
    v0 = m.estimator(z, mask, mu, t)
    v1 = m.estimator(pad(z), pad(mask), pad(mu), t)
    assert (v1[:, :, 20:] == 0).all()
    assert rel_l2(v1[:, :, :20], v0) > 1e-3


def test_padded_state_keeps_noise():
    """The velocity is masked, so padded frames of the solve output equal the injected z."""
    case = load_golden("tiny_rk4_padded")
    mu, mask, z, lengths = syn.make_inputs(case["lengths"], seed=case["input_seed"], T=case["T"])
    out = torch.as_tensor(case["solve_out"])
    for b, L in enumerate(case["lengths"]):
        assert torch.equal(out[b, :, L:], z[b, :, L:])
