"""CPU tests of the host logic and of the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import types

import pytest
import torch

import matcha_tts_24k_b200 as P
from matcha_tts_24k_b200 import _native as N
from conftest import ROOT, cfm_params


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfm_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = N.load_library(build_if_missing=True)
    names = declared_symbols()
    assert "cfm_solve" in names and "cfm_create" in names
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/cfm_b200.h but not exported"
    assert sorted(N.EXPORTS) == names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = N.load_library(build_if_missing=True)
    cfg = N.Config(200, 100, 384, 6, 64, 2, 2, 0, 0, 0)
    h = ctypes.c_void_p()
    rc = lib.cfm_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in lib.cfm_last_error(None)


def test_create_rejects_bad_config():
    lib = N.load_library(build_if_missing=True)
    h = ctypes.c_void_p()
    for bad in (N.Config(200, 100, 100, 6, 64, 2, 2, 0, 0, 0), N.Config(200, 100, 384, 6, 48, 2, 2, 0, 0, 0),
                N.Config(200, 100, 384, 6, 64, 2, 2, 7, 0, 0)):
        assert lib.cfm_create(ctypes.byref(bad), ctypes.byref(h)) == -1
        assert lib.cfm_last_error(None)


def test_cfm_surface_matches_reference():
    m = P.CFM(in_channels=200, out_channel=100, cfm_params=cfm_params("midpoint"), decoder_params=P.synthetic.PROD)
    assert m.solver == "midpoint" and m.n_feats == 200 and m.sigma_min == 1e-4 and m.use_mu_prior is True
    keys = list(m.state_dict().keys())
    assert all(k.startswith("estimator.") for k in keys) and len(keys) == 270
    assert "estimator.down_blocks.0.1.0.ff._orig_mod.net.0.alpha" in keys
    m.solver = "euler"  # reference cli.py:94 / server.py:109 assign it at run time
    # reference server.py:47 re-assigns .estimator to a torch.compile wrapper; weights must stay reachable
    before = {k: v.data_ptr() for k, v in m._weights.state_dict().items()}
    m.estimator = torch.compile(m.estimator, dynamic=True)
    assert {k: v.data_ptr() for k, v in m._weights.state_dict().items()} == before
    with pytest.raises(ValueError):
        P.CFM(200, 100, cfm_params(), dict(P.synthetic.PROD, down_block_type="conformer"))


def test_argument_checks_precede_the_library_call():
    """The library packs with its own B / F / T; mismatching tensors must raise on the host (reference: conv shape error)."""
    m = P.CFM(200, 100, cfm_params(), P.synthetic.PROD).eval()
    mu80, mu100 = torch.zeros(2, 80, 10), torch.zeros(2, 100, 10)
    with pytest.raises(ValueError, match="80 mel channels"):
        m._check_shapes(mu80, mu80)
    with pytest.raises(ValueError, match="disagree"):
        m._check_shapes(mu100, torch.zeros(2, 100, 12))
    assert m._check_shapes(mu100, mu100) == (2, 100, 10)
    with pytest.raises(ValueError, match="1 entries"):
        m._check_lengths([5], 2, 10)
    with pytest.raises(ValueError):
        m._check_lengths([5, 11], 2, 10)
    with pytest.raises(ValueError, match="speaker channels"):
        m._check_spks(torch.zeros(2, 96), 2, prep=False)
    ms = P.CFM(296, 100, cfm_params(), P.synthetic.PROD).eval()
    with pytest.raises(ValueError, match="missing"):
        ms._check_spks(None, 2, prep=False)
    with pytest.raises(ValueError, match="shape"):
        ms._check_spks(torch.zeros(2, 95), 2, prep=False)


def test_cpu_tensors_are_refused():
    m = P.CFM(200, 100, cfm_params(), P.synthetic.PROD).eval()
    mu, mask, z, _ = P.synthetic.make_inputs([10, 7], T=10)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(mu, mask, 2)


def test_parent_state_dict_loading_like_inference_py():
    """reference matcha/inference.py:52,193-194: the CFM sits at `.decoder` and is filled by
    load_state_dict(strict=False); every decoder.* key must be consumed."""
    class Parent(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.decoder = P.CFM(200, 100, cfm_params(), P.synthetic.DEFAULT)

    a, b = Parent(), Parent()
    P.synthetic.fill_named_seed(a.decoder.estimator, 5)
    res = b.load_state_dict(a.state_dict(), strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(a.decoder.estimator.final_proj.weight, b.decoder.estimator.final_proj.weight)


def test_lengths_from_mask():
    lengths = torch.tensor([5, 8, 1])
    mask = P.synthetic.sequence_mask(lengths, 8).unsqueeze(1).float()
    assert P.lengths_from_mask(mask) == [5, 8, 1]
    bad = mask.clone()
    bad[0, 0, 6] = 1.0
    with pytest.raises(ValueError):
        P.lengths_from_mask(bad)
    with pytest.raises(ValueError):
        P.lengths_from_mask(mask.bool())
    with pytest.raises(ValueError):
        P.lengths_from_mask(mask * 0.5)
    # a half-precision mask of cfg4's length: summing it in its own dtype would round (bf16 has 8 bits of mantissa)
    long = P.synthetic.sequence_mask(torch.tensor([2812, 300]), 2812).unsqueeze(1)
    assert P.lengths_from_mask(long.to(torch.bfloat16)) == [2812, 300]
    assert P.lengths_from_mask(long.to(torch.float16)) == [2812, 300]


def test_sharding_is_balanced_and_complete():
    lengths = P.synthetic.config_lengths("cfg3")
    assert len(lengths) == 256 and min(lengths) >= 188 and max(lengths) <= 1125
    for world in (1, 2, 4, 8):
        shards = P.shard_utterances(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(256))
        loads = [sum(P.sharding.utterance_cost(lengths[i], 384) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.02
    outs = [[f"o{i}" for i in s] for s in P.shard_utterances(lengths, 4)]
    assert P.gather_outputs(outs, P.shard_utterances(lengths, 4), 256) == [f"o{i}" for i in range(256)]


def test_flop_model_matches_survey():
    assert abs(P.synthetic.algorithmic_flops([938] * 32, 384, 10) / 14.93e12 - 1) < 2e-3
    assert abs(P.synthetic.algorithmic_flops([150], 384, 10) / 63.7e9 - 1) < 5e-3


def test_precision_names_match_the_header():
    """cfm_precision in include/cfm_b200.h <-> the names the host module accepts (bf16 / fp32 / fp32_tc)."""
    import re
    from matcha_tts_24k_b200 import _native as N
    src = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()
    enum = re.search(r"typedef enum \{([^}]*)\} cfm_precision;", src).group(1)
    vals = {k.strip().lower().replace("cfm_prec_", ""): int(v) for k, v in (item.split("=") for item in enum.split(","))}
    assert vals == N.PREC
    with pytest.raises(ValueError):
        P.CFM(200, 100, cfm_params("euler"), P.synthetic.PROD, precision="fp16")
