"""The pad-aware packed schedule (tests/packed_model.py == what csrc/ implements) reproduces the
dense padded reference semantics exactly (fp64), including GroupNorm-over-padding and the
additive attention mask."""
import pytest
import torch

from conftest import cfm_params, rel_l2
from oracle import cfm_oracle as O
from matcha_tts_24k_b200 import synthetic as syn
from packed_model import PackedEstimator

DEC = dict(channels=(64, 64), dropout=0.05, attention_head_dim=32, n_blocks=2, num_mid_blocks=2, num_heads=2)


@pytest.mark.parametrize("lengths,T", [([20], 20), ([21], 22), ([20], 22), ([17, 40, 33], 40), ([9, 30], 64),
                                        ([1, 2, 3], 4), ([5], 50)])
def test_packed_equals_padded_dense(lengths, T):
    m = O.CFM(200, 100, cfm_params(), DEC).eval()
    syn.fill_named_seed(m.estimator, 99)
    m = m.double()
    mu, mask, z, _ = syn.make_inputs(lengths, seed=5, T=T)
    mu, mask, z = mu.double(), mask.double(), z.double()
    t = torch.tensor(0.37, dtype=torch.float64)
    with torch.inference_mode():
        v_ref = m.estimator(z, mask, mu, t)
        pk = PackedEstimator(m.estimator.state_dict(), 2, 32, 2, 2)
        emb = O.sinusoidal_embedding(t, 200).double()[0]
        for b, L in enumerate(lengths):
            v = pk(z[b, :, :L].T.contiguous(), mu[b, :, :L].T.contiguous(), emb, L, T)
            assert rel_l2(v.T, v_ref[b, :, :L]) < 1e-11, (b, L, T)
