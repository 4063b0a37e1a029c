"""GPU tests of the round-2 host-facing features, all through the C ABI and against the CPU oracle:

  * estimator call with one time value per utterance, t of shape (B,)  (reference flow_matching.py:84-97)
  * compute_loss forward value                                         (reference flow_matching.py:65-107)
  * plan cache: alternating request shapes reuse their plans / graphs  (reference server.py:93-119 request loop)
  * calls on different streams are ordered by the library             (one workspace per plan)
  * host-buffer entry: fresh result tensors, speaker vectors on the host path
  * the per-launch timeline used by bench.py's roofline.kernel
"""
import pytest
import torch

import matcha_tts_24k_b200 as P
from matcha_tts_24k_b200 import synthetic as syn
from conftest import cfm_params, rel_l2
from oracle import cfm_oracle as O

pytestmark = pytest.mark.gpu

SMALL = dict(channels=(128, 128), dropout=0.05, attention_head_dim=64, n_blocks=1, num_mid_blocks=1, num_heads=2)


def pair(dec, solver="euler", precision="fp32", in_channels=200, seed=1234):
    ora = O.CFM(in_channels, 100, cfm_params(solver), dec).eval()
    syn.fill_named_seed(ora.estimator, seed)
    m = P.CFM(in_channels, 100, cfm_params(solver), dec, precision=precision).eval()
    m.estimator.load_state_dict(ora.estimator.state_dict())
    return ora, m.cuda()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
def test_estimator_with_per_utterance_time(precision, tol):
    lengths = [70, 41, 64, 9]
    ora, m = pair(syn.PROD, precision=precision)
    mu, mask, z, _ = syn.make_inputs(lengths, seed=7)
    t = torch.tensor([0.05, 0.5, 0.93, 0.31])
    with torch.inference_mode():
        ref = ora.estimator(z, mask, mu, t)
    v = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t.cuda())
    err = rel_l2(v.cpu(), ref)
    print(f"per-utterance t [{precision}]: rel_l2={err:.3e} max_abs={float((v.cpu() - ref).abs().max()):.3e}")
    assert err <= tol
    # each utterance equals a scalar-t call with its own time (the time embedding is the only thing that differs)
    for b in (0, 2):
        vb = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t[b])
        if precision == "fp32":
            assert rel_l2(vb[b].cpu(), v[b].cpu()) <= 1e-6
    # the planned time grid survives an estimator call: a solve afterwards still matches the oracle
    m.close()


def test_compute_loss_forward_value_matches_oracle():
    lengths = [50, 33]
    ora, m = pair(SMALL, precision="fp32")
    mu, mask, x1, _ = syn.make_inputs(lengths, seed=11)
    g = torch.Generator().manual_seed(5)
    t = torch.rand(2, generator=g)
    x0 = mu + torch.randn(mu.shape, generator=g)
    # oracle value with the same draws (restating compute_loss with t and x0 injected)
    tt = t.reshape(2, 1, 1)
    y = (1 - (1 - ora.sigma_min) * tt) * x0 + tt * x1
    u = x1 - (1 - ora.sigma_min) * x0
    with torch.inference_mode():
        pred = ora.estimator(y, mask, mu, t)
    ref = torch.nn.functional.mse_loss(pred * mask, u * mask, reduction="sum") / (mask.sum() * u.shape[1])
    loss = m.compute_loss(x1.cuda(), mask.cuda(), mu.cuda(), t=t.cuda(), x0=x0.cuda())
    print(f"compute_loss: ours={float(loss):.6f} oracle={float(ref):.6f}")
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    # with its own random draws it runs and is finite (the draws come from torch's CUDA generator, as in the reference)
    assert torch.isfinite(m.compute_loss(x1.cuda(), mask.cuda(), mu.cuda()))
    m.close()


def test_plan_cache_alternating_shapes():
    ora, m = pair(SMALL, precision="fp32")
    cases = []
    for i, lengths in enumerate(([40], [57], [40, 33])):
        mu, mask, z, _ = syn.make_inputs(lengths, seed=20 + i)
        ts = torch.linspace(0, 1, 4)
        cases.append((lengths, mu.cuda(), mask.cuda(), z.cuda(), ts, ora.solve(z, ts, mu, mask)))
    first = {}
    for rnd in range(4):  # A B C A B C ...: from the second round on every plan is cached and its graph replays
        for i, (lengths, mu, mask, z, ts, ref) in enumerate(cases):
            out = m.solve(z, ts, mu, mask)
            assert rel_l2(out.cpu(), ref) <= 2e-5
            if rnd == 1:
                first[i] = out.clone()
            if rnd > 1:
                assert torch.equal(out, first[i])  # graph replay of the cached plan: bitwise repeatable
    m.set_option("plan_cache", 1)  # a one-entry cache evicts on every change of shape and still decodes correctly
    for lengths, mu, mask, z, ts, ref in cases + cases:
        assert rel_l2(m.solve(z, ts, mu, mask).cpu(), ref) <= 2e-5
    m.close()


def test_calls_on_different_streams_are_ordered():
    ora, m = pair(SMALL, precision="fp32")
    lengths = [64, 50]
    mu, mask, z, _ = syn.make_inputs(lengths, seed=31)
    mu2, _, z2, _ = syn.make_inputs(lengths, seed=32)
    ts = torch.linspace(0, 1, 3)
    ref1, ref2 = ora.solve(z, ts, mu, mask), ora.solve(z2, ts, mu2, mask)
    mu_d, z_d, mask_d = mu.cuda(), z.cuda(), mask.cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    for _ in range(3):
        with torch.cuda.stream(side):
            out1 = m.solve(z_d, ts, mu_d, mask_d)          # asynchronous, on a side stream
        out2 = m.solve_host(z2, ts, mu2, lengths)            # immediately after, on the library's own stream
        side.synchronize()
        assert rel_l2(out1.cpu(), ref1) <= 2e-5 and rel_l2(out2, ref2) <= 2e-5
    m.close()


def test_solve_host_returns_fresh_tensors_and_takes_speakers():
    S = 16
    ora, m = pair(SMALL, precision="fp32", in_channels=200 + S)
    lengths = [30, 22]
    mu, mask, z, _ = syn.make_inputs(lengths, seed=41)
    spks = torch.randn(2, S, generator=torch.Generator().manual_seed(1))
    ts = torch.linspace(0, 1, 3)
    ref = ora.solve(z, ts, mu, mask, spks)
    a = m.solve_host(z, ts, mu, lengths, spks=spks)
    b = m.solve_host(z, ts, mu, lengths, spks=spks)
    assert a.data_ptr() != b.data_ptr() and torch.equal(a, b)
    assert rel_l2(a, ref) <= 2e-5
    buf = torch.empty_like(mu).pin_memory()
    assert m.solve_host(z, ts, mu, lengths, spks=spks, out=buf) is buf and torch.equal(buf, a)
    with pytest.raises(ValueError, match="speaker"):
        m.solve_host(z, ts, mu, lengths)  # an estimator with speaker channels never reuses a stale device pointer
    with pytest.raises(ValueError):
        m.solve_host(z, ts, mu[:, :80], lengths, spks=spks)
    m.close()


def test_timeline_lists_every_launch():
    _, m = pair(SMALL, precision="bf16")
    lengths = [200, 150]
    mu, mask, z, _ = syn.make_inputs(lengths, seed=51, device="cuda")
    ts = torch.linspace(0, 1, 3)
    m.solve(z, ts, mu, mask)
    rows = m.timeline(z, ts, mu, lengths)
    tags = {r[0] for r in rows}
    assert {"qkv", "attention", "out_proj", "ff1_snake", "ff2_copy", "res_conv", "final_proj_ode", "conv_s2", "conv_transpose"} <= tags
    assert all(r[5] >= 0 for r in rows) and sum(r[4] for r in rows) > 0
    m.close()


@pytest.mark.parametrize("dec,lengths", [(syn.PROD, [300, 171, 64]), (syn.DEFAULT, [257, 130]), (SMALL, [150, 97])])
def test_fused_feed_forward_kernel_matches_two_gemm_schedule(dec, lengths):
    """ff_fused.cuh (FF1 -> SnakeBeta -> FF2 in one kernel, hidden activation in tensor memory, A operand of the second GEMM read
    from TMEM) against the default two-GEMM schedule and the oracle: solve x2, bf16 tolerance; "small_tiles" 0 sends every
    resolution through the fused kernel."""
    ora, m = pair(dec, precision="bf16")
    mu, mask, z, _ = syn.make_inputs(lengths, seed=61)
    ts = torch.linspace(0, 1, 3)
    ref = ora.solve(z, ts, mu, mask)
    outs = {}
    m.refresh(torch.device("cuda", 0))
    for fused in (0, 1):
        m.set_option("small_tiles", 0)
        m.set_option("ff_fused", fused)
        outs[fused] = m.solve(z.cuda(), ts, mu.cuda(), mask.cuda()).cpu()
        err = rel_l2(outs[fused], ref)
        print(f"ff_fused={fused}: rel_l2 vs oracle {err:.3e}")
        assert err <= 1e-2
    assert rel_l2(outs[1], outs[0]) <= 1e-2
    m.close()


@pytest.mark.parametrize("dec,lengths", [(syn.PROD, [300, 171, 64, 1]), (SMALL, [150, 97])])
def test_fused_linear_layernorm_kernel_matches_unfused_schedule(dec, lengths):
    """rowln.cuh (out-proj / FF2 + residual add + the following LayerNorm in one kernel, whole rows per CTA) against the default
    schedule and the oracle: solve x2, bf16 tolerance; "rowln" 2 sends every resolution through the fused kernel."""
    ora, m = pair(dec, precision="bf16")
    mu, mask, z, _ = syn.make_inputs(lengths, seed=62)
    ts = torch.linspace(0, 1, 3)
    ref = ora.solve(z, ts, mu, mask)
    outs = {}
    m.refresh(torch.device("cuda", 0))
    for fused in (0, 2):
        m.set_option("rowln", fused)
        outs[fused] = m.solve(z.cuda(), ts, mu.cuda(), mask.cuda()).cpu()
        err = rel_l2(outs[fused], ref)
        print(f"rowln={fused}: rel_l2 vs oracle {err:.3e}")
        assert err <= 1e-2
    assert rel_l2(outs[2], outs[0]) <= 1e-2
    m.close()


@pytest.mark.parametrize("dec,lengths,in_ch", [(syn.DEFAULT, [257, 130, 3], 200), (syn.PROD, [150, 97], 216), (SMALL, [150, 97, 64], 200)])
def test_fp32_on_the_tensor_pipe_matches_oracle_and_fp32_reference(dec, lengths, in_ch):
    """precision "fp32_tc" (fp32 storage, bf16 x 3 split operands through the bf16 GEMM kernels and the split-operand attention
    kernel): <= 1e-3 of the oracle (north_star's fp32 tolerance; measured ~1e-5) and of the library's fp32-FMA mode, for the default
    C = 320 estimator, speaker channels (S = 16) and a 128-wide estimator; solve, per-utterance-t estimator call, option toggling."""
    ora, m = pair(dec, precision="fp32_tc", in_channels=in_ch)
    S = in_ch - 200
    mu, mask, z, _ = syn.make_inputs(lengths, seed=63)
    spks = torch.randn(len(lengths), S, generator=torch.Generator().manual_seed(2)) if S else None
    ts = torch.linspace(0, 1, 4)
    ref = ora.solve(z, ts, mu, mask, spks) if S else ora.solve(z, ts, mu, mask)
    kw = {"spks": spks.cuda()} if S else {}
    out = m.solve(z.cuda(), ts, mu.cuda(), mask.cuda(), **kw).cpu()
    err = rel_l2(out, ref)
    print(f"fp32_tc solve: rel_l2 vs oracle {err:.3e} max_abs {float((out - ref).abs().max()):.3e}")
    assert torch.isfinite(out).all() and err <= 1e-3
    for b, L in enumerate(lengths):
        assert torch.equal(out[b, :, L:], z[b, :, L:])
    if not S:
        t = torch.rand(len(lengths), generator=torch.Generator().manual_seed(4))
        v = m.estimator(z.cuda(), mask.cuda(), mu.cuda(), t).cpu()
        with torch.inference_mode():
            v_ref = ora.estimator(z, mask, mu, t)
        assert rel_l2(v, v_ref) <= 1e-3
        m.set_option("fp32_tc", 0)  # the same handle as the fp32-FMA reference mode
        exact = m.solve(z.cuda(), ts, mu.cuda(), mask.cuda()).cpu()
        assert rel_l2(exact, ref) <= 2e-5 and rel_l2(out, exact) <= 1e-3
    m.close()


def test_solve_into_a_caller_owned_result_tensor():
    """CFM.solve(..., out=): the result lands in the caller's tensor (no allocation per call); wrong shapes / devices raise."""
    ora, m = pair(SMALL, precision="fp32")
    mu, mask, z, _ = syn.make_inputs([90, 41], seed=64)
    ts = torch.linspace(0, 1, 3)
    ref = ora.solve(z, ts, mu, mask)
    out = torch.full_like(mu, float("nan")).cuda()
    res = m.solve(z.cuda(), ts, mu.cuda(), mask.cuda(), out=out)
    assert res is out and rel_l2(out.cpu(), ref) <= 2e-5
    with pytest.raises(ValueError):
        m.solve(z.cuda(), ts, mu.cuda(), mask.cuda(), out=torch.empty(1, 100, 8, device="cuda"))
    with pytest.raises(ValueError):
        m.solve(z.cuda(), ts, mu.cuda(), mask.cuda(), out=torch.empty_like(mu))
    m.close()


def _sharded_case(devices):
    lengths = [120, 64, 97, 33, 150, 88, 140]
    ora, single = pair(SMALL, precision="fp32")
    mu, mask, z, _ = syn.make_inputs(lengths, seed=71)
    ts = torch.linspace(0, 1, 3)
    ref = ora.solve(z, ts, mu, mask)
    one = single.solve(z.cuda(), ts, mu.cuda(), mask.cuda()).cpu()
    sh = P.ShardedCFM(200, 100, cfm_params("euler"), SMALL, devices=devices, precision="fp32")
    sh.load_estimator_state_dict(ora.estimator.state_dict())
    out = torch.full_like(mu, float("nan")).pin_memory()
    res = sh.solve_host(z.pin_memory(), ts, mu.pin_memory(), lengths, out=out)
    assert res is out and torch.isfinite(out).all()
    assert sorted(i for s in sh.last_shards for i in s) == list(range(len(lengths))) and all(sh.last_shards)
    assert rel_l2(out, ref) <= 2e-5
    # utterances do not interact given T: the sharded decode equals the one-GPU decode of the whole batch (fp32 mode: the
    # stand-alone statistics pass sums in fp64, order-independent to ~1e-16)
    assert rel_l2(out, one) <= 1e-6
    again = sh.solve_host(z, ts, mu, lengths)  # pageable inputs, fresh result tensor, cached plans / graphs
    assert torch.equal(again, out)
    sh.close()
    single.close()


def test_sharded_decode_two_replicas_on_one_gpu():
    """The in-process multi-GPU path (ShardedCFM -> cfm_solve_host_indexed / cfm_synchronize) with two handles on device 0."""
    _sharded_case([0, 0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_decode_two_gpus():
    _sharded_case([0, 1])


def test_front_and_back_of_the_decode_match_reference_and_oracle():
    """cfm_front_durations / cfm_front_expand / cfm_denormalize against the golden vectors made with the reference's own
    functions (inference.py:146-172) and, on random ragged inputs incl. an all-masked row and odd / even fine lengths, the oracle."""
    import numpy as np
    import os
    from oracle import front_back as FB
    _, m = pair(SMALL, precision="fp32")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "front_back.npz"))
    mu_x, dur = torch.from_numpy(g["mu_x"]), torch.from_numpy(g["phoneme_durations"])
    mu_y, y_mask, y_lengths = m.expand_encoder_output(mu_x.cuda(), dur.cuda())
    assert y_lengths == g["y_lengths"].tolist() and tuple(mu_y.shape) == g["mu_y"].shape
    assert torch.equal(y_mask.cpu(), torch.from_numpy(g["y_mask"]))
    assert torch.allclose(mu_y.cpu(), torch.from_numpy(g["mu_y"]), atol=1e-6, rtol=0)
    mel = m.denormalize(torch.from_numpy(g["decoder_outputs"]).cuda(), float(g["mel_mean"]), float(g["mel_std"]), int(g["y_max_length"]))
    assert torch.allclose(mel.cpu(), torch.from_numpy(g["mel"]), atol=1e-6, rtol=0)
    gen = torch.Generator().manual_seed(3)
    for B, Tx in ((5, 61), (2, 300), (1, 7)):
        x_lengths = torch.randint(1, Tx + 1, (B,), generator=gen)
        x_lengths[0] = Tx
        x_mask = FB.sequence_mask(x_lengths, Tx).unsqueeze(1).float()
        mu_x = torch.randn(B, 100, Tx, generator=gen) * x_mask
        dur = (torch.rand(B, Tx, generator=gen) * 9).round().clamp(min=1) * x_mask.squeeze(1)
        ref_mu, ref_mask, ref_len, _ = FB.front(mu_x, dur, x_mask)
        mu_y, y_mask, y_lengths = m.expand_encoder_output(mu_x.cuda(), dur.cuda())
        assert y_lengths == ref_len.tolist() and torch.equal(y_mask.cpu(), ref_mask)
        assert torch.allclose(mu_y.cpu(), ref_mu, atol=1e-6, rtol=0)
    m.close()


def test_synthesise_mel_equals_front_decode_back_of_the_oracle():
    from oracle import front_back as FB
    ora, m = pair(SMALL, precision="fp32")
    gen = torch.Generator().manual_seed(9)
    B, Tx = 2, 40
    x_lengths = torch.tensor([40, 28])
    x_mask = FB.sequence_mask(x_lengths, Tx).unsqueeze(1).float()
    mu_x = torch.randn(B, 100, Tx, generator=gen) * x_mask
    dur = (torch.rand(B, Tx, generator=gen) * 5).round().clamp(min=1) * x_mask.squeeze(1)
    mu_y, y_mask, y_lengths, y_max = FB.front(mu_x, dur, x_mask)
    # the decode draws its seed-42 noise on the GPU (reference flow_matching.py:43-55): replay those draws for the oracle
    gdev = torch.Generator(device="cuda")
    gdev.manual_seed(42)
    z = mu_y + torch.randn(mu_y.shape, generator=gdev, device="cuda").cpu()
    ref = FB.back(ora.solve(z, torch.linspace(0, 1, 4), mu_y, y_mask), y_max, -5.5, 2.0)
    mel, lens = m.synthesise_mel(mu_x.cuda(), dur.cuda(), 3, mel_mean=-5.5, mel_std=2.0)
    assert lens == y_lengths.tolist() and tuple(mel.shape) == tuple(ref.shape)
    assert rel_l2(mel.cpu(), ref) <= 2e-5
    m.close()
