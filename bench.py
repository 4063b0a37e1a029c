"""Benchmark of the CFM decode hot path (BASELINE.json metric: mel-frames/s at 10 Euler steps).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5|cfg1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full decode (pack -> CUDA graph of the 10-step Euler loop -> unpack) of one synthetic batch.
  value     whole-job mel-frames/s with mu / z / mask resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the C-ABI host-buffer call cfm_solve_host: pinned host mu, z -> H2D -> decode -> D2H mel
  roofline  algorithmic FLOPs of the decode / device time of the graph launch, against the measured
            bf16 tensor peak of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restated reference PyTorch path) on the box's host cores, bounded sample, rank 0 only
`--impl reference` times that CPU path alone (the reference has no importable CPU/GPU build here: DESIGN.md).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "cfm_decode_mel_frames_per_s_10_euler_steps"
UNIT = "mel-frames/s"
N_STEPS_ODE = 10
FRAME_SECONDS = 256.0 / 24000.0  # hop / sample rate (reference configs/data/corpus-24k.yaml:20-22)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"], "hbm_gbs": p["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.rows.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def mark(self):
        """Index of the next sample: call at the start and end of the timed region."""
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.rows[lo:hi]:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None, "power_w_max": max(pw) if pw else None}


def workload(name: str):
    import matcha_tts_24k_b200 as P
    lengths = P.synthetic.config_lengths(name)
    desc = {"cfg1": "cfg1: B=1 L=150", "cfg2": "cfg2: B=32 x ~10 s (L=T=938)", "cfg3": "cfg3: B=256 mixed 2-12 s, mask-packed",
            "cfg4": "cfg4: B=16 x 30 s (L=T=2812)", "cfg5": "cfg5: B=64 x ~10 s"}[name]
    return lengths, desc


def cpu_oracle_throughput(lengths, n_timed: int, threads: int):
    """The reference's PyTorch path restated (oracle/cfm_oracle.py) on the host cores: fp32, inference_mode."""
    import matcha_tts_24k_b200 as P
    from oracle import cfm_oracle as O
    torch.set_num_threads(threads)
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    ora = O.CFM(200, 100, cp, P.synthetic.PROD).eval()
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1)
    ts = torch.linspace(0, 1, N_STEPS_ODE + 1)
    ora.solve(z[:1, :, :64].contiguous(), ts[:2], mu[:1, :, :64].contiguous(), mask[:1, :, :64].contiguous())  # warm the allocator
    times = []
    for _ in range(n_timed):
        t0 = time.perf_counter()
        ora.solve(z, ts, mu, mask)
        times.append(time.perf_counter() - t0)
    return sum(lengths) / statistics.median(times), statistics.median(times)


def workload_config(desc, args, T):
    """`config` of the JSON line: names the workload only, identical for both arms."""
    return {"workload": desc, "estimator": "prod C=384 H=6 d=64 n_blocks=2 mid=2 (37.03 M params, random init + N(0,0.1) 1-D)",
            "ode": f"{args.solver} x{args.ode_steps}", "spks": args.spks, "t_pad": T,
            "l2": "activations of one decode exceed the 126 MB L2 (cfg2 workspace 667 MiB): no flush between steps"}


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of the path = the oracle restatement (the
    reference is not importable here), all host threads, one STEP = one full decode of a bounded SAMPLE of the workload
    (2 utterances); exactly --warmup untimed and --steps timed steps."""
    if rank != 0:
        return
    import matcha_tts_24k_b200 as P
    from oracle import cfm_oracle as O
    threads = len(os.sched_getaffinity(0))
    lengths, desc = workload(args.workload)
    T = 2 * ((max(lengths) + 1) // 2)
    sample = lengths[:2] if len(lengths) > 2 else lengths  # bounded sample of the same workload
    torch.set_num_threads(threads)
    cp = types.SimpleNamespace(solver=args.solver, sigma_min=1e-4, use_mu_prior=True)
    ora = O.CFM(200 + args.spks, 100, cp, P.synthetic.PROD).eval()
    mu, mask, z, _ = P.synthetic.make_inputs(sample, seed=1, T=T if args.workload != "cfg3" else None)
    ts = torch.linspace(0, 1, args.ode_steps + 1)
    spks = torch.randn(len(sample), args.spks, generator=torch.Generator().manual_seed(3)) if args.spks else None
    step = (lambda: ora.solve(z, ts, mu, mask, spks)) if args.spks else (lambda: ora.solve(z, ts, mu, mask))
    for _ in range(args.warmup):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    fps = sum(sample) / sec
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(desc, args, T),
            "note": "reference CPU path = oracle restatement of the reference PyTorch CFM (the reference itself is not importable "
                    "here: torchdiffeq / diffusers absent, and it hard-codes CUDA); value = frames of the sample / mean step time",
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"each step = full {args.ode_steps}-step {args.solver} decode of {len(sample)} utterance(s) "
                                       f"({sum(sample)} frames) of {desc}; {args.steps} timed steps, mean"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rtf": sec / (sum(sample) * FRAME_SECONDS)}
    print(json.dumps(line), flush=True)


def gpu_pytorch_baseline(dev, names=("cfg2", "cfg1"), n_timed=5, with_compile=True):
    """The baseline SURVEY.md section 2a names: the reference's PyTorch path (oracle restatement) on the SAME B200 - eager fp32,
    eager under torch.autocast(bf16), and with the estimator wrapped in torch.compile(dynamic=True) as reference server.py:47
    does.  Full 10-step Euler decode, CUDA-event timed; runs after (never inside) the repo arm's timed regions."""
    import matcha_tts_24k_b200 as P
    from oracle import cfm_oracle as O
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    ora = O.CFM(200, 100, cp, P.synthetic.PROD).eval().to(dev)
    out = {}

    def time_it(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    for name in names:
        lengths, desc = workload(name)
        mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device=dev)
        ts = torch.linspace(0, 1, N_STEPS_ODE + 1, device=dev)
        rec = {"workload": desc, "frames": sum(lengths)}
        try:
            ms = time_it(lambda: ora.solve(z, ts, mu, mask), n_timed)
            rec["eager_fp32"] = {"ms_per_decode": ms, "value": sum(lengths) / ms * 1e3, "unit": UNIT}

            def amp():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return ora.solve(z, ts, mu, mask)
            ms = time_it(amp, n_timed)
            rec["eager_autocast_bf16"] = {"ms_per_decode": ms, "value": sum(lengths) / ms * 1e3, "unit": UNIT}
        except Exception as e:  # noqa: BLE001 - a baseline must never take the bench down
            rec["error"] = f"{type(e).__name__}: {e}"[:300]
        out[name] = rec
    if with_compile:
        try:
            t0 = time.perf_counter()
            eager = ora.estimator
            ora.estimator = torch.compile(eager, dynamic=True)  # reference server.py:47
            for name in names:
                lengths, _ = workload(name)
                mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1, device=dev)
                ts = torch.linspace(0, 1, N_STEPS_ODE + 1, device=dev)

                def amp_c():
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return ora.solve(z, ts, mu, mask)
                ms = time_it(amp_c, n_timed)
                out[name]["compile_autocast_bf16"] = {"ms_per_decode": ms, "value": sum(lengths) / ms * 1e3, "unit": UNIT}
            out["compile_s"] = round(time.perf_counter() - t0, 1)
            ora.estimator = eager
        except Exception as e:  # noqa: BLE001
            out["compile_error"] = f"{type(e).__name__}: {e}"[:300]
    return out


def cfg3_strong_scaling(n_gpus: int, precision: str):
    """BASELINE config 3 as the product runs it: ONE process drives every GPU of the box (ShardedCFM: one library handle and host
    thread per GPU, utterances dealt by cost, per-utterance H2D / D2H at the original indices, no collective).  Measured by rank 0
    while the other ranks of the torchrun job wait on a CPU barrier.  t1 = the same 256-utterance batch on one GPU through the same
    host-buffer entry.  Wall clock around the call (it returns when the result is on the host)."""
    import matcha_tts_24k_b200 as P
    lengths = P.synthetic.config_lengths("cfg3")
    cp = types.SimpleNamespace(solver="euler", sigma_min=1e-4, use_mu_prior=True)
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=5)
    mu, z = mu.pin_memory(), z.pin_memory()
    out = torch.empty_like(mu).pin_memory()
    ts = torch.linspace(0, 1, N_STEPS_ODE + 1)

    def run(devices, reps):
        sh = P.ShardedCFM(200, 100, cp, P.synthetic.PROD, devices=devices, precision=precision)
        for r in sh.replicas:
            P.synthetic.fill_named_seed(r.estimator, 1234)
        for _ in range(3):  # plan, direct-launch decode, graph capture
            sh.solve_host(z, ts, mu, lengths, out=out)
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            sh.solve_host(z, ts, mu, lengths, out=out)
            times.append((time.perf_counter() - t0) * 1e3)
        shards = sh.last_shards
        sh.close()
        return statistics.median(times), shards

    t1, _ = run([0], 3)
    tn, shards = run(list(range(n_gpus)), 7)
    cost = [sum(P.sharding.utterance_cost(lengths[i], 384) for i in s) for s in shards]
    return {"workload": "cfg3: B=256 mixed 2-12 s, mask-packed, 10 Euler steps, host buffers (H2D + decode + D2H per GPU)",
            "api": "ShardedCFM.solve_host -> cfm_solve_host_indexed / cfm_synchronize, one process, one host thread per GPU",
            "n_gpus": n_gpus, "t1_ms": t1, "tn_ms": tn, "speedup": t1 / tn, "value": sum(lengths) / (tn * 1e-3), "unit": UNIT,
            "imbalance": max(cost) / (sum(cost) / len(cost)), "utterances_per_gpu": [len(s) for s in shards],
            "frames_per_gpu": [sum(lengths[i] for i in s) for s in shards]}


def kernel_table(rows, peak_tflops):
    """Aggregates cfm_debug_timeline rows by launch class: in-situ time, share of the step, achieved TFLOP/s."""
    agg = {}
    for tag, M, N, K, flops, us in rows:
        a = agg.setdefault(tag, {"launches": 0, "us": 0.0, "flops": 0.0})
        a["launches"] += 1
        a["us"] += us
        a["flops"] += flops
    total = sum(a["us"] for a in agg.values()) or 1.0
    table = []
    for tag, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        tf = a["flops"] / (a["us"] * 1e-6) / 1e12 if a["us"] > 0 else 0.0
        table.append({"kernel": tag, "launches": a["launches"], "ms": a["us"] / 1e3, "share": a["us"] / total,
                      "tflops": tf, "frac": tf / peak_tflops})
    return table, total / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32_tc"],
                    help="bf16 (default) | fp32 = fp32-FMA parity reference | fp32_tc = fp32 storage, bf16 x 3 split operands on the tensor pipe")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the PyTorch-on-the-same-GPU baseline leg")
    ap.add_argument("--no-compile-baseline", action="store_true", help="skip the torch.compile part of that leg")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the extra cfg3 strong-scaling record (ShardedCFM driven by rank 0)")
    ap.add_argument("--ode-steps", type=int, default=N_STEPS_ODE, help="Euler steps per decode (BASELINE config 5 sweeps 2/4/10/32)")
    ap.add_argument("--solver", default="euler", choices=["euler", "midpoint", "heun3", "rk4"],
                    help="fixed-grid solver (the reference ships midpoint with 4 steps: matcha/inference.py:39-40)")
    ap.add_argument("--spks", type=int, default=0, help="speaker-vector width S for upstream-style conditioning (config 5: 96)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import matcha_tts_24k_b200 as P
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    all_lengths, desc = workload(args.workload)
    if args.workload == "cfg3" and world > 1:  # one batch sharded by utterance; T stays the batch max
        T = 2 * ((max(all_lengths) + 1) // 2)
        lengths = [all_lengths[i] for i in P.shard_utterances(all_lengths, world)[rank]]
        scaling, job_frames = "strong", sum(all_lengths)
    else:  # every rank decodes its own copy of the workload
        T = 2 * ((max(all_lengths) + 1) // 2)
        lengths, scaling, job_frames = all_lengths, "weak", sum(all_lengths) * world
    cp = types.SimpleNamespace(solver=args.solver, sigma_min=1e-4, use_mu_prior=True)
    model = P.CFM(200 + args.spks, 100, cp, P.synthetic.PROD, precision=args.precision, flags=args.flags).eval()
    P.synthetic.fill_named_seed(model.estimator, 1234)
    model = model.to(dev)
    mu, mask, z, _ = P.synthetic.make_inputs(lengths, seed=1 + rank, T=T)
    mu_h, z_h = mu.pin_memory(), z.pin_memory()
    mu, mask, z = mu.to(dev), mask.to(dev), z.to(dev)
    # the time grid is host-side metadata of the plan: a CPU tensor keeps the per-step call free of a device-to-host sync, so the K
    # timed steps are queued back to back (with a device tensor every step waited for the previous one, and the nvidia-smi
    # sampler's driver calls then showed up as 2-7 ms of idle GPU per step)
    ts = torch.linspace(0, 1, args.ode_steps + 1)
    spks = torch.randn(len(lengths), args.spks, generator=torch.Generator().manual_seed(3)).to(dev) if args.spks else None

    def barrier():
        torch.cuda.synchronize(dev)
        if dist:
            dist.barrier()
            torch.cuda.synchronize(dev)

    step_ms, host_ms = [], []

    def timed(fn, steps, lead_in=0):
        """K steps queued back to back between two events.  lead_in: untimed steps queued right before the first event without a
        synchronisation in between, their results held at the same time.  The loop below keeps the previous result alive while the
        next decode allocates its output, so the SECOND queued decode needs a second output block from torch's caching allocator;
        when the cache has none that is a cudaMalloc (+ a release of cached blocks, which waits for the device) inside the call -
        1-140 ms of host time (`run.step_enqueue_ms`) during which the queue ran dry and the second step looked 2-5x longer.  The
        lead-in puts both blocks into the cache before the first event."""
        gc.collect()
        gc.disable()  # no collection pauses between the enqueues of the timed steps
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]  # one per step: spread of the steps inside the region
        hold = [fn() for _ in range(lead_in)]  # two results alive at once, as in the loop below (`out` + the new one)
        del hold
        w0 = time.perf_counter()
        e0.record()
        host_ms.clear()
        for i in range(steps):
            h0 = time.perf_counter()
            out = fn()
            host_ms.append((time.perf_counter() - h0) * 1e3)
            marks[i].record()
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        ms = e0.elapsed_time(e1)
        step_ms[:] = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(steps)]
        gc.enable()
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out

    dev_step = lambda: model.solve(z, ts, mu, mask, lengths=lengths, spks=spks)
    out_pinned = torch.empty_like(mu_h).pin_memory()  # the caller's result buffer (pinned: asynchronous device-to-host copy)
    spks_h = spks.cpu() if spks is not None else None
    host_step = lambda: model.solve_host(z_h, ts, mu_h, lengths, device=dev, spks=spks_h, out=out_pinned)
    # The sampler (an nvidia-smi child) starts BEFORE the warm-up so that its start-up never lands in the timed region;
    # only the samples taken between the two marks are reported.
    with ClockSampler(local_rank) as clk:
        t_w = time.perf_counter()
        n_w = 0
        while n_w < args.warmup or time.perf_counter() - t_w < 2.5:  # >= W steps and >= 2.5 s: the decode runs at the board's power
            # cap, and the cap controller needs a couple of seconds to settle (sporadic +15-20 % first blocks were seen with 1.5 s)
            out = dev_step()
            torch.cuda.synchronize(dev)
            n_w += 1
        lo = clk.mark()
        ms_total, _, out = timed(dev_step, args.steps, lead_in=2)
        dev_step_ms, dev_host_ms = list(step_ms), list(host_ms)
        time.sleep(0.12)
        hi = clk.mark()
    clocks = clk.summary(lo, max(hi, lo + 1))
    for _ in range(3):
        host_step()
    _, wall_ms_e2e, out_h = timed(host_step, args.steps)
    if not bool(torch.isfinite(out).all()) or not bool(torch.isfinite(out_h).all()):
        raise SystemExit("bench.py: non-finite decode output")

    info = model.plan_info()
    # ---- the call the reference's callers make: CFM.forward(mu, mask, n) with device tensors (mask -> lengths on the host, seed-42
    # noise, cached time grid): matcha/inference.py:169.  Also for cfg1 (B = 1, the server's request shape) incl. never-seen lengths.
    fwd = {}
    if not args.spks and args.solver == "euler":
        for _ in range(2):
            model(mu, mask, args.ode_steps)
        ms_f, _, _ = timed(lambda: model(mu, mask, args.ode_steps), max(3, args.steps // 2))
        fwd[args.workload] = {"ms_per_call": ms_f / max(3, args.steps // 2), "api": "CFM.forward(mu, mask, n_timesteps), device tensors"}
        if world == 1:
            mu1, mask1, _, _ = P.synthetic.make_inputs([150], seed=2, device=dev)
            for _ in range(3):
                model(mu1, mask1, N_STEPS_ODE)
            ms1, _, _ = timed(lambda: model(mu1, mask1, N_STEPS_ODE), 20)
            news = []
            for L in (131, 163, 197, 211, 89, 240):  # first request of a length: plan (tables + workspace) + direct-launch decode
                mu_n, mask_n, _, _ = P.synthetic.make_inputs([L], seed=3, device=dev)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                model(mu_n, mask_n, N_STEPS_ODE)
                torch.cuda.synchronize(dev)
                news.append((time.perf_counter() - t0) * 1e3)
            fwd["cfg1"] = {"ms_per_call": ms1 / 20, "ms_first_call_new_length": statistics.median(news),
                           "api": "CFM.forward, B=1 L=150, 10 Euler steps (graph replay) / six never-seen lengths (wall clock, median)"}
            model.solve(z, ts, mu, mask, lengths=lengths, spks=spks)  # make the benchmark plan current again
    # ---- in-situ per-kernel times of one direct-launch decode (cfm_debug_timeline), for roofline.kernel
    ktable, k_total_ms = [], None
    try:
        rows = model.timeline(z, ts, mu, lengths) if not args.spks else []
        ktable, k_total_ms = kernel_table(rows, peaks()["tflops_sustained"])
    except Exception as e:  # noqa: BLE001
        ktable = [{"error": f"{type(e).__name__}: {e}"[:200]}]
    ms_step = ms_total / args.steps
    value = job_frames / (ms_step * 1e-3)
    e2e_ms = wall_ms_e2e / args.steps
    pk = peaks()
    nfe = args.ode_steps * {"euler": 1, "midpoint": 2, "heun3": 3, "rk4": 4}[args.solver]
    flops = P.synthetic.algorithmic_flops(lengths, 384, nfe) + 8.0 * args.spks * 384 * sum(lengths) * nfe
    achieved = flops / (ms_step * 1e-3) / 1e12
    n_bytes = mu_h.numel() * 4
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_cfg2.json")
    if (args.workload == "cfg2" and args.precision == "bf16" and args.ode_steps == N_STEPS_ODE and args.solver == "euler"
            and not args.spks and os.path.exists(tpath)):
        traffic = json.load(open(tpath)).get("dram_bytes_per_decode")  # ncu dram__bytes_{read,write}.sum over every launch of one decode
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": workload_config(desc, args, T),
        "run": {"batch_per_gpu": len(lengths), "frames_per_gpu": sum(lengths), "rows_full": info["rows_full"],
                "workspace_mib": round(info["workspace_bytes"] / 2**20), "schedule": "whole ODE loop = one CUDA graph",
                "step_ms": [round(v, 3) for v in dev_step_ms], "step_enqueue_ms": [round(v, 3) for v in dev_host_ms],
                "parallelism": "utterance-sharded replicas, no collective" if world > 1 else "single GPU"},
        "rtf": (ms_step * 1e-3) / (job_frames * FRAME_SECONDS),
        "e2e": {"value": job_frames / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 2 * n_bytes, "d2h_bytes_per_step": n_bytes, "api": "cfm_solve_host (C ABI, pinned host buffers)" if not args.spks else "cfm_solve_host_spks (C ABI, pinned host buffers)"},
        "gpu_launches": int(info["kernels_per_solve"]) * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "peak_source": pk["source"] + ", sustained bf16",
                     "scope": "whole decode = one graph launch (the tensor-core GEMM and attention kernels carry > 97 % of the FLOPs)",
                     "algorithmic_flops_per_launch": flops},
        "clocks": clocks,
    }
    dom = next((k for k in ktable if k.get("tflops", 0) > 0), None)
    if dom:  # the launch class with the largest share of the step, timed in situ (direct launches, one event per launch)
        line["roofline"].update({"kernel": dom["kernel"], "kernel_achieved": dom["tflops"], "kernel_frac": dom["frac"],
                                 "kernel_share_of_step": dom["share"], "kernel_launches_per_step": dom["launches"],
                                 "kernel_timing": "CUDA events around every launch of one direct-launch decode (cfm_debug_timeline)"})
    # north_star's target is worded on the estimator GEMMs: every launch class that carries FLOPs except attention, in situ.  The
    # per-launch events of the direct-launch decode include the host's launch gaps (timeline_ms > ms_per_step), so the same figure
    # is also given with all classes scaled by ms_per_step / timeline_ms (the decode as the graph runs it).
    gemm = [k for k in ktable if k.get("tflops", 0) > 0 and k.get("kernel") not in ("attention", "time_mlp")]
    if gemm and k_total_ms:
        g_ms = sum(k["ms"] for k in gemm)
        g_fl = sum(k["tflops"] * k["ms"] for k in gemm)  # TFLOP/s * ms = GFLOP
        scale = ms_step / k_total_ms if world == 1 else 1.0
        line["roofline"]["gemm"] = {"classes": len(gemm), "ms_timeline": g_ms, "achieved_timeline": g_fl / g_ms,
                                    "frac_timeline": g_fl / g_ms / pk["tflops_sustained"],
                                    "achieved_in_graph_est": g_fl / (g_ms * scale), "frac_in_graph_est": g_fl / (g_ms * scale) / pk["tflops_sustained"],
                                    "share_of_step": g_ms / k_total_ms}
    line["kernels"] = {"timeline_ms": k_total_ms, "by_launch_class": ktable}
    if fwd:
        line["forward_api"] = fwd
    if rank == 0 and world == 1 and not args.no_gpu_baseline and args.workload == "cfg2" and not args.spks:
        model.close()  # release the workspace before PyTorch allocates its own activations
        line["gpu_pytorch_baseline"] = gpu_pytorch_baseline(dev, with_compile=not args.no_compile_baseline)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        sample = lengths[:2]
        fps, sec = cpu_oracle_throughput(sample, 2, threads)
        torch.set_num_threads(threads)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{len(sample)} utterances of {desc}, full 10-step Euler decode, median of 2 (~{2 * sec:.0f} s)"}
    if dist and args.workload == "cfg2" and not args.no_strong:
        # one process (rank 0) drives all GPUs through the product's multi-GPU API; the others wait on the CPU
        model.close()
        torch.cuda.synchronize(dev)
        cpu_group = dist.new_group(backend="gloo")
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                line["extra"] = {"cfg3_strong": cfg3_strong_scaling(world, args.precision)}
            except Exception as e:  # noqa: BLE001
                line["extra"] = {"cfg3_strong": {"error": f"{type(e).__name__}: {e}"[:300]}}
        dist.barrier(group=cpu_group)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
