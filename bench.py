#!/usr/bin/env python
"""bench.py — TI train images/s of the ViewNeTI hot path (SD-2.1, 512^2 => 64x64x4 latents, 77x1024 contexts).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--latent 64]
    N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one train image per GPU: frozen UNet forward + fp32 MSE + dgrad-only backward to the 32 XTI context
tensors (reference training/coach.py:197-214), batch-parallel over ranks (per-GPU batch 1, BASELINE config 2/3),
plus — for N > 1 — the single NCCL all-reduce of the flat mapper-gradient buffer.
Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM (CUDA-graph replay);
`e2e` = the same metric through the drop-in Python API (UNet2DConditionModel.__call__ + mse_loss + backward) with
pinned-host inputs copied in every step and the loss read back.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per train image (SURVEY.md 8d / BASELINE.md 3): 2*MAC of every conv / linear / attention GEMM
GFLOP_TRAIN = {64: 1671.3, 32: 362.9}
GFLOP_FWD = {64: 804.3, 32: 181.1}
MAPPER_GRAD_ELEMS = 2 * 141696          # M_v + active M_o
METRIC = "TI train images/sec (SD2.1, 512^2)"


def workload(L: int) -> str:
    """config.workload - the SAME string in both arms (BASELINE config 2, per GPU)."""
    return (f"mode 2 single-scene TI step, SD2.1 {L * 8}x{L * 8}, bs=1 per GPU ({L}x{L}x4 latents, 16+16 contexts 77x1024): "
            f"UNet fwd + fp32 MSE + backward to the 32 contexts")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def oracle_train_step_time(latent: int, steps: int, warmup: int, budget_s: float):
    """Times the fp32 CPU oracle (our restatement of the reference's diffusers path; the reference itself cannot be
    installed here) on the same workload.  Returns (seconds per image, images timed, threads)."""
    from oracle.unet_sd21 import UNetOracle, train_step_oracle
    from tests.unet_parity import ctx_to, make_inputs
    from view_neti_b200.sd21 import SD21, init_state_dict
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    unet = UNetOracle(SD21)
    unet.load_state_dict(init_state_dict(SD21, 0))
    lat, t, tgt, ctx = make_inputs(SD21, 1, latent, latent, seed=1)
    t_start = time.time()
    done_w, times = 0, []
    while done_w < warmup and (time.time() - t_start) < budget_s * 0.4:
        train_step_oracle(unet, lat, t, tgt, ctx_to(ctx, "cpu"))
        done_w += 1
    while len(times) < steps:
        t0 = time.time()
        train_step_oracle(unet, lat, t, tgt, ctx_to(ctx, "cpu"))
        times.append(time.time() - t0)
        if time.time() - t_start > budget_s:
            break
    return sum(times) / len(times), len(times), threads, done_w


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    per, n, threads, done_w = oracle_train_step_time(args.latent, max(1, args.steps), max(1, args.warmup), budget_s=200.0)
    v = 1.0 / per
    sample = (f"{n} timed + {done_w} warm-up train images (fwd + fp32 MSE + autograd backward to the 32 contexts) at "
              f"{args.latent}x{args.latent} latents, B=1, fp32 PyTorch eager on the host CPU; wall budget 200 s "
              f"(requested steps={args.steps}, warmup={args.warmup})")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": n,
        "warmup": done_w, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": workload(args.latent),
                   "note": "reference arm = oracle port of the reference's diffusers CPU path (diffusers/accelerate "
                           "are not installable offline; see DESIGN.md)"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def profile_categories(plan):
    """One instrumented eager train step after the timed region: per-kernel device durations from CUPTI
    (torch.profiler, no replay), grouped per kernel family; GEMM / conv FLOPs are counted from the call arguments."""
    from torch.profiler import ProfilerActivity, profile
    from view_neti_b200 import ops
    flops = {"gemm": 0.0, "conv": 0.0}
    orig = {"gemm": ops.gemm, "conv3x3": ops.conv3x3}

    def gemm(A, B, D, **k):
        flops["gemm"] += 2.0 * (A.numel() // A.shape[-1]) * B.shape[0] * B.shape[1]
        return orig["gemm"](A, B, D, **k)

    def conv3x3(x, Wk, D, **k):
        flops["conv"] += 2.0 * (x.numel() // x.shape[-1]) * Wk.shape[0] * Wk.shape[1]
        return orig["conv3x3"](x, Wk, D, **k)

    ops.gemm, ops.conv3x3 = gemm, conv3x3
    ops.set_pdl(False)          # isolate kernels: with PDL a kernel's duration includes waiting for its predecessor
    try:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            plan.train_step()
            torch.cuda.synchronize()
    finally:
        ops.set_pdl(True)
        ops.gemm, ops.conv3x3 = orig["gemm"], orig["conv3x3"]
    fam = {"vn_gemm_kernel": "gemm", "attn_": "attn", "gn_": "groupnorm", "ln_kernel": "layernorm", "geglu": "geglu"}
    ms, n = {}, {}
    for e in prof.key_averages():
        dt = getattr(e, "self_device_time_total", 0) or 0
        if dt <= 0:
            continue
        cat = next((v for k, v in fam.items() if k in e.key), "other")
        ms[cat] = ms.get(cat, 0.0) + dt / 1e3
        n[cat] = n.get(cat, 0) + e.count
    return ms, n, flops


def _emit(line: dict, real_stdout) -> None:
    """The ONE JSON line, on the process's original stdout (see run_ours for why fd 1 may have been redirected)."""
    text = json.dumps(line) + "\n"
    if real_stdout is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(real_stdout, text.encode())


def _finish(world: int) -> None:
    """End of a rank.  Multi-process runs leave without tearing the process group down: ncclCommDestroy with the step graph
    (which holds the captured all-reduce) still alive blocked rank 0 for the whole 900 s limit of one test run, and nothing
    after this point needs the communicator; buffers are flushed and the process exits with status 0."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def _time_replays(graph_or_fn, n: int, e0, e1) -> float:
    """ms per call over n back-to-back calls (CUDA events on the current stream)."""
    call = graph_or_fn.replay if hasattr(graph_or_fn, "replay") else graph_or_fn
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist
    import torch.nn.functional as F
    from view_neti_b200 import ops
    from view_neti_b200.sd21 import SD21, init_state_dict
    from view_neti_b200.unet import UNet2DConditionModel
    from tests.unet_parity import make_inputs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    real_stdout = None
    if world > 1:
        # stdout carries ONE JSON line, but NCCL writes its version banner and its NCCL_DEBUG=INFO log (which the driver reads
        # the rank count from - it is left exactly as the environment sets it) to file descriptor 1.  For the whole run fd 1
        # is therefore pointed at stderr; the JSON line is written to the saved original descriptor at the end.
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ["NCCL_DEBUG"] = "INFO"               # the communicator's own account of itself (nranks, NVLS, channels) -> stderr
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
        print(f"bench.py: rank {rank} of {dist.get_world_size()} (backend {dist.get_backend()}) on cuda:{local_rank}", file=sys.stderr, flush=True)
    L = args.latent
    cfg = SD21
    model = UNet2DConditionModel(init_state_dict(cfg, 0), cfg, dev)
    plan = model.engine.plan(1, L, L)
    lat, t, tgt, ctx = make_inputs(cfg, 1, L, L, seed=1 + rank)       # per-rank samples (weak scaling)
    plan.latents.copy_(lat); plan.timesteps.copy_(t); plan.target.copy_(tgt)
    for i in range(cfg.num_cross_layers):
        plan.ctx[0, i].copy_(ctx[f"CONTEXT_TENSOR_{i}"]); plan.ctx[1, i].copy_(ctx[f"CONTEXT_TENSOR_BYPASS_{i}"])
    flat = torch.zeros(MAPPER_GRAD_ELEMS, device=dev)               # flat mapper-gradient buffer (M_v + M_o)

    graph = plan.capture("train")
    launches_per_step = plan.launches["train"]
    tail = None
    tail_note = None
    if world > 1:
        # The exchange of the step: ONE all-reduce of a buffer of the mapper-gradient size that is
        # data-dependent on this step's d_ctx, as the TAIL of the step's CUDA graph (NCCL is graph-capturable): pack,
        # all-reduce, scale - three nodes after the backward, no host round trip.  (The REAL mapper gradients travel in the
        # `full_step` leg below, through Coach.train_step.)
        def exchange():
            flat.copy_(plan.d_ctx[:, :2].reshape(-1)[:MAPPER_GRAD_ELEMS])
            dist.all_reduce(flat)
            flat.mul_(1.0 / world)

        exchange()
        torch.cuda.synchronize()
        try:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                plan.train_step()
                exchange()
            tail, tail_note = g2, "all-reduce captured as the tail of the step graph"
            launches_per_step += 3
        except Exception as e:          # capture refused by this NCCL / torch build: eager exchange after the replay
            tail_note = "all-reduce issued eagerly after the graph replay (capture failed: %s)" % (f"{type(e).__name__}: {e}"[:80])
            torch.cuda.synchronize()

    def step():
        if tail is not None:
            tail.replay()
            return
        graph.replay()
        if world > 1:
            exchange()

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_total = float(ms)
        # (f) the same loop over >= 500 steps (a sustained figure next to the K-step value the driver asks for); still inside
        # the clock-sampling window
        n_long = max(500, args.steps)
        long_ms = torch.tensor([_time_replays(step, n_long, e0, e1)], device=dev)
        if world > 1:
            dist.all_reduce(long_ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms_total / args.steps
    value = world * args.steps / (ms_total / 1e3)
    sustained = {"steps": n_long, "ms_per_step": float(long_ms), "value": world * 1e3 / float(long_ms), "unit": "images/s"}
    loss_dev = float(plan.loss)

    # ---- end to end through the drop-in API, host-resident inputs ------------------------------------------
    pin = lambda x: x.clone().pin_memory()      # noqa: E731
    h_lat, h_t, h_tgt = pin(lat), pin(t), pin(tgt)
    nl = cfg.num_cross_layers
    # the host keeps the 16 + 16 context tensors of a step in one pinned block (one H2D copy); the dict handed to the
    # UNet holds per-layer views of it, exactly the reference's {"this_idx", "CONTEXT_TENSOR_i", "..._BYPASS_i"} protocol
    h_ctx = pin(torch.stack([torch.stack([ctx[f"CONTEXT_TENSOR_{i}"] for i in range(nl)]),
                             torch.stack([ctx[f"CONTEXT_TENSOR_BYPASS_{i}"] for i in range(nl)])]))
    h2d = sum(x.numel() * x.element_size() for x in [h_lat, h_t, h_tgt, h_ctx])

    def e2e_step():
        d_lat = h_lat.to(dev, non_blocking=True)
        d_t = h_t.to(dev, non_blocking=True)
        d_tgt = h_tgt.to(dev, non_blocking=True)
        d_all = h_ctx.to(dev, non_blocking=True)                       # stands in for the mapper/CLIP output
        d_ctx = {"this_idx": 0}
        for i in range(nl):                                             # leaf tensors (views of the block, no copies)
            d_ctx[f"CONTEXT_TENSOR_{i}"] = d_all[0, i].detach().requires_grad_(True)
            d_ctx[f"CONTEXT_TENSOR_BYPASS_{i}"] = d_all[1, i].detach().requires_grad_(True)
        pred = model(d_lat, d_t, d_ctx).sample                       # coach.py:197-198
        loss = F.mse_loss(pred.float(), d_tgt.float(), reduction="mean")   # :211-213
        loss.backward()                                                    # :214
        if world > 1:
            flat.copy_(torch.cat([d_ctx[k].grad.view(-1) for k in ("CONTEXT_TENSOR_0", "CONTEXT_TENSOR_BYPASS_0",
                                                                   "CONTEXT_TENSOR_1", "CONTEXT_TENSOR_BYPASS_1")]
                                 )[:MAPPER_GRAD_ELEMS])
            dist.all_reduce(flat)
        return float(loss.detach().cpu())                                 # D2H read of the step's result (:257)

    ops.launch_count_reset()
    for _ in range(3):
        e2e_loss = e2e_step()
    e2e_launches_eager = ops.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k2 = max(3, min(args.steps, 20))
    e0.record()
    for _ in range(k2):
        e2e_loss = e2e_step()
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * k2 / (float(ms2) / 1e3)
    e2e_note = "inputs copied at the start of their step"
    pf_steps = 0

    # The same loop as a training job feeds it (pinned-memory loader with prefetch): the H2D copy of step i+1 is issued on
    # a copy stream while step i computes.  Every timed step still copies one step's inputs and reads its loss back.
    # Single-process only (no collective inside, so a failure here cannot desynchronise ranks); kept only if the loss is
    # the same number and the loop is faster, else the figure above stands.
    if world == 1:
        try:
            copy_stream = torch.cuda.Stream()
            main_stream = torch.cuda.current_stream()

            def stage():
                with torch.cuda.stream(copy_stream):
                    bufs = [h.to(dev, non_blocking=True) for h in (h_lat, h_t, h_tgt, h_ctx)]
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                for b in bufs:
                    b.record_stream(main_stream)
                return bufs, ev

            staged = [stage()]

            def e2e_step_prefetch():
                (d_lat, d_t, d_tgt, d_all), ev = staged.pop()
                main_stream.wait_event(ev)
                d_ctx = {"this_idx": 0}
                for i in range(nl):
                    d_ctx[f"CONTEXT_TENSOR_{i}"] = d_all[0, i].detach().requires_grad_(True)
                    d_ctx[f"CONTEXT_TENSOR_BYPASS_{i}"] = d_all[1, i].detach().requires_grad_(True)
                pred = model(d_lat, d_t, d_ctx).sample
                loss = F.mse_loss(pred.float(), d_tgt.float(), reduction="mean")
                loss.backward()
                staged.append(stage())                                    # next step's inputs travel under this step
                return float(loss.detach().cpu())

            for _ in range(3):
                pf_loss = e2e_step_prefetch()
                pf_steps += 1
            torch.cuda.synchronize()
            e0.record()
            for _ in range(k2):
                pf_loss = e2e_step_prefetch()
                pf_steps += 1
            e1.record()
            torch.cuda.synchronize()
            pf_value = k2 / (e0.elapsed_time(e1) / 1e3)
            if abs(pf_loss - e2e_loss) <= 1e-6 * abs(e2e_loss) and pf_value > e2e_value:
                e2e_note = ("inputs of step i+1 prefetched on a copy stream during step i (each step still copies one "
                            "step's inputs and reads its loss); without prefetch: %.2f images/s" % e2e_value)
                e2e_value = pf_value
            else:
                e2e_note += "; prefetch variant not used (%.2f images/s, loss %.6f)" % (pf_value, pf_loss)
        except Exception as e:          # the plain figure stands
            e2e_note += "; prefetch variant failed: " + f"{type(e).__name__}: {e}"[:120]
    api_plan = model.engine.plan(1, L, L)
    e2e_launches = api_plan.launches.get("fwd", 0) + api_plan.launches.get("bwd", 0)
    extra_launches = 0

    # ---- the COMPLETE train step (SURVEY 8f #1/#2 widening): batched conditioning path (CUDA mappers + 23-layer CLIP
    # encoder) -> UNet -> MSE -> backward into the mapper parameters -> all-reduce of the REAL mapper gradients (world > 1)
    # -> AdamW, through Coach.train_step.  Runs on every rank count: this is the step whose scaling the reference's DDP
    # wrapper is about (coach.py:97-99,214). ------------
    full_step = None
    try:
        from view_neti_b200.training.coach import Coach
        from view_neti_b200.training.synthetic import build_conditioning, synthetic_prompt
        cond = build_conditioning(dev)
        coach = Coach(cfg=None, unet=model, conditioning=cond, optimizer=torch.optim.AdamW(cond.parameters(), lr=1e-3),
                      generator=torch.Generator(device=dev).manual_seed(1 + rank))
        prompt = synthetic_prompt(1, dev)
        lat0 = torch.randn(1, 4, L, L, device=dev)
        for _ in range(4):
            coach.train_step(lat0, prompt)
        if world > 1:
            dist.barrier()
        kf = max(3, min(args.steps, 20))
        fs = torch.tensor([_time_replays(lambda: coach.train_step(lat0, prompt), kf, e0, e1)], device=dev)
        if world > 1:
            dist.all_reduce(fs, op=dist.ReduceOp.MAX)
        fs_ms = float(fs)
        fl_loss = coach.train_step(lat0, prompt)
        full_step = {"value": world * 1e3 / fs_ms, "unit": "images/s", "ms_per_step": fs_ms, "steps": kf, "n_gpus": world,
                     "what": "Coach.train_step: NeTI mappers + batched 16-layer CLIP-H conditioning (23-layer encoder on "
                             "[16,77,1024]) + UNet fwd/bwd + mapper gradients" +
                             (" + all-reduce of the %d mapper gradients" % sum(p.numel() for p in cond.parameters()) if world > 1 else "") +
                             " + AdamW; synthetic prompt, seeded weights, per-rank noise / timesteps",
                     "trainable_params": sum(p.numel() for p in cond.parameters()), "loss": float(fl_loss)}
        if world > 1:
            # DDP semantics hold: parameters stay bit-identical across ranks
            chk = torch.cat([p.detach().reshape(-1) for p in cond.parameters()]).double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            full_step["params_identical_across_ranks"] = bool(float(lo) == float(hi))
        if world == 1:
            # the same step started from the image, as reference coach.py:165-169 does every step: VAE encode first
            from view_neti_b200.models.vae import SD21_VAE, AutoencoderKL, init_state_dict as vae_init
            coach.vae = AutoencoderKL(vae_init(SD21_VAE, 0), SD21_VAE, dev)
            pb = dict(prompt)
            pb["pixel_values"] = torch.rand(1, 3, 8 * L, 8 * L, device=dev) * 2 - 1
            for _ in range(3):
                coach.train_step(batch=pb)
            px_ms = _time_replays(lambda: coach.train_step(batch=pb), kf, e0, e1)
            full_step["from_pixel_values"] = {"value": 1e3 / px_ms, "unit": "images/s", "ms_per_step": px_ms,
                                              "what": "the same step with vae.encode(pixel_values [1,3,%d,%d]) in front" % (8 * L, 8 * L)}
        del coach, cond
    except Exception as e:          # the headline metric above must survive a failure of the widened path
        full_step = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- BASELINE config 4 (mode 3 multi-scene pretraining, train_m3.yaml): shared M_v + 14 per-scene M_o resident, ONE object
    # active per step and the same one on every rank (rank 0 draws, everybody follows), all-reduce of M_v + that M_o only -----
    mode3 = None
    try:
        from types import SimpleNamespace
        from view_neti_b200.training.coach import Coach
        from view_neti_b200.training.synthetic import build_conditioning, object_token_id, synthetic_prompt
        n_obj = 14
        cond3 = build_conditioning(dev, n_objects=n_obj)

        class _Objects:            # the part of the mode-3 dataset the step needs: which object the next batch shows
            learnable_mode = 3

            def __init__(self):
                self.g = torch.Generator().manual_seed(5 + rank)
                self.current_object_idx = 0

            def reset_sampled_object(self, idx=None):
                self.current_object_idx = int(torch.randint(0, n_obj, (1,), generator=self.g)) if idx is None else int(idx)
                return self.current_object_idx

        coach3 = Coach(cfg=SimpleNamespace(learnable_mode=3), unet=model, conditioning=cond3,
                       optimizer=torch.optim.AdamW(cond3.parameters(), lr=1e-3),
                       generator=torch.Generator(device=dev).manual_seed(2 + rank), train_dataset=_Objects())
        prompts = [synthetic_prompt(1, dev, object_id=object_token_id(i)) for i in range(n_obj)]
        lat3 = torch.randn(1, 4, L, L, device=dev)
        seen = []

        def step3():
            i = coach3.reset_sampled_object()
            seen.append(i)
            return coach3.train_step(lat3, prompts[i])

        for _ in range(4):
            step3()
        if world > 1:
            dist.barrier()
        k3 = max(3, min(args.steps, 20))
        m3 = torch.tensor([_time_replays(step3, k3, e0, e1)], device=dev)
        if world > 1:
            dist.all_reduce(m3, op=dist.ReduceOp.MAX)
        mode3 = {"value": world * 1e3 / float(m3), "unit": "images/s", "ms_per_step": float(m3), "steps": k3, "n_gpus": world,
                 "object_mappers": n_obj, "objects_visited": len(set(seen)),
                 "trainable_params": sum(p.numel() for p in cond3.parameters()),
                 "what": "Coach.train_step in learnable mode 3: 14 object mappers + 1 view mapper resident, the step's object drawn on "
                         "rank 0 and broadcast, gradients of M_v + the active M_o all-reduced (283 392 floats), inactive mappers untouched"}
        if world > 1:
            chk = torch.cat([p.detach().reshape(-1) for p in cond3.parameters()]).double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            mode3["params_identical_across_ranks"] = bool(float(lo) == float(hi))
        del coach3, cond3
    except Exception as e:
        mode3 = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- BASELINE config 5 (single process): pretrained M_v FROZEN (requires_grad False: no gradient is computed for it) +
    # learn M_o, then novel-view inference = PromptManager.embed_prompt over the 50 timesteps + sd_pipeline_call (50 steps x
    # batched CFG UNet forward + fused guidance / sampler update + VAE decode) ---------------------------------------------
    config5 = None
    if world == 1:
        try:
            from view_neti_b200.models.vae import SD21_VAE, AutoencoderKL, init_state_dict as vae_init
            from view_neti_b200.prompt_manager import PromptManager
            from view_neti_b200.schedulers import DPMSolverMultistepScheduler
            from view_neti_b200.sd_pipeline_call import ViewNeTIPipeline, sd_pipeline_call
            from view_neti_b200.training.coach import Coach
            from view_neti_b200.training.synthetic import OBJECT_TOKEN_ID, VIEW_TOKEN_IDS, build_conditioning, synthetic_prompt
            cond5 = build_conditioning(dev)
            cond5.mapper_view.requires_grad_(False)                     # mode 5 (coach.py:661-669): M_v loaded and frozen
            coach5 = Coach(cfg=None, unet=model, conditioning=cond5,
                           optimizer=torch.optim.AdamW([p for p in cond5.parameters() if p.requires_grad], lr=1e-3),
                           generator=torch.Generator(device=dev).manual_seed(3))
            prompt5 = synthetic_prompt(1, dev)
            lat5 = torch.randn(1, 4, L, L, device=dev)
            for _ in range(4):
                coach5.train_step(lat5, prompt5)
            k5 = max(3, min(args.steps, 20))
            t5 = _time_replays(lambda: coach5.train_step(lat5, prompt5), k5, e0, e1)
            n_inf = 50
            sched = DPMSolverMultistepScheduler("v_prediction")           # what the reference's inference scripts install (validate.py:568)
            sched.set_timesteps(n_inf)
            pm = PromptManager(tokenizer=None, text_encoder=cond5, timesteps=[int(t) for t in sched.timesteps],
                               placeholder_view_token_ids=VIEW_TOKEN_IDS, placeholder_object_token_ids=[OBJECT_TOKEN_ID], chunk=10)
            vae5 = AutoencoderKL(vae_init(SD21_VAE, 0), SD21_VAE, dev)
            neg = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(3)).to(dev)
            pipe = ViewNeTIPipeline(model, sched, negative_prompt_embeds=neg, vae=vae5)

            def infer(seed):
                emb = pm.embed_prompt(prompt5["input_ids"])
                return sd_pipeline_call(pipe, emb, height=8 * L, width=8 * L, num_inference_steps=n_inf, guidance_scale=7.5,
                                        generator=torch.Generator(device=dev).manual_seed(seed), output_type="np")

            infer(0)
            torch.cuda.synchronize()
            t0 = time.time()
            n_img = 2
            for i in range(n_img):
                img = infer(1 + i).images
            torch.cuda.synchronize()
            per_img = (time.time() - t0) / n_img
            config5 = {"train": {"value": 1e3 / t5, "unit": "images/s", "ms_per_step": t5, "steps": k5,
                                 "trainable_params": sum(p.numel() for p in cond5.parameters() if p.requires_grad),
                                 "what": "Coach.train_step with the view mapper frozen (mode 5): only M_o receives gradient"},
                       "inference": {"value": 1.0 / per_img, "unit": "images/s", "s_per_image": per_img, "images": n_img,
                                     "steps": n_inf, "what": f"embed_prompt (50 timesteps x 16 layers, batched) + 50-step DPM-Solver++ CFG 7.5 "
                                                             f"denoise at {8 * L}x{8 * L} + VAE decode, wall clock per image",
                                     "finite": bool(torch.isfinite(torch.as_tensor(img)).all())}}
            del coach5, cond5, vae5, pipe
        except Exception as e:
            config5 = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        _finish(world)
        return

    # ---- secondary legs (rank 0 of a single-process run) -----------------------------------------------------
    forward_leg = batch3 = None
    peaks, peak_src = measured_peaks()
    clk = clocks.summary()
    # (c) denominator from the clocks record: a timed region that ran at the maximum SM clock with no power cap is a burst
    # measurement; otherwise the sustained figure applies.  Both fractions are printed.
    at_max = bool(clk.get("sm_mhz") and clk.get("sm_max_mhz") and clk["sm_mhz"] >= 0.97 * clk["sm_max_mhz"]
                  and "sw_power_cap" not in clk.get("reasons", []))
    p_burst = peaks.get("bf16_tflops")
    p_sust = peaks.get("bf16_tflops_sustained", p_burst)
    peak = p_burst if at_max else p_sust
    peak_kind = "bf16_tflops (burst)" if at_max else "bf16_tflops_sustained"
    if world == 1:
        try:
            gf = plan.capture("fwd")
            for _ in range(3):
                gf.replay()
            f_ms = _time_replays(gf, max(100, args.steps), e0, e1)
            extra_launches += plan.launches["fwd"] * (3 + max(100, args.steps))
            f_tf = GFLOP_FWD.get(L, 0.0) * 1e9 / (f_ms * 1e-3) / 1e12
            forward_leg = {"ms": f_ms, "forwards_per_s": 1e3 / f_ms, "launches": plan.launches["fwd"],
                           "gflop": GFLOP_FWD.get(L), "achieved_tflops": f_tf,
                           "frac_of_burst": f_tf / p_burst if p_burst else None, "frac_of_sustained": f_tf / p_sust if p_sust else None,
                           "what": f"UNet forward only ({L}x{L} latents, B=1; the denoise-loop call of sd_pipeline_call.py:78-94), "
                                   f"forward CUDA graph, inputs resident; north_star target 0.5 of the tensor roofline"}
        except Exception as e:
            forward_leg = {"error": f"{type(e).__name__}: {e}"[:200]}
        try:
            # the reference's own micro-batch (input_configs/train.yaml:48: train_batch_size 3) on one GPU - SECONDARY: the
            # headline metric is quoted at per-GPU batch 1 (BASELINE configs 2 / 3)
            nb3 = 3
            plan3 = model.engine.plan(nb3, L, L)
            l3, t3, g3, c3 = make_inputs(cfg, nb3, L, L, seed=11)
            plan3.latents.copy_(l3); plan3.timesteps.copy_(t3); plan3.target.copy_(g3)
            for i in range(cfg.num_cross_layers):
                plan3.ctx[0, i].copy_(c3[f"CONTEXT_TENSOR_{i}"]); plan3.ctx[1, i].copy_(c3[f"CONTEXT_TENSOR_BYPASS_{i}"])
            g3g = plan3.capture("train")
            for _ in range(3):
                g3g.replay()
            b_ms = _time_replays(g3g, max(10, min(args.steps, 50)), e0, e1)
            extra_launches += plan3.launches["train"] * (3 + max(10, min(args.steps, 50)))
            b_tf = nb3 * GFLOP_TRAIN.get(L, 0.0) * 1e9 / (b_ms * 1e-3) / 1e12
            batch3 = {"per_gpu_batch": nb3, "ms_per_step": b_ms, "value": nb3 * 1e3 / b_ms, "unit": "images/s",
                      "launches_per_step": plan3.launches["train"], "achieved_tflops": b_tf,
                      "frac_of_burst": b_tf / p_burst if p_burst else None, "frac_of_sustained": b_tf / p_sust if p_sust else None,
                      "what": "same train step at the reference's default micro-batch 3 per GPU (train.yaml:48) - secondary"}
            del plan3
            model.engine._plans.pop((nb3, L, L), None)
        except Exception as e:
            batch3 = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- roofline of the dominant kernel (vn_gemm_kernel: every conv / linear), measured live ----------------
    cat_ms, cat_n, flops = profile_categories(plan)
    tot_ms = sum(cat_ms.values())
    gemm_ms = cat_ms.get("gemm", 0.0)
    gemm_flops = flops["gemm"] + flops["conv"]
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_tflops = GFLOP_TRAIN.get(L, 0.0) * 1e9 / (ms_per_step * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("vn_gemm_kernel_dram_bytes_per_launch")
    roofline = {
        "bound": "tensor", "kernel": "vn_gemm_kernel (tcgen05 GEMM + implicit-GEMM conv; all Linear/Conv fwd + dgrad)",
        "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s", "frac": gemm_tflops / peak if peak else None,
        "peak_source": f"{peak_src} {peak_kind}: the timed region ran at {clk.get('sm_mhz')} MHz of {clk.get('sm_max_mhz')} MHz, "
                       f"throttle reasons {clk.get('reasons')}",
        "frac_of_burst": gemm_tflops / p_burst if p_burst else None,
        "frac_of_sustained": gemm_tflops / p_sust if p_sust else None,
        "traffic": traffic,
        "launches_per_step": cat_n.get("gemm", 0),
        "flops_per_step": gemm_flops, "avg_launch_us": 1e3 * gemm_ms / max(1, cat_n.get("gemm", 0)),
        "share_of_step": gemm_ms / tot_ms if tot_ms else None,
        "timing": "CUPTI kernel durations of one eager step (PDL overlap off) run right after the timed region (same process, same clocks)",
        "kernel_ms_per_step": {k: round(v, 4) for k, v in cat_ms.items()},
        "category_launches": cat_n,
        "whole_step": {"gflop_per_image": GFLOP_TRAIN.get(L), "achieved": step_tflops,
                       "frac": step_tflops / peak if peak else None,
                       "frac_of_burst": step_tflops / p_burst if p_burst else None,
                       "frac_of_sustained": step_tflops / p_sust if p_sust else None},
    }

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample, same sampling as the reference arm ------
    cpu = None
    if world == 1:
        per, n, threads, done_w = oracle_train_step_time(L, steps=5, warmup=1, budget_s=25.0)
        cpu = {"value": 1.0 / per, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"{n} timed + {done_w} warm-up train image(s) (fwd + MSE + backward to 32 contexts) at {L}x{L} latents, "
                         f"fp32 PyTorch eager, {threads} threads, wall budget 25 s"}

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload(L),
                   "exchange": (tail_note if world > 1 else "none (one process)"),
                   "global_batch": world, "parallelism": f"dp{world}", "weights": "seeded random, SD-2.1 shapes (865.9M)",
                   "l2": "weights streamed per step (2 x 1.73 GB fwd + dgrad copies) exceed the 126 MB L2",
                   "graph": "one CUDA graph per step"},
        "sustained": sustained,
        "roofline": roofline, "forward": forward_leg, "batch3": batch3, "cpu_baseline": cpu, "full_step": full_step,
        "mode3": mode3, "config5": config5,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "steps": k2, "copies": e2e_note, "api": "UNet2DConditionModel.__call__ + F.mse_loss + backward (CUDA-graph replay inside)"},
        "gpu_launches": launches_per_step * (args.steps + n_long) + e2e_launches * (k2 + pf_steps) + e2e_launches_eager + extra_launches,
        "launches_per_step": launches_per_step,
        "clocks": clk,
        "loss": loss_dev, "e2e_loss": e2e_loss,
    }
    _emit(line, real_stdout)
    _finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--latent", type=int, default=64)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    if world != args.gpus:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
