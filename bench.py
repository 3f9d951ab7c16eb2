#!/usr/bin/env python
"""bench.py -- headline benchmark of the NeuCam hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

metric : scene-MLP fwd+bwd samples/s (1 sample = one scene-MLP query = one of the K taps of a pixel)
workload: BASELINE.json configs[1] -- config_mfme_blender.txt in its 9-image 3-focus x 3-exposure form:
          N,H,W = 9,300,450, sample_frac 3.5e-2 -> B = 42 525 px/step, K = 9 taps, Q = 382 725 samples/step,
          blur + focal + exposure + CRF + random-gamma loss + Adam.  Synthetic seeded scene, random-init nets.
value  : whole-job samples/s, step inputs already resident in HBM (device-timed with CUDA events,
          max over ranks).  A step = forward + backward + (all-reduce) + Adam.
e2e    : same metric through the public API (FusedTrainStep.step) with pinned HOST batches: the H2D copy of
          every step's inputs and a D2H read of the loss are inside the timed region.
roofline: dominant tensor-core kernel, algorithmic FLOPs / CUDA-event time (events around the kernel inside real
          eager steps) vs MEASURED_PEAKS.json's sustained peak.
cpu_baseline: the oracle port of the reference step timed on the host cores on a bounded sample; the same baseline
          leg also reports gpu_eager_baseline = that port in torch eager fp32 on the same B200 at the full workload
          size (SURVEY.md section 8d "same-box number to beat").  Both run AFTER the product measurement, on their own
          copies of the models; `--no-cpu-baseline` skips the leg.
--config me_dynamic | video_deblur | video_hdr: the uv-mapping / two-layer workloads (LayeredTrainStep), px/s.

`--impl reference` times the reference's CPU implementation of the same step: /root/reference is Python and
cannot travel to the GPU box, so this arm runs the oracle port (oracle/neucam_oracle.py, pinned to the
reference by tests/golden) with all host threads, on a bounded pixel sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_FWD_BWD = 4733952      # 3 x 788 992 MAC x 2
FLOP_PER_SAMPLE_ONE_PASS = 1577984     # forward == dgrad == wgrad algorithmic FLOPs per sample
METRIC = "scene_mlp_fwd_bwd_samples_per_s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


def make_host_batches(cfg, B, n, seed, pin):
    from neucam_b200 import harness
    scene = harness.SyntheticScene(cfg["shape"], cfg["focal_list"], seed=0)
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        inp, gt = scene.sample(B, gen)
        gamma = torch.rand(1, generator=gen) * 100 + 1.0           # training.py:181
        b = {"coords": inp["coords"].contiguous(), "coord_idx": inp["coord_idx"].contiguous(),
             "gamma_weights": inp["gamma_weights"].contiguous(), "img": gt["img"].contiguous(), "gamma": gamma}
        if pin:
            b = {k: v.pin_memory() for k, v in b.items()}
        out.append(b)
    return out


# ------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference step)
# ------------------------------------------------------------------------------------------------------
def cpu_reference_steps(cfg, B_sample, steps, warmup, threads):
    """Times `steps` reference-semantics training steps (fwd + autograd bwd + Adam) on the host."""
    from neucam_b200 import harness
    from oracle import neucam_oracle as orc
    torch.set_num_threads(threads)
    models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
    state = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    ocfg = orc.StepConfig(cfg["shape"], use_random_gamma=cfg["use_random_gamma"], with_grad_loss=False)
    batches = make_host_batches(cfg, B_sample, steps + warmup, seed=1, pin=False)
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    times = []
    for i, b in enumerate(batches):
        t0 = time.perf_counter()
        batch = {"coords": b["coords"], "coord_idx": b["coord_idx"], "img": b["img"], "gamma_weights": b["gamma_weights"]}
        _, grads = orc.step_grads(state, batch, ocfg, b["gamma"] if cfg["use_random_gamma"] else None)
        for s, d in state.items():
            for k in d:
                p, m, v = orc.adam_update(d[k], grads[s][k], mom[s][k][0], mom[s][k][1], i + 1, 1e-4)
                d[k] = p
                mom[s][k] = (m, v)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return 9 * B_sample * len(times) / total, total / len(times) * 1e3


def gpu_eager_reference_steps(cfg, B, steps, dev):
    """The oracle port of the reference step (fp32 torch ops + autograd + Adam, TF32 off) on the SAME B200 at the
    full workload size: the same-box eager baseline.  Returns (samples/s, ms/step)."""
    from neucam_b200 import harness
    from oracle import neucam_oracle as orc
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
    state = {s: {k: v.detach().clone().to(dev) for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    ocfg = orc.StepConfig(cfg["shape"], use_random_gamma=cfg["use_random_gamma"], with_grad_loss=False)
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    batches = [{k: v.to(dev) for k, v in b.items()} for b in make_host_batches(cfg, B, 2, seed=1, pin=False)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(steps + 1):
        if i == 1:
            e0.record()
        b = batches[i % 2]
        _, grads = orc.step_grads(state, b, ocfg, b["gamma"] if cfg["use_random_gamma"] else None)
        for s, d in state.items():
            for k in d:
                d[k], m, v = orc.adam_update(d[k], grads[s][k], mom[s][k][0], mom[s][k][1], i + 1, 1e-4)
                mom[s][k] = (m, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return 9 * B / (ms * 1e-3), ms


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B_sample = 4096                       # ~37k scene-MLP samples per step: about a second of CPU work
    sps, ms = cpu_reference_steps(cfg, B_sample, args.steps, max(args.warmup, 1), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "sample": f"{B_sample} of the workload's pixels per step (K=9 taps each)"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{B_sample} px/step x {args.steps} steps, oracle port of training.py:95-271 (torch CPU fp32)"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch.distributed as dist
    from neucam_b200 import harness, ops
    from neucam_b200.step import FusedTrainStep, StepOptions

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    B = harness.pixels_per_step(cfg)       # per-GPU pixels: weak scaling
    K = 9
    Q = K * B
    models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
    opts = StepOptions(patch_size=9, offset_size=5, use_random_gamma=cfg["use_random_gamma"], fixed_value=0.0, lr=1e-4)
    step = FusedTrainStep(models, cfg["shape"], opts, B, device=dev, process_group=pg)

    n_batches = min(args.steps + args.warmup, 8)
    host = make_host_batches(cfg, B, n_batches, seed=100 + rank, pin=True)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def load_resident(i):
        b = resident[i % n_batches]
        step.coords.copy_(b["coords"].view(-1, 3))
        step.coord_idx.copy_(b["coord_idx"].view(-1))
        step.img.copy_(b["img"].view(-1, 3))
        step.gamma_w.copy_(b["gamma_weights"].view(-1))
        step.gamma.copy_(b["gamma"])

    if not args.no_graph:
        load_resident(0)
        step.capture()          # forward + backward + all-reduce + Adam as ONE CUDA graph

    def one_step(i):
        load_resident(i)
        step.run_resident()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for i in range(args.warmup):
        one_step(i)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for i in range(args.steps):
            one_step(args.warmup + i)
        e1.record()
        sync_all()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * Q / (ms_step * 1e-3)
    loss_now = step.losses()["total"].item()

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H loss ----
    for i in range(2):
        step.step(host[i % n_batches], host[i % n_batches], gamma=host[i % n_batches]["gamma"])["total"].item()
    sync_all()
    e0.record()
    for i in range(args.steps):
        b = host[(args.warmup + i) % n_batches]
        losses = step.step(b, b, gamma=b["gamma"])
        _ = losses["total"].item()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    e2e_value = world * Q / (e2e_ms * 1e-3)

    # ---- per-kernel timing of the tensor-core kernels (rank 0, CUDA events on the launch stream) ----
    kern = {}
    peaks, peak_src = load_peaks()
    if rank == 0:
        # whole eager steps on the timed region's batches, the three tensor kernels bracketed by CUDA events on the
        # launch stream: durations in the context they run in (caches, clocks, neighbours), not back-to-back reruns
        evs = []
        for i in range(20):
            load_resident(i)
            step.forward_backward(events=evs)
            step.optimizer_step(reduce=False)
        torch.cuda.synchronize()
        acc = {}
        for name, a, b in evs[3 * 4:]:            # first 4 steps = warm-up of the eager path
            acc.setdefault(name, []).append(a.elapsed_time(b))
        for name, v in acc.items():
            ms = sum(v) / len(v)
            kern[name] = {"ms": ms, "launches_timed": len(v),
                          "tflops": Q * FLOP_PER_SAMPLE_ONE_PASS / (ms * 1e-3) / 1e12}
    if world > 1:
        dist.barrier()

    if rank == 0:
        dom = max(kern, key=lambda k: kern[k]["ms"])
        peak = peaks["bf16_tflops_sustained"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)   # DRAM bytes per launch from the committed ncu capture
        cpu_sps, cpu_ms = (None, None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            B_sample = 4096
            cpu_sps, cpu_ms = cpu_reference_steps(cfg, B_sample, 40, 1, threads)     # ~10 s of host work
            cpu = {"value": cpu_sps, "unit": "samples/s", "cores": threads, "kind": "port",
                   "sample": f"{B_sample} px/step (K=9 -> {9 * B_sample} samples) x 40 steps of the same workload, "
                             "oracle port of training.py:95-271 on torch CPU fp32"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": args.config, "shape_NHW": list(cfg["shape"]), "pixels_per_step_per_gpu": B,
                       "taps": K, "samples_per_step_per_gpu": Q, "random_gamma": cfg["use_random_gamma"],
                       "l2": "per-step working set (4.2 GB of fp16 activations) >> 126 MB L2; no explicit flush",
                       "parallelism": f"dp{world}: pixel batch sharded, one NCCL all-reduce of the flat gradient",
                       "cuda_graph": not args.no_graph},
            "px_per_s": value / K,
            "tflops_algorithmic": value * FLOP_PER_SAMPLE_FWD_BWD / 1e12,
            "frac_of_tensor_peak_whole_step": value / world * FLOP_PER_SAMPLE_FWD_BWD / 1e12 / peaks["bf16_tflops_sustained"],
            "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": step.h2d_bytes(), "d2h_bytes_per_step": 4},
            "gpu_launches": 12 * args.steps,
            "clocks": clk.summary(),
            "roofline": {"bound": "tensor", "kernel": dom, "achieved": kern[dom]["tflops"], "peak": peak,
                         "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / peak, "traffic": traffic,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside full steps), {peak_src}",
                         "flops_per_launch": Q * FLOP_PER_SAMPLE_ONE_PASS,
                         "traffic_unit": "DRAM bytes per launch (ncu), vs the algorithmic "
                                         "7 KB/sample forward, 6 KB/sample backward, 6 KB/sample wgrad"},
            "kernels": kern,
            "loss": loss_now,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_cpu_baseline:
            del step, resident
            torch.cuda.empty_cache()
            sps, ms = gpu_eager_reference_steps(cfg, B, 3, dev)
            line["gpu_eager_baseline"] = {"value": sps, "unit": "samples/s", "ms_per_step": ms, "kind": "port",
                                          "what": "oracle port of training.py:95-271 (torch eager fp32, TF32 off, autograd "
                                                  "+ Adam) on the same B200 at the full workload size, 3 steps"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# the uv-mapping / two-layer workloads (SURVEY.md section 8f "next" rows): --config me_dynamic|video_deblur|video_hdr
# ------------------------------------------------------------------------------------------------------
def layered_host_batches(kind, B, n, seed, pin):
    from neucam_b200 import harness
    spec = harness.LAYERED_CONFIGS[kind]
    N, H, W = spec["shape"]
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        idx = torch.randint(0, N * H * W, (B,), generator=gen)
        t, rem = idx // (H * W), idx % (H * W)
        tyx = torch.stack((t / max(N - 1, 1), (rem // W) / (H - 1), (rem % W) / (W - 1)), -1).float() * 2 - 1
        b = {"coords": tyx[None].contiguous(), "coord_idx": idx[None].contiguous(),
             "gamma_weights": torch.ones(1, B, 1), "img": torch.rand(1, B, 3, generator=gen) * 2 - 1,
             "gamma": torch.rand(1, generator=gen) * 100 + 1.0}
        out.append({k: v.pin_memory() for k, v in b.items()} if pin else b)
    return out


def layered_cpu_steps(kind, B_sample, steps, warmup, threads):
    from neucam_b200 import harness
    from oracle import neucam_oracle as orc
    torch.set_num_threads(threads)
    spec = harness.LAYERED_CONFIGS[kind]
    models = harness.build_layered_models(kind, seed=0)
    state = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    nets = {s: dict(last=spec[s][3], skip=(4,), num_frequencies=spec[s][2] if spec[s][1] else None)
            for s in ("uv_b", "uv_f", "alpha") if s in spec}
    o = spec["options"]
    ocfg = orc.StepConfig(spec["shape"], use_random_gamma=spec["use_random_gamma"], with_grad_loss=False,
                          uv_mapping_scale=o["uv_mapping_scale"], rigidity_loss_w=o.get("rigidity_loss_w"),
                          sparsity_loss_w=o.get("sparsity_loss_w"), alpha_loss_w=o.get("alpha_loss_w"), nets=nets)
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    times = []
    for i, b in enumerate(layered_host_batches(kind, B_sample, steps + warmup, 1, False)):
        t0 = time.perf_counter()
        _, grads = orc.step_grads(state, b, ocfg, b["gamma"] if spec["use_random_gamma"] else None)
        for s, d in state.items():
            for k in d:
                p, m, v = orc.adam_update(d[k], grads[s][k], mom[s][k][0], mom[s][k][1], i + 1, 1e-4)
                d[k], mom[s][k] = p, (m, v)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return B_sample * len(times) / sum(times), sum(times) / len(times) * 1e3


def run_layered(args, kind):
    import torch.distributed as dist
    from neucam_b200 import harness
    from neucam_b200.layered import LayeredTrainStep
    from neucam_b200.step import StepOptions

    spec = harness.LAYERED_CONFIGS[kind]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            threads = os.cpu_count() or 1
            B_sample = 2048
            pps, ms = layered_cpu_steps(kind, B_sample, args.steps, max(args.warmup, 1), threads)
            print(json.dumps({"impl": "reference", "metric": "train_pixels_per_second", "value": pps, "unit": "px/s",
                              "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": kind, "sample": f"{B_sample} px/step"},
                              "cpu_baseline": {"value": pps, "unit": "px/s", "cores": threads, "kind": "port",
                                               "sample": f"{B_sample} px/step x {args.steps} steps, oracle port"},
                              "e2e": {"value": pps, "unit": "px/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    N, H, W = spec["shape"]
    B = int(spec["sample_frac"] * N * H * W)
    K = 9 if spec["deblur"] else 1
    atlases = 2 if "uv_f" in spec else 1
    models = harness.build_layered_models(kind, seed=0)
    opts = StepOptions(patch_size=9, offset_size=5, use_random_gamma=spec["use_random_gamma"], lr=1e-4, **spec["options"])
    step = LayeredTrainStep(models, spec["shape"], opts, device=dev, process_group=pg)
    n_batches = min(args.steps + args.warmup, 8)
    host = layered_host_batches(kind, B, n_batches, 100 + rank, True)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    use_gamma = spec["use_random_gamma"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if not args.no_graph:
        step.capture(resident[0], resident[0], resident[0]["gamma"] if use_gamma else None)

    def one_step(i, src):
        b = src[i % n_batches]
        if args.no_graph:
            return step.step(b, b, b["gamma"] if use_gamma else None)
        step.load_resident(b, b, b["gamma"] if use_gamma else None)
        return step.run_resident()

    for i in range(args.warmup):
        one_step(i, resident)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for i in range(args.steps):
            one_step(args.warmup + i, resident)
        e1.record()
        sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    loss_now = step.losses()["total"].item()
    # end to end: pinned host batch -> H2D -> step -> loss D2H
    for i in range(2):
        one_step(i, host)["total"].item()
    sync_all()
    e0.record()
    for i in range(args.steps):
        _ = one_step(args.warmup + i, host)["total"].item()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for k, v in host[0].items() if k != "gamma" or use_gamma)
        line = {"metric": "train_pixels_per_second", "value": world * B / (ms_step * 1e-3), "unit": "px/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 scene nets / f32 glue",
                "data": "synthetic",
                "config": {"workload": kind, "shape_NHW": [N, H, W], "pixels_per_step_per_gpu": B, "taps": K,
                           "scene_networks": atlases, "scene_samples_per_step_per_gpu": K * B * atlases,
                           "cuda_graph": not args.no_graph},
                "scene_samples_per_s": world * K * B * atlases / (ms_step * 1e-3),
                "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "px/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "clocks": clk.summary(), "loss": loss_now}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            pps, _ = layered_cpu_steps(kind, 2048, 10, 1, threads)
            line["cpu_baseline"] = {"value": pps, "unit": "px/s", "cores": threads, "kind": "port",
                                    "sample": "2048 px/step x 10 steps, oracle port on torch CPU fp32"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mfme_blender")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from neucam_b200 import harness
    if args.config in harness.LAYERED_CONFIGS:
        return run_layered(args, args.config)
    cfg = harness.CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
