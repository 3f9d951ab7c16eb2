#!/usr/bin/env python
"""bench.py -- headline benchmark of the NeuCam hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]
                    [--scaling weak|strong] [--only-headline]

metric : scene-MLP fwd+bwd samples/s (1 sample = one scene-MLP query = one of the K taps of a pixel)
workload: BASELINE.json configs[1] -- config_mfme_blender.txt in its 9-image 3-focus x 3-exposure form:
          N,H,W = 9,300,450, sample_frac 3.5e-2 -> B = 42 525 px/step, K = 9 taps, Q = 382 725 samples/step,
          blur + focal + exposure + CRF + random-gamma loss + Adam.  Synthetic seeded scene, random-init nets.
value  : whole-job samples/s, step inputs already resident in HBM (device-timed with CUDA events, max over ranks): the
          video is attached to the step (attach_dataset) and every step's batch is drawn by the on-device sampler.
          A step = sampler + forward + backward + (all-reduce) + Adam, the latter four replayed from ONE CUDA graph.
e2e    : same metric through the public API (FusedTrainStep.step) with pinned HOST batches: the H2D copy of every
          step's inputs and a D2H read of the loss are inside the timed region.
roofline: the slowest tensor-core kernel, algorithmic FLOPs / CUDA-event time (events around the kernel inside real
          eager steps on the launch stream) vs MEASURED_PEAKS.json: the BURST peak when the sampled SM clock sat at its
          maximum with no power cap, the SUSTAINED peak when sw_power_cap was sampled -- `peak_source` says which.
cpu_baseline: the oracle port of the reference step timed on the host cores on a bounded sample; the same leg also
          reports gpu_eager_baseline = that port in torch eager fp32 on the same B200 at the full workload size (SURVEY.md
          section 8d "same-box number to beat") and parity_at_scale = the fused step's gradients against that port's
          gradients on the SAME full-size batch (the line is marked rejected above 1e-2).
configs: ONE JSON line is printed (the driver's contract); the other four shipped configs (mf_static and the three
          uv-mapping / two-layer ones) are measured the same way and nested under "configs" in that line
          (`--only-headline` skips them, `--config X` makes X the line).
strong : the line also carries "strong_scaling": the native global batch of the headline and of mf_static split over the
          N ranks, next to the weak-scaling `value`.  `--scaling strong` makes that the line's value.

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
ADAM_BYTES_PER_PARAM = 28              # read p, g, m, v; write p, m, v


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json"
    # /opt/skills/guides/B200_PROFILING.md fallback figures
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "B200_PROFILING.md fallback"


def pick_peak(peaks, src, clocks):
    """Burst peak for a kernel that ran with the SM clock at its maximum and no power cap; sustained otherwise."""
    capped = "sw_power_cap" in (clocks.get("reasons") or [])
    sm, mx = clocks.get("sm_mhz"), clocks.get("sm_max_mhz")
    at_max = sm is not None and mx is not None and sm >= 0.97 * mx
    if at_max and not capped:
        return peaks["bf16_tflops"], f"{src} bf16_tflops (burst: SM clock {sm:.0f} of {mx:.0f} MHz, no power cap sampled)"
    why = "sw_power_cap sampled" if capped else f"SM clock {sm} of {mx} MHz under load"
    return peaks["bf16_tflops_sustained"], f"{src} bf16_tflops_sustained ({why})"


class ClockSampler:
    """SM clock, power and throttle reasons DURING a timed region, polled through NVML every 2 ms (nvidia-smi's process
    start-up is slower than a whole 50 ms timed region; NVML is what it reads anyway).  The power sensor is a ~100 ms
    moving average: on a region shorter than that, `power_w_max` lags the real draw."""

    def __init__(self, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        nv = self.nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": mx, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= int(r[2])
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
        reasons = [n for n, attr in names if bits & int(getattr(nv, attr, 0))]
        return {"sm_mhz": sm[len(sm) // 2], "sm_mhz_min": sm[0], "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(r[1] for r in self.rows)}


class Dist:
    """The process's data-parallel context (one rank per GPU, NCCL)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.pg = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.pg = dist.group.WORLD

    def sync(self):
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def max_ms(self, ms):
        t = torch.tensor([ms], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def checksums_equal(self, c):
        """True iff every rank holds the same parameter checksum (bit-identical replicas)."""
        if self.world == 1:
            return True
        import torch.distributed as dist
        lo, hi = c.clone(), c.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool((lo == hi).item())

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


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
# CPU arm (oracle port of the reference step) and the same-box eager baseline
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


def gpu_eager_and_parity(cfg, B, steps, dev, fused_grads, state0, batch0):
    """(1) The oracle port of the reference step (fp32 torch ops + autograd + Adam, TF32 off) on the SAME B200 at the
    full workload size: the same-box eager baseline.  (2) parity at scale: its gradients on
    `batch0` from parameters `state0` against `fused_grads` (what FusedTrainStep left in .grad for the same batch and
    parameters): relative L2 error per tensor, bound 1e-2 (BASELINE.json north_star)."""
    from oracle import neucam_oracle as orc
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ocfg = orc.StepConfig(cfg["shape"], use_random_gamma=cfg["use_random_gamma"], with_grad_loss=False)
    _, ref_g = orc.step_grads(state0, batch0, ocfg, batch0["gamma"] if cfg["use_random_gamma"] else None)
    worst, worst_name, n_t = 0.0, "", 0
    for slot, d in ref_g.items():
        for k, g in d.items():
            if g.abs().max() == 0:
                continue
            e = ((fused_grads[slot][k].double() - g.double()).norm() / g.double().norm()).item()
            n_t += 1
            if e > worst:
                worst, worst_name = e, f"{slot}.{k}"
    parity = {"what": "relative L2 error of every parameter gradient of the fused step vs the oracle port (torch eager fp32, "
                      "TF32 off) on the same full-size batch and parameters",
              "samples": 9 * B, "tensors": n_t, "worst": worst, "worst_tensor": worst_name, "bound": 1e-2,
              "ok": bool(worst <= 1e-2)}
    del ref_g
    state = {s: {k: v.clone() for k, v in d.items()} for s, d in state0.items()}
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(steps + 1):
        if i == 1:
            e0.record()
        _, grads = orc.step_grads(state, batch0, ocfg, batch0["gamma"] if cfg["use_random_gamma"] else None)
        for s, d in state.items():
            for k in d:
                d[k], m, v = orc.adam_update(d[k], grads[s][k], mom[s][k][0], mom[s][k][1], i + 1, 1e-4)
                mom[s][k] = (m, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eager = {"value": 9 * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "kind": "port",
             "what": f"oracle port of training.py:95-271 (torch eager fp32, TF32 off, autograd + Adam) on the same B200 at "
                     f"the full workload size, {steps} steps"}
    return eager, parity


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
# static multi-focus / multi-exposure workloads (FusedTrainStep): mfme_blender (headline), mf_static
# ------------------------------------------------------------------------------------------------------
def measure_static(D, cfg, B, global_pixels, steps, warmup, with_kernels=True, keep_for_parity=False):
    """One measurement of the static step at B pixels per rank.  Returns a dict (rank 0 fills everything)."""
    from neucam_b200 import harness, ops
    from neucam_b200.step import FusedTrainStep, StepOptions

    K = 9
    Q = K * B
    dev = D.dev
    models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
    opts = StepOptions(patch_size=9, offset_size=5, use_random_gamma=cfg["use_random_gamma"], fixed_value=0.0, lr=1e-4)
    step = FusedTrainStep(models, cfg["shape"], opts, B, device=dev,
                          process_group=None if os.environ.get("NEUCAM_BENCH_NO_COMM") else D.pg,   # diagnostics only
                          global_pixels=global_pixels, overlap_allreduce=not os.environ.get("NEUCAM_BENCH_NO_OVERLAP"))
    n_batches = min(steps + warmup, 8)
    host = make_host_batches(cfg, B, n_batches, seed=100 + D.rank, pin=True)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def load_resident(i):
        b = resident[i % n_batches]
        step.coords.copy_(b["coords"].view(-1, 3))
        step.coord_idx.copy_(b["coord_idx"].view(-1))
        step.img.copy_(b["img"].view(-1, 3))
        step.gamma_w.copy_(b["gamma_weights"].view(-1))
        step.gamma.copy_(b["gamma"])

    parity_inputs = None
    if keep_for_parity and D.rank == 0 and D.world == 1:
        # gradients of the very first step (initial parameters, batch 0) for parity_at_scale
        load_resident(0)
        state0 = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
        step.forward_backward()
        torch.cuda.synchronize()
        fused = {s: {k: p.grad.detach().clone() for k, p in m.named_parameters()} for s, m in models.items() if m is not None}
        b0 = resident[0]
        batch0 = {"coords": b0["coords"], "coord_idx": b0["coord_idx"], "img": b0["img"],
                  "gamma_weights": b0["gamma_weights"], "gamma": b0["gamma"]}
        parity_inputs = (fused, state0, batch0)

    n0 = ops.launch_count()
    load_resident(0)
    step.capture()          # forward + backward + all-reduce + Adam + re-pack as ONE CUDA graph (2 warm-up steps + capture)
    launches_per_step = (ops.launch_count() - n0) // 3

    # device-resident loop = the product's resident-data path: the video lives in HBM (attach_dataset) and every step's
    # batch is drawn there by the on-device sampler (neucam_sample_batch, one launch; dataio.py:356-393 semantics),
    # plus the step's random gamma (4 bytes, device to device)
    scene = harness.SyntheticScene(cfg["shape"], cfg["focal_list"], seed=0)
    step.attach_dataset(scene.data, gamma_weights=scene.gamma_weights if cfg["use_random_gamma"] else None)

    def one_step(i):
        step.sample_resident(seed=100 + D.rank, step=i, gamma=resident[i % n_batches]["gamma"])
        step.run_resident()

    for i in range(warmup):
        one_step(i)
    D.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(D.local_rank) as clk:
        e0.record()
        for i in range(steps):
            one_step(warmup + i)
        e1.record()
        D.sync()
    ms_step = D.max_ms(e0.elapsed_time(e1)) / steps
    loss_now = step.losses()["total"].item()
    same = D.checksums_equal(step.params.checksum())

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H loss ----
    for i in range(2):
        step.step(host[i % n_batches], host[i % n_batches], gamma=host[i % n_batches]["gamma"])["total"].item()
    D.sync()
    e0.record()
    for i in range(steps):
        b = host[(warmup + i) % n_batches]
        _ = step.step(b, b, gamma=b["gamma"])["total"].item()
    e1.record()
    D.sync()
    e2e_ms = D.max_ms(e0.elapsed_time(e1)) / steps

    out = {"B": B, "Q": Q, "ms_per_step": ms_step, "e2e_ms": e2e_ms, "loss": loss_now, "clocks": clk.summary(),
           "launches_per_step": launches_per_step, "h2d": step.h2d_bytes(), "replicas_identical": same,
           "params": step.params.numel, "parity_inputs": parity_inputs}

    # ---- per-kernel timing (rank 0): whole eager steps, the tensor kernels bracketed by CUDA events on the launch
    #      stream -- durations in the context they run in (caches, clocks, neighbours), not back-to-back reruns ----
    kern = {}
    kern_clocks = None
    if with_kernels and D.rank == 0:
        evs = []
        n_eager = 16
        torch.cuda.synchronize()
        time.sleep(0.5)      # let the power-management state of the previous (longer) regions decay: the kernels are
                             # timed in a burst of the same length as the driver's 20-step timed region
        with ClockSampler(D.local_rank) as kclk:
            for i in range(n_eager):
                load_resident(i)
                step.forward_backward(events=evs)
                step.optimizer_step(reduce=False)
            torch.cuda.synchronize()
        kern_clocks = kclk.summary()
        acc = {}
        for name, a, b in evs[3 * 4:]:            # first 4 steps = warm-up of the eager path
            acc.setdefault(name, []).append(a.elapsed_time(b))
        for name, v in acc.items():
            ms = sum(v) / len(v)
            kern[name] = {"ms": ms, "launches_timed": len(v), "tflops": Q * FLOP_PER_SAMPLE_ONE_PASS / (ms * 1e-3) / 1e12}
        # the HBM-bound Adam kernel alone: 28 B / parameter
        P = step.params
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.adam_step(P.flat, P.grad, P.exp_avg, P.exp_avg_sq, P.adam_state, 0.0)
        ea.record()
        for _ in range(20):
            ops.adam_step(P.flat, P.grad, P.exp_avg, P.exp_avg_sq, P.adam_state, 0.0)
        eb.record()
        torch.cuda.synchronize()
        ms = ea.elapsed_time(eb) / 20
        kern["adam"] = {"ms": ms, "gbs": P.numel * ADAM_BYTES_PER_PARAM / (ms * 1e-3) / 1e9,
                        "bytes": P.numel * ADAM_BYTES_PER_PARAM,
                        "note": "3.2 MB of parameters + state: L2-resident, launch-latency bound, not HBM-bound"}
    if D.world > 1:
        import torch.distributed as dist
        dist.barrier()
    out["kernels"] = kern
    out["kernel_clocks"] = kern_clocks
    del step
    torch.cuda.empty_cache()
    return out


def static_line(D, args, name, cfg, peaks, peak_src, B, global_pixels, scaling, with_baselines):
    m = measure_static(D, cfg, B, global_pixels, args.steps, args.warmup, with_kernels=True,
                       keep_for_parity=with_baselines)
    if D.rank != 0:
        return None
    K, Q, world = 9, m["Q"], D.world
    value = world * Q / (m["ms_per_step"] * 1e-3)
    e2e_value = world * Q / (m["e2e_ms"] * 1e-3)
    kern = m["kernels"]
    tensor_k = {k: v for k, v in kern.items() if "tflops" in v}
    dom = max(tensor_k, key=lambda k: tensor_k[k]["ms"])
    # the roofline kernel is timed in its own region: the peak it is compared with follows THAT region's clocks
    peak, peak_source = pick_peak(peaks, peak_src, m["kernel_clocks"] or m["clocks"])
    step_peak, _ = pick_peak(peaks, peak_src, m["clocks"])
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and name == "mfme_blender":
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get(dom), tj.get("source")
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": name, "shape_NHW": list(cfg["shape"]), "pixels_per_step_per_gpu": B,
                   "pixels_per_step_global": global_pixels, "taps": K, "samples_per_step_per_gpu": Q,
                   "random_gamma": cfg["use_random_gamma"],
                   "l2": f"per-step working set ({11 * Q / 1e6:.0f} MB of fp16 activations / phases / dz) vs 126 MB L2: "
                         + ("inputs larger than L2, no explicit flush" if 11e3 * Q > 2 * 126e6 else
                            "NOT larger than L2 -- this small config is partly L2-resident by nature"),
                   "parallelism": f"dp{world}: pixel batch sharded, NCCL all-reduce of the flat gradient captured in the "
                                  "step graph, atlas bucket overlapped with the backward's tails",
                   "cuda_graph": True,
                   "resident_input": "video in HBM, batch drawn per step by neucam_sample_batch (on-device sampler)"},
        "px_per_s": value / K,
        "tflops_algorithmic": value * FLOP_PER_SAMPLE_FWD_BWD / 1e12,
        "frac_of_tensor_peak_whole_step": value / world * FLOP_PER_SAMPLE_FWD_BWD / 1e12 / step_peak,
        "e2e": {"value": e2e_value, "unit": "samples/s", "ms_per_step": m["e2e_ms"],
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4},
        "gpu_launches": m["launches_per_step"] * args.steps,
        "gpu_launches_per_step": m["launches_per_step"],
        "gpu_launches_how": "neucam_launch_count() (the library counts its own kernel launches) over the graph's "
                            "warm-up + capture, per step, x steps replayed",
        "clocks": m["clocks"],
        "kernel_timing_clocks": m["kernel_clocks"],
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": tensor_k[dom]["tflops"], "peak": peak,
                     "unit": "TFLOP/s", "frac": tensor_k[dom]["tflops"] / peak, "traffic": traffic,
                     "peak_source": peak_source, "flops_per_launch": Q * FLOP_PER_SAMPLE_ONE_PASS,
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_sample": "input 8 B (uv) + parameters; the kernels additionally stream fp16 "
                                                     "scratch for the backward: forward 7 KB, backward 6 KB, wgrad 6 KB "
                                                     "per sample (DESIGN.md section 3)"},
        "kernels": kern,
        "replicas_bit_identical": m["replicas_identical"],
        "loss": m["loss"],
    }
    if with_baselines and world == 1:
        threads = os.cpu_count() or 1
        B_sample = 4096
        cpu_sps, _ = cpu_reference_steps(cfg, B_sample, 40 if name == "mfme_blender" else 15, 1, threads)
        line["cpu_baseline"] = {"value": cpu_sps, "unit": "samples/s", "cores": threads, "kind": "port",
                                "sample": f"{B_sample} px/step (K=9 -> {9 * B_sample} samples) of the same workload, "
                                          "oracle port of training.py:95-271 on torch CPU fp32"}
        fused, state0, batch0 = m["parity_inputs"]
        eager, parity = gpu_eager_and_parity(cfg, B, 3, D.dev, fused, state0, batch0)
        line["gpu_eager_baseline"] = eager
        line["parity_at_scale"] = parity
        if not parity["ok"]:
            line["rejected"] = "parity_at_scale above 1e-2: this line's performance numbers do not count"
    return line


# ------------------------------------------------------------------------------------------------------
# the uv-mapping / two-layer workloads (SURVEY.md section 8f "next" rows): me_dynamic | video_deblur | video_hdr
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


def layered_flops_per_step(kind, B):
    """Algorithmic FLOPs of one step: scene networks 4 733 952 per sample (fwd + bwd); uv /
    alpha SimpleMLPs 6 x MACs per evaluated point (K B taps, + 2 B shifted pixels per uv network for the rigidity
    loss); split-precision passes and recomputation are NOT counted."""
    from neucam_b200 import harness
    spec = harness.LAYERED_CONFIGS[kind]
    K = 9 if spec["deblur"] else 1
    atlases = 2 if "uv_f" in spec else 1
    total = K * B * atlases * FLOP_PER_SAMPLE_FWD_BWD

    def macs(hp, out_dim):
        layers, enc, freqs, _ = hp
        e = 3 + 6 * freqs if enc else 3
        dims = [e] + [256] * (layers - 1) + [out_dim]
        m = sum((dims[i] + (e if i == 4 else 0)) * dims[i + 1] for i in range(layers))
        return m

    rigid = spec["options"].get("rigidity_loss_w") is not None
    for slot, od in (("uv_b", 2), ("uv_f", 2), ("alpha", 1)):
        if slot in spec:
            pts = K * B + (2 * B if (rigid and slot != "alpha") else 0)
            total += 6 * macs(spec[slot], od) * pts
    return total


def measure_layered(D, args, kind, peaks, peak_src, with_baselines):
    from neucam_b200 import harness, ops
    from neucam_b200.layered import LayeredTrainStep
    from neucam_b200.step import StepOptions

    spec = harness.LAYERED_CONFIGS[kind]
    rank, world, dev = D.rank, D.world, D.dev
    N, H, W = spec["shape"]
    B = int(spec["sample_frac"] * N * H * W)
    K = 9 if spec["deblur"] else 1
    atlases = 2 if "uv_f" in spec else 1
    models = harness.build_layered_models(kind, seed=0)
    opts = StepOptions(patch_size=9, offset_size=5, use_random_gamma=spec["use_random_gamma"], lr=1e-4, **spec["options"])
    step = LayeredTrainStep(models, spec["shape"], opts, device=dev, process_group=D.pg)
    n_batches = min(args.steps + args.warmup, 8)
    host = layered_host_batches(kind, B, n_batches, 100 + rank, True)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    use_gamma = spec["use_random_gamma"]

    # device-resident loop = the product's resident-data path: a (synthetic, uniform-random) video lives in HBM, every
    # step's batch is drawn there by the on-device sampler straight into the graph's static buffers (capture(adopt=True))
    gvid = torch.Generator(device=dev).manual_seed(7 + rank)
    video = torch.rand(N * H * W, 3, device=dev, generator=gvid) * 2 - 1
    step.attach_dataset(video, B, gamma_weights=torch.ones(N * H * W, device=dev) if use_gamma else None)
    gamma_dev = resident[0]["gamma"].clone() if use_gamma else None
    mi, gt_res = step.sample_resident(seed=100 + rank, step=0)
    n0 = ops.launch_count()
    step.capture(mi, gt_res, gamma_dev, adopt=True)
    launches_per_step = (ops.launch_count() - n0) // 3 + 1      # + the sampler

    def one_step_resident(i):
        step.sample_resident(seed=100 + rank, step=i)
        if use_gamma:
            gamma_dev.copy_(resident[i % n_batches]["gamma"])
        return step.run_resident()

    def one_step(i, src):
        b = src[i % n_batches]
        step.load_resident(b, b, b["gamma"] if use_gamma else None)
        return step.run_resident()

    for i in range(args.warmup):
        one_step_resident(i)
    D.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(D.local_rank) as clk:
        e0.record()
        for i in range(args.steps):
            one_step_resident(args.warmup + i)
        e1.record()
        D.sync()
    ms_step = D.max_ms(e0.elapsed_time(e1)) / args.steps
    loss_now = step.losses()["total"].item()
    same = D.checksums_equal(step.params.checksum())
    for i in range(2):
        one_step(i, host)["total"].item()
    D.sync()
    e0.record()
    for i in range(args.steps):
        _ = one_step(args.warmup + i, host)["total"].item()
    e1.record()
    D.sync()
    e2e_ms = D.max_ms(e0.elapsed_time(e1)) / args.steps
    line = None
    if rank == 0:
        h2d = sum(v.numel() * v.element_size() for k, v in host[0].items() if k != "gamma" or use_gamma)
        clocks = clk.summary()
        peak, peak_source = pick_peak(peaks, peak_src, clocks)
        flops = layered_flops_per_step(kind, B)
        tf = flops / (ms_step * 1e-3) / 1e12
        line = {"metric": "train_pixels_per_second", "value": world * B / (ms_step * 1e-3), "unit": "px/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 (split hi+lo for the ReLU networks) / f32 accumulate", "data": "synthetic",
                "config": {"workload": kind, "shape_NHW": [N, H, W], "pixels_per_step_per_gpu": B, "taps": K,
                           "scene_networks": atlases, "scene_samples_per_step_per_gpu": K * B * atlases,
                           "cuda_graph": True,
                           "resident_input": "video in HBM, batch drawn per step by neucam_sample_batch into the graph's "
                                             "static buffers"},
                "scene_samples_per_s": world * K * B * atlases / (ms_step * 1e-3),
                "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "px/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                "gpu_launches_how": "neucam_launch_count() over warm-up + capture; torch glue kernels are NOT counted",
                "roofline": {"bound": "tensor", "kernel": "whole step (every kernel of the graph)", "achieved": tf,
                             "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                             "peak_source": peak_source, "flops_per_step": flops,
                             "note": "algorithmic FLOPs (SURVEY.md section 8d): hi/lo split passes of the ReLU chains "
                                     "are not counted"},
                "clocks": clocks, "replicas_bit_identical": same, "loss": loss_now}
        if with_baselines and world == 1:
            threads = os.cpu_count() or 1
            pps, _ = layered_cpu_steps(kind, 2048, 6, 1, threads)
            line["cpu_baseline"] = {"value": pps, "unit": "px/s", "cores": threads, "kind": "port",
                                    "sample": "2048 px/step x 6 steps, oracle port on torch CPU fp32"}
    del step
    torch.cuda.empty_cache()
    return line


def layered_reference_arm(args, kind):
    if int(os.environ.get("RANK", "0")) != 0:
        return
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


# ------------------------------------------------------------------------------------------------------
def strong_scaling(D, args, peaks, peak_src):
    """The native global batch split over the N ranks: headline and mf_static."""
    from neucam_b200 import harness
    from neucam_b200.dist import shard_range
    out = {}
    for name in ("mfme_blender", "mf_static"):
        cfg = harness.CONFIGS[name]
        Bg = harness.pixels_per_step(cfg)
        lo, hi = shard_range(Bg, D.rank, D.world)
        m = measure_static(D, cfg, hi - lo, Bg, args.steps, args.warmup, with_kernels=False)
        if D.rank == 0:
            tiles = -(-9 * (hi - lo) // 128)
            out[name] = {"value": 9 * Bg / (m["ms_per_step"] * 1e-3), "unit": "samples/s", "ms_per_step": m["ms_per_step"],
                         "pixels_global": Bg, "pixels_rank0": hi - lo, "tiles_rank0": tiles,
                         "waves_of_148_sm": tiles / 148.0, "replicas_bit_identical": m["replicas_identical"]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mfme_blender")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the nested measurements of the other configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from neucam_b200 import harness
    layered = args.config in harness.LAYERED_CONFIGS
    if args.impl == "reference":
        if layered:
            return layered_reference_arm(args, args.config)
        return run_reference_arm(args, harness.CONFIGS[args.config])

    D = Dist()
    peaks, peak_src = load_peaks()
    with_baselines = not args.no_cpu_baseline
    if layered:
        line = measure_layered(D, args, args.config, peaks, peak_src, with_baselines)
    else:
        cfg = harness.CONFIGS[args.config]
        Bg = harness.pixels_per_step(cfg)
        if args.scaling == "strong":
            from neucam_b200.dist import shard_range
            lo, hi = shard_range(Bg, D.rank, D.world)
            line = static_line(D, args, args.config, cfg, peaks, peak_src, hi - lo, Bg, "strong", with_baselines)
        else:
            line = static_line(D, args, args.config, cfg, peaks, peak_src, Bg, Bg * D.world, "weak", with_baselines)
    default_run = args.config == "mfme_blender" and args.scaling == "weak" and not args.only_headline
    if default_run:
        # the other shipped configs, measured the same way, nested in the one line the driver parses
        nested = {}
        mf = static_line(D, args, "mf_static", harness.CONFIGS["mf_static"], peaks, peak_src,
                         harness.pixels_per_step(harness.CONFIGS["mf_static"]),
                         harness.pixels_per_step(harness.CONFIGS["mf_static"]) * D.world, "weak", with_baselines)
        if D.rank == 0:
            nested["mf_static"] = mf
        if D.world == 1:
            for kind in ("me_dynamic", "video_deblur", "video_hdr"):
                nested[kind] = measure_layered(D, args, kind, peaks, peak_src, with_baselines)
        strong = strong_scaling(D, args, peaks, peak_src) if D.world > 1 else None
        if D.rank == 0:
            line["configs"] = nested
            if strong is not None:
                line["strong_scaling"] = strong
    if D.rank == 0:
        print(json.dumps(line))
    D.close()
    if D.rank == 0 and line.get("rejected"):
        sys.exit(3)


if __name__ == "__main__":
    main()
