"""Kernel-by-kernel timeline of ONE graph-replayed step (static or layered config) (CUPTI through torch.profiler): start offset,
duration and the gap to the previous kernel's end -- where the step's wall time goes beyond the sum of its kernels."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions
from torch.profiler import ProfilerActivity, profile

name = sys.argv[1] if len(sys.argv) > 1 else "mfme_blender"
if name in harness.LAYERED_CONFIGS:
    import bench
    from neucam_b200.layered import LayeredTrainStep
    spec = harness.LAYERED_CONFIGS[name]
    N, H, W = spec["shape"]
    B = int(spec["sample_frac"] * N * H * W)
    models = harness.build_layered_models(name, seed=0)
    opts = StepOptions(use_random_gamma=spec["use_random_gamma"], lr=1e-4, **spec["options"])
    step = LayeredTrainStep(models, spec["shape"], opts, device="cuda")
    b = {k: v.cuda() for k, v in bench.layered_host_batches(name, B, 1, 0, False)[0].items()}
    step.capture(b, b, b["gamma"] if spec["use_random_gamma"] else None)
else:
    cfg = harness.CONFIGS[name]
    B = harness.pixels_per_step(cfg)
    models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
    step = FusedTrainStep(models, cfg["shape"], StepOptions(use_random_gamma=cfg["use_random_gamma"]), B, device="cuda")
    scene = harness.SyntheticScene(cfg["shape"], cfg["focal_list"], seed=0)
    inp, gt = scene.sample(B, torch.Generator().manual_seed(1))
    step.load_batch(inp, gt, torch.tensor([30.0]))
    step.capture()
for _ in range(5):
    step.run_resident()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step.run_resident()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
# take the middle replay: every replay launches the same kernels, so it is the middle third of the events
n = len(evs) // 3
seg = evs[n:2 * n] if len(evs) % 3 == 0 else evs
t0 = seg[0].time_range.start
end_prev = t0
print(f"{'start us':>9} {'dur us':>8} {'gap us':>7}  kernel")
tot = 0.0
for e in seg:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:9.1f} {d:8.1f} {e.time_range.start - end_prev:7.1f}  {e.name[:90]}")
    end_prev = max(end_prev, e.time_range.end)
    tot += d
print(f"step span {end_prev - t0:.1f} us, sum of kernel durations {tot:.1f} us (side-stream kernels overlap)")
