"""aten-level op counts of one eager LayeredTrainStep step (what torch glue is left).  python tools/profile_layered_ops.py video_deblur"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from neucam_b200 import harness
from neucam_b200.layered import LayeredTrainStep
from neucam_b200.step import StepOptions
from torch.profiler import ProfilerActivity, profile
kind = sys.argv[1] if len(sys.argv) > 1 else "video_deblur"
spec = harness.LAYERED_CONFIGS[kind]
N, H, W = spec["shape"]
B = int(spec["sample_frac"] * N * H * W)
models = harness.build_layered_models(kind, seed=0)
step = LayeredTrainStep(models, spec["shape"], StepOptions(use_random_gamma=spec["use_random_gamma"], lr=1e-4, **spec["options"]), device="cuda")
b = {k: v.cuda() for k, v in bench.layered_host_batches(kind, B, 1, 0, False)[0].items()}
gamma = b["gamma"] if spec["use_random_gamma"] else None
for _ in range(3):
    step.step(b, b, gamma)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step.step(b, b, gamma)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.self_device_time_total) for e in prof.key_averages() if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"aten ops with device time: {sum(r[1] for r in rows)} calls, {tot:.0f} us")
for k, n, t in rows[:22]:
    print(f"{k:32s} {n:4d} calls {t:8.1f} us")
