"""Kernel-time table of one eager LayeredTrainStep step (torch.profiler, CUDA activities only).
    python tools/profile_layered.py video_hdr"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from neucam_b200 import harness  # noqa: E402
from neucam_b200.layered import LayeredTrainStep  # noqa: E402
from neucam_b200.step import StepOptions  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "video_hdr"
spec = harness.LAYERED_CONFIGS[kind]
N, H, W = spec["shape"]
B = int(spec["sample_frac"] * N * H * W)
models = harness.build_layered_models(kind, seed=0)
opts = StepOptions(use_random_gamma=spec["use_random_gamma"], lr=1e-4, **spec["options"])
step = LayeredTrainStep(models, spec["shape"], opts, device="cuda")
b = {k: v.cuda() for k, v in bench.layered_host_batches(kind, B, 1, 0, False)[0].items()}
gamma = b["gamma"] if spec["use_random_gamma"] else None
for _ in range(3):
    step.step(b, b, gamma)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step.step(b, b, gamma)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
