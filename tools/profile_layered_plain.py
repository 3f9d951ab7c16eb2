"""A few eager steps of a layered workload for ncu:  python tools/profile_layered_plain.py video_deblur [steps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from neucam_b200 import harness
from neucam_b200.layered import LayeredTrainStep
from neucam_b200.step import StepOptions

kind = sys.argv[1] if len(sys.argv) > 1 else "video_deblur"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
spec = harness.LAYERED_CONFIGS[kind]
N, H, W = spec["shape"]
B = int(spec["sample_frac"] * N * H * W)
models = harness.build_layered_models(kind, seed=0)
step = LayeredTrainStep(models, spec["shape"], StepOptions(use_random_gamma=spec["use_random_gamma"], lr=1e-4, **spec["options"]), device="cuda")
b = {k: v.cuda() for k, v in bench.layered_host_batches(kind, B, 1, 0, False)[0].items()}
for _ in range(steps):
    losses = step.step(b, b, b["gamma"] if spec["use_random_gamma"] else None)
torch.cuda.synchronize()
print(kind, "loss", losses["total"].item())
