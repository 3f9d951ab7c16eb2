"""A few eager steps of a static workload (default: the headline mfme_blender, Q = 382 725) for ncu:
    ncu --metrics gpu__time_duration.sum ... python tools/profile_static.py [config] [steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions

name = sys.argv[1] if len(sys.argv) > 1 else "mfme_blender"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = harness.CONFIGS[name]
B = harness.pixels_per_step(cfg)
models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
step = FusedTrainStep(models, cfg["shape"], StepOptions(use_random_gamma=cfg["use_random_gamma"]), B, device="cuda")
scene = harness.SyntheticScene(cfg["shape"], cfg["focal_list"], seed=0)
gen = torch.Generator().manual_seed(1)
for i in range(steps):
    inp, gt = scene.sample(B, gen)
    losses = step.step(inp, gt, gamma=torch.tensor([30.0]))
torch.cuda.synchronize()
print(name, "loss", losses["total"].item())
