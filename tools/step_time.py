"""Graph-replayed step time of a static workload (CUDA events, bursts of 20 replays): python tools/step_time.py [config]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions

name = sys.argv[1] if len(sys.argv) > 1 else "mfme_blender"
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
best = []
for rep in range(3):
    time.sleep(0.5)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step.run_resident()
    e1.record(); torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 20)
print(name, "ms/step", " ".join(f"{b:.4f}" for b in best), "loss", step.losses()["total"].item())
