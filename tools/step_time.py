"""Graph-replayed step time of a static workload (CUDA events, three bursts of 20 replays after a 0.5 s pause):
    python tools/step_time.py [config] [same|cycle]
same  = one resident batch replayed (the kernels alone); cycle = 8 resident batches with the five small device copies
bench.py makes per step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions

name = sys.argv[1] if len(sys.argv) > 1 else "mfme_blender"
mode = sys.argv[2] if len(sys.argv) > 2 else "same"
cfg = harness.CONFIGS[name]
B = harness.pixels_per_step(cfg)
models = harness.build_static_models(cfg["shape"], cfg["focal_list"], hidden=512, seed=0)
step = FusedTrainStep(models, cfg["shape"], StepOptions(use_random_gamma=cfg["use_random_gamma"]), B, device="cuda")
host = bench.make_host_batches(cfg, B, 8, seed=100, pin=True)
resident = [{k: v.cuda() for k, v in b.items()} for b in host]


def load(i):
    b = resident[i % 8]
    step.coords.copy_(b["coords"].view(-1, 3))
    step.coord_idx.copy_(b["coord_idx"].view(-1))
    step.img.copy_(b["img"].view(-1, 3))
    step.gamma_w.copy_(b["gamma_weights"].view(-1))
    step.gamma.copy_(b["gamma"])


load(0)
step.capture()
for i in range(5):
    load(i)
    step.run_resident()
torch.cuda.synchronize()
out = []
for rep in range(3):
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        if mode == "cycle":
            load(i)
        step.run_resident()
    e1.record()
    torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / 20)
print(name, mode, "ms/step", " ".join(f"{b:.4f}" for b in out), "loss", step.losses()["total"].item())
