"""Run-to-run spread of the PSNR-after-N-iterations check: OUR step repeated (float-atomics order is the only
difference between runs) vs the oracle under several gradient-noise seeds.  python tools/psnr_spread.py [runs]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_psnr_parity as T
from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.backends.cuda.matmul.allow_tf32 = False
shape, focal, B = (2, 20, 30), [-1.0, 1.0], 400
scene = harness.SyntheticScene(shape, focal, exposures=[-0.5, 0.5], seed=3, max_radius=2)
gen = torch.Generator().manual_seed(4)
batches = [scene.sample(B, gen) for _ in range(T.ITERS)]
for r in range(runs):
    models = harness.build_static_models(shape, focal, hidden=512, seed=1)
    step = FusedTrainStep(models, shape, StepOptions(lr=1e-4), B, device="cuda")
    for inp, gt in batches:
        step.step(inp, gt)
    torch.cuda.synchronize()
    m = T.render_metrics({s: {k: v.detach().float() for k, v in mm.state_dict().items()} for s, mm in models.items() if mm is not None}, scene, shape)
    print("ours   run %d: LDR %.3f  HDR(max-norm) %.3f  HDR(fitted) %.3f  focus %s" % (r, m[0], m[1], m[2], models["focal"].focus.detach().cpu().flatten().tolist()))
for seed in range(runs):
    noise = 0.0 if seed == 0 else 2e-3
    m = T.run_oracle(scene, shape, focal, batches, noise, seed)
    print("oracle noise %.0e seed %d: LDR %.3f  HDR(max-norm) %.3f  HDR(fitted) %.3f" % (noise, seed, m[0], m[1], m[2]))
