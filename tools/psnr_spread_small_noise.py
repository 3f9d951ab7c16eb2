"""Basin statistics of the fp32 oracle under tiny gradient noise (companion of tools/psnr_spread.py)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_psnr_parity as T
from neucam_b200 import harness
torch.backends.cuda.matmul.allow_tf32 = False
shape, focal, B = (2, 20, 30), [-1.0, 1.0], 400
scene = harness.SyntheticScene(shape, focal, exposures=[-0.5, 0.5], seed=3, max_radius=2)
gen = torch.Generator().manual_seed(4)
batches = [scene.sample(B, gen) for _ in range(T.ITERS)]
for noise in (1e-6, 1e-4):
    vals = []
    for seed in range(1, 9):
        m = T.run_oracle(scene, shape, focal, batches, noise, seed)
        vals.append(round(m[2], 2))
    print("oracle + %.0e relative gradient noise, seeds 1..8: exposure-fitted HDR PSNR %s" % (noise, vals))
