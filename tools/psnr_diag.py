"""Diagnostic: how far do training trajectories drift apart under small gradient perturbations?
Runs, on identical batches: the reference-semantics oracle (fp32, on the GPU for speed, TF32 off), the
oracle with relative Gaussian gradient noise, and the fused CUDA step; prints LDR / all-in-focus HDR PSNR
at several iteration counts."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import neucam_oracle as orc
from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
CHECK = [int(x) for x in os.environ.get("CHECK", "150,500,1500").split(",")]
shape, focal, B = (2, 20, 30), [-1.0, 1.0], 400
scene = harness.SyntheticScene(shape, focal, exposures=[-0.5, 0.5], seed=3, max_radius=2)
gen = torch.Generator().manual_seed(4)
batches = [scene.sample(B, gen) for _ in range(max(CHECK))]
dev = "cuda"

def metrics(state):
    n, h, w = shape
    st = {s: {k: v.to(dev) for k, v in d.items()} for s, d in state.items()}
    grid = scene.mgrid.to(dev); idx = torch.arange(grid.shape[0], device=dev)
    with torch.no_grad():
        cfg = orc.StepConfig(shape, with_grad_loss=False)
        batch = {"coords": grid[None], "coord_idx": idx[None], "img": scene.data.to(dev)[None]}
        _, inter = orc.step_losses(st, batch, cfg, return_intermediates=True)
        ldr = inter["ldr"] / 2 + 0.5; gt = scene.data.to(dev) / 2 + 0.5
        p1 = harness.psnr(ldr, gt).item()
        hdr = torch.exp(orc.sine_mlp(st["atlas_b"], grid[: h * w, 1:])).t().reshape(3, h, w)
        rad = scene.radiance.to(dev)
        p2 = harness.psnr(hdr / hdr.max(), rad / rad.max()).item()
    return p1, p2

def run_oracle(noise, seed=0):
    models = harness.build_static_models(shape, focal, hidden=512, seed=1)
    state = {s: {k: v.detach().clone().to(dev) for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    cfg = orc.StepConfig(shape, with_grad_loss=False)
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    g = torch.Generator(device=dev).manual_seed(seed)
    out = {}
    for i, (inp, gt) in enumerate(batches):
        b = {k: v.to(dev) for k, v in inp.items()}; b["img"] = gt["img"].to(dev)
        _, grads = orc.step_grads(state, b, cfg)
        for s, d in state.items():
            for k in d:
                gr = grads[s][k]
                if noise > 0:
                    gr = gr + noise * gr.norm() / max(gr.numel(), 1) ** 0.5 * torch.randn(gr.shape, generator=g, device=dev)
                p, m, v = orc.adam_update(d[k], gr, mom[s][k][0], mom[s][k][1], i + 1, cfg.lr)
                d[k] = p; mom[s][k] = (m, v)
        if (i + 1) in CHECK:
            out[i + 1] = metrics(state)
    return out

def run_ours():
    models = harness.build_static_models(shape, focal, hidden=512, seed=1)
    step = FusedTrainStep(models, shape, StepOptions(lr=1e-4), B, device=dev)
    out = {}
    for i, (inp, gt) in enumerate(batches):
        step.step(inp, gt)
        if (i + 1) in CHECK:
            torch.cuda.synchronize()
            out[i + 1] = metrics({s: {k: v.detach().float() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None})
    return out

res = {"oracle": run_oracle(0.0), "oracle+1e-3": run_oracle(1e-3), "oracle+5e-3": run_oracle(5e-3, 1),
       "oracle+5e-3 b": run_oracle(5e-3, 2), "ours": run_ours()}
for name, r in res.items():
    print(f"{name:16s} " + "  ".join(f"it{k}: LDR {v[0]:.3f} HDR {v[1]:.3f}" for k, v in r.items()))
