"""North-star end criterion: after the SAME number of training iterations on the SAME batches the fused GPU
step reaches the quality of the reference-semantics step (oracle: fp32 autograd + Adam).

Measured fact this test is built around (tools/psnr_diag.py, profiles/r1_psnr_parity.md): training this
model is chaotic at the trajectory level -- the SAME fp32 oracle run on CPU vs GPU (different summation
order only) differs by 0.3 dB after 150 iterations, and 1e-3 relative gradient noise moves the LDR PSNR by
0.1-0.6 dB.  "Within 0.1 dB of the reference" is therefore only resolvable against the reference's own
perturbation band: the fused step must not be worse than the worst member of {oracle, oracle + 2e-3 relative
gradient noise (2 seeds)} by more than 0.1 dB (2e-3 is the measured gradient error of the fp16 tensor-core
path at init).  LDR PSNR per utils.py:198; all-in-focus HDR PSNR = exp(atlas_b(uv)) at the unshifted centre
sample (utils.py:164), max-normalised (utils.py:733-736) -- reported, but gated on its exposure-fitted variant
because the max-normalisation alone moves it by 2 dB between two runs of the same binary."""
import pytest
import torch

from oracle_replay import orc

from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions

pytestmark = pytest.mark.gpu

ITERS = 300
DEV = "cuda"


def render_metrics(state, scene, shape):
    n, h, w = shape
    st = {s: {k: v.to(DEV) for k, v in d.items()} for s, d in state.items()}
    grid = scene.mgrid.to(DEV)
    idx = torch.arange(grid.shape[0], device=DEV)
    with torch.no_grad():
        cfg = orc.StepConfig(shape, with_grad_loss=False)
        batch = {"coords": grid[None], "coord_idx": idx[None], "img": scene.data.to(DEV)[None]}
        _, inter = orc.step_losses(st, batch, cfg, return_intermediates=True)
        psnr_ldr = harness.psnr(inter["ldr"] / 2 + 0.5, scene.data.to(DEV) / 2 + 0.5).item()
        hdr = torch.exp(orc.sine_mlp(st["atlas_b"], grid[: h * w, 1:])).t().reshape(3, h, w)
        rad = scene.radiance.to(DEV)
        psnr_hdr = harness.psnr(hdr / hdr.max(), rad / rad.max()).item()
        # the same comparison with the least-squares exposure scale instead of the max-normalisation: one bright
        # pixel moves hdr.max() and with it the reference's metric by decibels (observed run to run: 10.5 .. 12.7 dB
        # for the SAME code, float atomics order only), which makes the max-normalised number unusable as a gate
        fit = (hdr * rad).sum() / (hdr * hdr).sum()
        psnr_hdr_fit = harness.psnr(fit * hdr / rad.max(), rad / rad.max()).item()
    return psnr_ldr, psnr_hdr, psnr_hdr_fit


def run_oracle(scene, shape, focal, batches, noise, seed):
    models = harness.build_static_models(shape, focal, hidden=512, seed=1)
    state = {s: {k: v.detach().clone().to(DEV) for k, v in m.state_dict().items()}
             for s, m in models.items() if m is not None}
    cfg = orc.StepConfig(shape, with_grad_loss=False)
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    g = torch.Generator(device=DEV).manual_seed(seed)
    for i, (inp, gt) in enumerate(batches):
        b = {k: v.to(DEV) for k, v in inp.items()}
        b["img"] = gt["img"].to(DEV)
        _, grads = orc.step_grads(state, b, cfg)
        for s, d in state.items():
            for k in d:
                gr = grads[s][k]
                if noise > 0:
                    gr = gr + noise * gr.norm() / max(gr.numel(), 1) ** 0.5 * torch.randn(gr.shape, generator=g, device=DEV)
                p, m, v = orc.adam_update(d[k], gr, mom[s][k][0], mom[s][k][1], i + 1, cfg.lr)
                d[k] = p
                mom[s][k] = (m, v)
    return render_metrics(state, scene, shape)


def test_psnr_after_same_iterations_matches_reference_band():
    torch.backends.cuda.matmul.allow_tf32 = False
    shape, focal, B = (2, 20, 30), [-1.0, 1.0], 400
    scene = harness.SyntheticScene(shape, focal, exposures=[-0.5, 0.5], seed=3, max_radius=2)
    gen = torch.Generator().manual_seed(4)
    batches = [scene.sample(B, gen) for _ in range(ITERS)]

    models = harness.build_static_models(shape, focal, hidden=512, seed=1)
    step = FusedTrainStep(models, shape, StepOptions(lr=1e-4), B, device=DEV)
    for inp, gt in batches:
        step.step(inp, gt)
    torch.cuda.synchronize()
    ours = render_metrics({s: {k: v.detach().float() for k, v in m.state_dict().items()}
                           for s, m in models.items() if m is not None}, scene, shape)

    band = [run_oracle(scene, shape, focal, batches, 0.0, 0),
            run_oracle(scene, shape, focal, batches, 2e-3, 1),
            run_oracle(scene, shape, focal, batches, 2e-3, 2)]
    ldr = [b[0] for b in band]
    hdr = [b[1] for b in band]
    fit = [b[2] for b in band]
    print(f"after {ITERS} iterations: LDR PSNR ours {ours[0]:.3f} dB, reference {ldr[0]:.3f} dB, "
          f"reference + 2e-3 grad noise {ldr[1]:.3f} / {ldr[2]:.3f} dB; all-in-focus HDR PSNR ours {ours[1]:.3f} dB, "
          f"reference {hdr[0]:.3f}, noisy {hdr[1]:.3f} / {hdr[2]:.3f} dB; exposure-fitted HDR PSNR ours {ours[2]:.3f} dB, "
          f"reference {fit[0]:.3f}, noisy {fit[1]:.3f} / {fit[2]:.3f} dB")
    assert ours[0] >= min(ldr) - 0.1
    assert ours[2] >= min(fit) - 0.25
    assert ours[1] >= min(hdr) - 3.0     # max-normalised (utils.py:733-736): reported, gated only loosely (see above)
    # and not absurdly different from the band either (same optimisation problem, same batches)
    assert abs(ours[0] - ldr[0]) <= max(1.0, 3 * (max(ldr) - min(ldr)))
