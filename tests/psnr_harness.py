"""Shared machinery of tests/test_psnr_parity.py and tools/psnr_at_scale.py: train the fused step and the oracle
(reference semantics, torch fp32 on the GPU, TF32 off) for the same number of iterations on the same injected index /
gamma streams at a BASELINE config shape, and score both with the reference's metrics.

TEST INFRASTRUCTURE: imports the oracle; nothing under neucam_b200/ imports this."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from neucam_b200 import harness  # noqa: E402
from neucam_b200.layered import LayeredTrainStep  # noqa: E402
from neucam_b200.step import FusedTrainStep, StepOptions  # noqa: E402
from oracle import neucam_oracle as orc  # noqa: E402

DEV = "cuda"


def hdr_psnr_reference_style(pred, gt):
    """evaluate_hdr (utils.py:733-736): pred /= max; pred = clip(mean(gt)/mean(pred) * pred, 0, 1); PSNR vs gt."""
    pred = pred / pred.max()
    pred = (gt.mean() / pred.mean() * pred).clamp(0, 1)
    return (10 * torch.log10(1 / torch.mean((pred - gt) ** 2))).item()


def index_stream(n_total, B, iters, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randint(0, n_total, (B,), generator=g) for _ in range(iters)]


def gammas(iters, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(1, generator=g) * 100 + 1.0 for _ in range(iters)]


def make_case(kind):
    if kind == "mf_static":
        cfg = harness.CONFIGS["mf_static"]
        shape, focal = cfg["shape"], cfg["focal_list"]
        scene = harness.SyntheticScene(shape, focal, exposures=[-0.5, 0.5], seed=3)
        B = harness.pixels_per_step(cfg)
        build = lambda: harness.build_static_models(shape, focal, hidden=512, seed=1)
        ocfg = orc.StepConfig(shape, use_random_gamma=False, with_grad_loss=False)
        opts = dict(use_random_gamma=False)
        return dict(kind=kind, shape=shape, scene=scene, B=B, build=build, ocfg=ocfg, opts=opts, layered=False)
    spec = harness.LAYERED_CONFIGS[kind]
    shape = spec["shape"]
    scene = harness.SyntheticScene(shape, [0.0] * shape[0], exposures=[-1.5, 0.0, 1.5][:shape[0]], seed=3)
    B = int(spec["sample_frac"] * shape[0] * shape[1] * shape[2])
    build = lambda: harness.build_layered_models(kind, seed=1)
    o = spec["options"]
    nets = {s: dict(last=spec[s][3], skip=(4,), num_frequencies=spec[s][2] if spec[s][1] else None)
            for s in ("uv_b", "uv_f", "alpha") if s in spec}
    ocfg = orc.StepConfig(shape, use_random_gamma=spec["use_random_gamma"], with_grad_loss=False,
                          uv_mapping_scale=o["uv_mapping_scale"], rigidity_loss_w=o.get("rigidity_loss_w"),
                          sparsity_loss_w=o.get("sparsity_loss_w"), alpha_loss_w=o.get("alpha_loss_w"), nets=nets)
    opts = dict(use_random_gamma=spec["use_random_gamma"], **o)
    return dict(kind=kind, shape=shape, scene=scene, B=B, build=build, ocfg=ocfg, opts=opts, layered=True)


@torch.no_grad()
def metrics(case, state):
    """All-in-focus LDR PSNR over every frame -- the summary renderer's metric (utils.py:95-198: blur disabled, uv warp,
    scene network, exposure, CRF, compared with the captured frames) -- and the all-in-focus HDR PSNR of frame 0's pixel
    grid with evaluate_hdr's normalisation (utils.py:733-736).  Evaluated with the ORACLE's fp32 networks for every
    contestant, so the metric itself is not part of the comparison."""
    scene, shape = case["scene"], case["shape"]
    n, h, w = shape
    st = {s: {k: v.to(DEV).float() for k, v in d.items()} for s, d in state.items()}
    grid = scene.mgrid.to(DEV).view(n, h * w, 3)
    frames = scene.frames.to(DEV).view(n, h * w, 3)
    nets = case["ocfg"].nets

    def radiance(pts):
        if case["layered"]:
            uv = (orc.simple_mlp(st["uv_b"], pts, nets["uv_b"]["last"], 3,
                                 num_frequencies=nets["uv_b"]["num_frequencies"]) + pts[:, 1:]) * 0.5
        else:
            uv = pts[:, 1:]
        return orc.sine_mlp(st["atlas_b"], uv)

    se, cnt, hdr0 = 0.0, 0, None
    exps = 3.0 * torch.tanh(st["exp"]["exps"])
    chunk = 1 << 17
    for f in range(n):
        rads = []
        for lo in range(0, h * w, chunk):
            hi = min(lo + chunk, h * w)
            rad = radiance(grid[f, lo:hi])
            rads.append(rad)
            ldr, _ = orc.tonemap(st["tm"], rad, exps[f].expand(hi - lo, 1))
            se += ((ldr / 2 + 0.5 - frames[f, lo:hi]) ** 2).sum().item()
            cnt += (hi - lo) * 3
        if f == 0:
            hdr0 = torch.exp(torch.cat(rads)).t().reshape(3, h, w)
    rad_gt = scene.radiance.to(DEV)
    return {"ldr_psnr": 10 * torch.log10(torch.tensor(cnt / se)).item(),
            "hdr_psnr": hdr_psnr_reference_style(hdr0, rad_gt / rad_gt.max())}


def run_fused(case, idxs, gams):
    models = case["build"]()
    scene, shape, B = case["scene"], case["shape"], case["B"]
    opts = StepOptions(lr=1e-4, **case["opts"])
    if case["layered"]:
        step = LayeredTrainStep(models, shape, opts, device=DEV)
        step.attach_dataset(scene.data, B, gamma_weights=scene.gamma_weights)
        for i, idx in enumerate(idxs):
            mi, gt = step.sample_resident(indices=idx.to(DEV))
            step.graph_step(mi, gt, gams[i] if opts.use_random_gamma else None)
    else:
        step = FusedTrainStep(models, shape, opts, B, device=DEV)
        step.attach_dataset(scene.data, gamma_weights=scene.gamma_weights)
        step.sample_resident(indices=idxs[0].to(DEV), gamma=gams[0])
        step.capture()
        for i, idx in enumerate(idxs):
            step.sample_resident(indices=idx.to(DEV), gamma=gams[i])
            step.run_resident()
    torch.cuda.synchronize()
    return {s: {k: v.detach().float().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}


def run_oracle(case, idxs, gams, permute_seed=None):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    models = case["build"]()
    scene, cfg = case["scene"], case["ocfg"]
    state = {s: {k: v.detach().clone().to(DEV) for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    mom = {s: {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in d.items()} for s, d in state.items()}
    grid, data, gw = scene.mgrid.to(DEV), scene.data.to(DEV), scene.gamma_weights.to(DEV)
    pg = torch.Generator().manual_seed(permute_seed) if permute_seed is not None else None
    for i, idx in enumerate(idxs):
        if pg is not None:
            idx = idx[torch.randperm(idx.numel(), generator=pg)]      # same pixels, another summation order
        idx = idx.to(DEV)
        batch = {"coords": grid[idx][None], "coord_idx": idx[None], "img": data[idx][None], "gamma_weights": gw[idx][None]}
        _, grads = orc.step_grads(state, batch, cfg, gams[i].to(DEV) if cfg.use_random_gamma else None)
        for s, d in state.items():
            for k in d:
                d[k], m, v = orc.adam_update(d[k], grads[s][k], mom[s][k][0], mom[s][k][1], i + 1, cfg.lr)
                mom[s][k] = (m, v)
    torch.cuda.synchronize()
    return state


