"""Checkpoint and evaluation helpers kept compatible with the reference's:

  save_model / load_model : the checkpoint layout of utils.save_model (utils.py:645-673): one torch.save dict
                            of state_dicts under fixed keys, so the reference's render_video.py can load what
                            this package trains and vice versa.
  render_frames           : chunked full-frame evaluation like utils.write_video_summary (utils.py:74-228) /
                            render_video.py:172-261 with the blur model disabled: LDR = tm(atlas_b(y,x), exp[f]),
                            all-in-focus HDR = exp(atlas_b(y,x)) at the un-shifted centre sample (utils.py:164).
  psnr                    : 10 log10(1 / mse) on [0,1] images (utils.py:198).
  warm_tm / warm_mapping  : the pre-training loops train_video.py:126-142 runs before the main loop (utils.py:265-334):
  / warm_alpha              CRF towards a fixed gamma curve, uv-mapping networks towards the (scaled) identity, alpha
                            network towards a given mask.  Same objectives, sample counts, learning rates and CPU RNG
                            consumption as the reference, so a seeded run reproduces the reference's warm start; the
                            network forward / backward runs on this package's kernels (tonemap.cu, relu_chain.cu) and
                            the optimizer is the fused flat Adam (adam.cu).
"""
import numpy as np
import torch

from . import harness

# slot name -> key in the checkpoint dict (utils.py:647-670)
CHECKPOINT_KEYS = {
    "atlas_b": "model_state_dict",
    "tm": "model_tm_state_dict",
    "uv_b": "model_mapping_s_state_dict",
    "uv_f": "model_mapping_d_state_dict",
    "blur": "model_blur_state_dict",
    "atlas_f": "model_f_state_dict",
    "alpha": "model_alpha_state_dict",
    "exp": "model_exp_state_dict",
    "focal": "model_focal_state_dict",
    "noise": "model_noise_state_dict",
    "depth": "model_depth_state_dict",
}


def save_model(models, path):
    """Writes {checkpoint key: state_dict} for every non-None model slot."""
    out = {}
    for slot, key in CHECKPOINT_KEYS.items():
        m = models.get(slot)
        if m is not None:
            out[key] = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    torch.save(out, path)


def load_model(models, path, map_location="cpu"):
    """Loads a checkpoint written by save_model (or by the reference's utils.save_model) into `models`.

    Parameters are written IN PLACE (they may be views of a training step's flat buffer).  A step that already exists
    keeps fp16 tensor-core operand copies of the atlas weights; FusedTrainStep notices the write through the flat
    buffer's version counter and re-packs before its next launch (step.repack() does it explicitly)."""
    ckpt = torch.load(path, map_location=map_location)
    for slot, key in CHECKPOINT_KEYS.items():
        m = models.get(slot)
        if m is not None and key in ckpt:
            with torch.no_grad():
                own = m.state_dict()
                for k, v in ckpt[key].items():
                    own[k].copy_(v.to(own[k].device, own[k].dtype))   # in place: parameters may alias a flat buffer
    return ckpt


psnr = harness.psnr


# ------------------------------------------------------------------------------------------------------------------
# warm-ups (reference utils.py:262-334, called from experiment_scripts/train_video.py:126-142)
# ------------------------------------------------------------------------------------------------------------------
def tonemapSimple(x):
    """The fixed target curve of warm_tm (utils.py:262-263)."""
    return (torch.exp(x) / (torch.exp(x) + 1)) ** (1 / 2.2)


class _WarmupAdam:
    """torch.optim.Adam(model.parameters(), lr) as the reference constructs it for the warm-ups, executed by the fused
    flat Adam kernel: the model's parameters are re-homed into one flat buffer (the training step re-homes them again
    later, copying the warmed values)."""

    def __init__(self, model, lr, device):
        from .step import FlatParameters
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("neucam_b200 warm-ups run on the CUDA kernels only (there is no CPU path)")
        model.to(dev)
        self.P = FlatParameters({"atlas_b": model}, dev)    # slot name is irrelevant here: one model, one flat buffer
        self.lr = lr

    def zero_grad(self):
        self.P.zero_grad()

    def step(self):
        from . import ops
        ops.adam_tick(self.P.adam_state)
        ops.adam_step(self.P.flat, self.P.grad, self.P.exp_avg, self.P.exp_avg_sq, self.P.adam_state, self.lr)


def warm_tm(model, pretrain_iters=1000, device="cuda", verbose=False):
    """utils.warm_tm (utils.py:265-282): fit the CRF to tonemapSimple on random radiance in [0,3] and exposure in
    [-3,3] stops, 1024 samples per iteration, Adam lr 5e-4.  Random draws come from the CPU generator in the reference's
    order (radiance, then exposure).  Returns the model; `history` of losses is attached as model.warm_history."""
    opt = _WarmupAdam(model, 5e-4, device)
    hist = []
    for i in range(pretrain_iters):
        opt.zero_grad()
        rand_radiance = torch.rand(1024, 3).to(device) * 3
        rand_exposure = torch.rand(1024, 1).to(device) * 6 - 3.0
        gt_rgb = tonemapSimple(rand_radiance + rand_exposure)
        rgb_l = model(rand_radiance, rand_exposure) / 2 + 0.5
        loss = (rgb_l - gt_rgb).norm(dim=1).mean()
        loss.backward()
        opt.step()
        hist.append(loss.detach())
        if verbose:
            print(f"step {i} pre-train loss: {loss.item()}")
    model.warm_history = torch.stack(hist) if hist else torch.zeros(0)
    return model


def warm_mapping(model_mapping, shape, uv_mapping_scale=0.8, pretrain_iters=100, device="cuda", verbose=False):
    """utils.warm_mapping (utils.py:284-305): pull the uv-mapping network towards uv = scale * (y, x) on 10000 random
    pixels of every frame per iteration, Adam lr 1e-4 (one optimizer step per frame)."""
    frames_num, img_h, img_w = shape
    opt = _WarmupAdam(model_mapping, 1e-4, device)
    hist = []
    for i in range(pretrain_iters):
        for f in range(frames_num):
            i_s_int = torch.randint(img_h, (np.int64(10000), 1))
            j_s_int = torch.randint(img_w, (np.int64(10000), 1))
            i_s = i_s_int / (img_h - 1) * 2 - 1
            j_s = j_s_int / (img_w - 1) * 2 - 1
            t_s = (f / (frames_num - 1 + 1e-16) * 2 - 1) * torch.ones_like(i_s)
            coords = torch.cat((t_s, i_s, j_s), dim=1).to(device)
            uv_temp = model_mapping(coords)
            opt.zero_grad()
            loss = (coords[:, 1:] * uv_mapping_scale - uv_temp).norm(dim=1).mean()
            loss.backward()
            opt.step()
            hist.append(loss.detach())
            if verbose:
                print(f"pre-train loss: {loss.item()}")
    model_mapping.warm_history = torch.stack(hist) if hist else torch.zeros(0)
    return model_mapping


def warm_alpha(model_alpha, dataset, pretrain_iters=1000, device="cuda", verbose=False):
    """utils.warm_alpha (utils.py:307-334): pull the alpha network towards the dataset's mask channel
    (dataset.vid[..., 3:4], [N,H,W,1]) on 10000 random pixels of every frame per iteration, Adam lr 1e-4."""
    frames_num, img_h, img_w = dataset.shape
    opt = _WarmupAdam(model_alpha, 1e-4, device)
    vid = dataset.vid
    mask_gt = (torch.from_numpy(vid[..., 3:4]) if isinstance(vid, np.ndarray) else vid[..., 3:4]).to(device)
    y = torch.linspace(-1, 1, img_h)
    x = torch.linspace(-1, 1, img_w)
    y, x = torch.meshgrid(y, x, indexing="ij")
    xy_grid = torch.stack([y, x], -1).reshape(-1, 2).to(device)
    hist = []
    for i in range(pretrain_iters):
        for f in range(frames_num):
            idx = torch.randint(xy_grid.shape[0], (10000,))
            i_s = xy_grid[idx, 0:1]
            j_s = xy_grid[idx, 1:2]
            t_s = (f / (frames_num - 1) * 2 - 1) * torch.ones_like(i_s)
            coords = torch.cat((t_s, i_s, j_s), dim=1).to(device)
            alpha_temp = model_alpha(coords)
            opt.zero_grad()
            loss = (mask_gt[f].reshape(-1, 1)[idx] - alpha_temp).norm(dim=1).mean()
            loss.backward()
            opt.step()
            hist.append(loss.detach())
            if verbose:
                print(f"pre-train loss: {loss.item()}, max alpha: {torch.max(alpha_temp).item()}")
    model_alpha.warm_history = torch.stack(hist) if hist else torch.zeros(0)
    return model_alpha


@torch.no_grad()
def render_frames(models, shape, chunk=65536, device="cuda"):
    """Returns (ldr [N,H,W,3] in [0,1], hdr [H,W,3] all-in-focus linear radiance of frame 0's pixel grid)."""
    n, h, w = shape
    grid = harness.mgrid(shape).to(device)                         # [N*H*W, 3]
    atlas, tm, exp = models["atlas_b"], models["tm"], models["exp"]
    exps = exp() if exp is not None else None
    ldr = torch.empty(n * h * w, 3, device=device)
    log_rad0 = torch.empty(h * w, 3, device=device)
    for lo in range(0, n * h * w, chunk):
        hi = min(lo + chunk, n * h * w)
        rgb = atlas({"coords": grid[None, lo:hi, 1:]})["model_out"].view(-1, 3)      # log radiance
        if lo < h * w:
            m = min(hi, h * w)
            log_rad0[lo:m] = rgb[: m - lo]
        frame = torch.div(torch.arange(lo, hi, device=device), h * w, rounding_mode="floor")
        e = exps[frame].view(-1, 1) if exps is not None else torch.zeros(hi - lo, 1, device=device)
        ldr[lo:hi] = tm(rgb, e) if tm is not None else rgb
    return (ldr / 2 + 0.5).view(n, h, w, 3), torch.exp(log_rad0).view(h, w, 3)


@torch.no_grad()
def render_layers(models, shape, chunk=65536, device="cuda", frames=None):
    """Full-frame evaluation of any model set at the unshifted pixel positions (the centre tap: all-in-focus), as
    utils.write_video_summary / render_video.py do it (utils.py:95-170): uv warps, both scene layers, alpha blend,
    exposure + CRF.  Returns a dict of device tensors: 'ldr' [F,H,W,3] in [0,1], 'hdr' [F,H,W,3] linear radiance,
    and, when the model set has them, 'alpha' [F,H,W,1], 'uv_b' / 'uv_f' [F,H,W,2].  Runs under no_grad: the fused
    networks store nothing."""
    n, h, w = shape
    frames = list(range(n)) if frames is None else list(frames)
    grid = harness.mgrid(shape).to(device).view(n, h * w, 3)
    exps = models["exp"]() if models.get("exp") is not None else None
    F_ = len(frames)
    out = {"ldr": torch.empty(F_, h * w, 3, device=device), "hdr": torch.empty(F_, h * w, 3, device=device)}
    if models.get("uv_b") is not None:
        out["uv_b"] = torch.empty(F_, h * w, 2, device=device)
    if models.get("uv_f") is not None:
        out["uv_f"] = torch.empty(F_, h * w, 2, device=device)
        out["alpha"] = torch.empty(F_, h * w, 1, device=device)
    for fi, f in enumerate(frames):
        for lo in range(0, h * w, chunk):
            hi = min(lo + chunk, h * w)
            pts = grid[f, lo:hi][None]                                              # [1,M,3]
            pos = pts[..., 1:]
            uv_b = (models["uv_b"](pts).view(1, -1, 2) + pos) * 0.5 if models.get("uv_b") is not None else pos
            rad = models["atlas_b"]({"coords": uv_b})["model_out"]
            if models.get("uv_f") is not None:
                uv_f = (models["uv_f"](pts).view(1, -1, 2) + pos) * 0.5
                front = models["atlas_f"]({"coords": uv_f})["model_out"]
                alpha = (models["alpha"](pts).view(1, -1, 1) if models.get("alpha") is not None
                         else torch.sigmoid(front[..., -1:]))
                rad = (1 - alpha) * rad + alpha * front[..., 0:3]
                out["uv_f"][fi, lo:hi], out["alpha"][fi, lo:hi] = uv_f[0], alpha[0]
            if "uv_b" in out:
                out["uv_b"][fi, lo:hi] = uv_b[0]
            out["hdr"][fi, lo:hi] = torch.exp(rad[0])
            e = exps[f].expand(hi - lo).view(-1, 1) if exps is not None else torch.zeros(hi - lo, 1, device=device)
            ldr = models["tm"](rad.view(-1, 3), e) if models.get("tm") is not None else rad.view(-1, 3)
            out["ldr"][fi, lo:hi] = ldr / 2 + 0.5
    return {k: v.view(F_, h, w, v.shape[-1]) for k, v in out.items()}
