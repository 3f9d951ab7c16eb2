"""Checkpoint and evaluation helpers kept compatible with the reference's:

  save_model / load_model : the checkpoint layout of utils.save_model (utils.py:645-673): one torch.save dict
                            of state_dicts under fixed keys, so the reference's render_video.py can load what
                            this package trains and vice versa.
  render_frames           : chunked full-frame evaluation like utils.write_video_summary (utils.py:74-228) /
                            render_video.py:172-261 with the blur model disabled: LDR = tm(atlas_b(y,x), exp[f]),
                            all-in-focus HDR = exp(atlas_b(y,x)) at the un-shifted centre sample (utils.py:164).
  psnr                    : 10 log10(1 / mse) on [0,1] images (utils.py:198).
"""
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
    """Loads a checkpoint written by save_model (or by the reference's utils.save_model) into `models`."""
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
