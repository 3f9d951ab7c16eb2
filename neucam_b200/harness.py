"""Host-side harness around the fused step: model construction as the reference's training script does
it, a seeded synthetic multi-focus / multi-exposure scene (no dataset files exist offline), and the
pixel sampler that reproduces the reference's input contract.

Reference anchors:
  * model slots and constructor arguments: experiment_scripts/train_video.py:27-169 + configs/*.txt
  * coordinate grid: dataio.get_mgrid (dataio.py:24-45): (t,y,x) = index/(dim-1)*2-1, t uses max(N-1,1)
  * sampling: dataio.Implicit3DWrapper.__getitem__ (dataio.py:356-393): B = int(frac*N*H*W) indices drawn
    WITH replacement, data = img*2-1, DataLoader adds a leading batch dim of 1.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import modules

SLOTS = ("atlas_b", "uv_b", "atlas_f", "uv_f", "alpha", "exp", "focal", "noise", "blur", "tm", "depth")

# (N, H, W), sample_frac, focal_list, use_random_gamma -- configs/*.txt
CONFIGS = {
    # config_mf_static.txt:6-8,16-24
    "mf_static": dict(shape=(2, 300, 450), sample_frac=3e-2, focal_list=[-1.0, 1.0], use_random_gamma=False),
    # config_mfme_blender.txt as shipped (3 frames)
    "mfme_blender_3": dict(shape=(3, 300, 450), sample_frac=3.5e-2, focal_list=[-1.0, 0.0, 1.0],
                           use_random_gamma=True),
    # BASELINE.json's 9-image 3 focus x 3 exposure variant of config_mfme_blender.txt
    "mfme_blender": dict(shape=(9, 300, 450), sample_frac=3.5e-2,
                         focal_list=[-1.0, -1.0, -1.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0], use_random_gamma=True),
}


# the uv-mapping / two-layer configs: configs/config_{me_dynamic,video_deblur,video_hdr}.txt.  Per SimpleMLP:
# (num_layer, use_encoding, num_frequencies, last nonlinearity); hidden 256, in 3, skip [4] throughout.
LAYERED_CONFIGS = {
    "me_dynamic": dict(shape=(3, 448, 680), sample_frac=6e-2, use_random_gamma=True, deblur=False,
                       uv_b=(4, False, None, "tanh"),
                       options=dict(uv_mapping_scale=1.0, rigidity_loss_w=1.0, rigidity_loss_steps=50000)),
    "video_deblur": dict(shape=(80, 360, 640), sample_frac=1e-3, use_random_gamma=False, deblur=True,
                         uv_b=(4, True, 3, "tanh"), uv_f=(4, True, 3, "tanh"), alpha=(4, True, 4, "sigmoid"),
                         options=dict(uv_mapping_scale=1.0, rigidity_loss_w=1e-2, rigidity_loss_steps=20000,
                                      sparsity_loss_w=1e-3, alpha_loss_w=5e-2)),
    "video_hdr": dict(shape=(80, 360, 640), sample_frac=1e-3, use_random_gamma=False, deblur=False,
                      uv_b=(4, False, None, "tanh"), uv_f=(6, True, 3, "tanh"), alpha=(4, True, 3, "sigmoid"),
                      options=dict(uv_mapping_scale=0.8, rigidity_loss_w=1e-2, rigidity_loss_steps=500000,
                                   sparsity_loss_w=5e-2)),
}


def build_layered_models(kind, shape=None, hidden=512, mlp_hidden=256, seed=0):
    """Model set of a uv-mapping config, constructed like train_video.py:27-122 does for it."""
    spec = LAYERED_CONFIGS[kind]
    torch.manual_seed(seed)
    n, h, w = shape if shape is not None else spec["shape"]
    m = OrderedDict((k, None) for k in SLOTS)

    def mlp(hp, out_dim):
        layers, enc, freqs, last = hp
        return modules.SimpleMLP(hidden_dim=mlp_hidden, num_layer=layers, in_dim=3, out_dim=out_dim,
                                 use_encoding=enc, num_frequencies=freqs, skip_layer=[4], nonlinear=last)

    m["atlas_b"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=hidden, num_hidden_layers=3,
                                        in_features=2, out_features=3, sidelength=(h, w), num_frequencies=7)
    m["uv_b"] = mlp(spec["uv_b"], 2)
    if "uv_f" in spec:
        m["atlas_f"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=hidden, in_features=2,
                                            out_features=3 if "alpha" in spec else 4, sidelength=(h, w),
                                            num_frequencies=7)
        m["uv_f"] = mlp(spec["uv_f"], 2)
        if "alpha" in spec:
            m["alpha"] = mlp(spec["alpha"], 1)
    m["exp"] = modules.LearnExps(n)
    if spec["deblur"]:
        m["focal"] = modules.LearnFocal(n)
        m["blur"] = modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                      num_frequencies=7, skip_layer=[4], nonlinear="both")
    m["tm"] = modules.TonemapNet()
    return m


def pixels_per_step(cfg):
    n, h, w = cfg["shape"]
    return int(cfg["sample_frac"] * n * h * w)   # dataio.py:351


def build_static_models(shape, focal_list, hidden=512, seed=0):
    """atlas_b + blur + focal + exp + tm, constructed like train_video.py:27-122 for the configs with
    `deblur`, `pre_hdr`, `learn_exp`, `learn_focal` set (config_mf_static.txt / config_mfme_blender.txt)."""
    torch.manual_seed(seed)
    n, h, w = shape
    m = OrderedDict((k, None) for k in SLOTS)
    m["atlas_b"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=hidden, num_hidden_layers=3,
                                        in_features=2, out_features=3, sidelength=(h, w), num_frequencies=7)
    m["exp"] = modules.LearnExps(n)
    m["focal"] = modules.LearnFocal(list(focal_list), True) if len(focal_list) > 1 else modules.LearnFocal(n)
    m["blur"] = modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                  num_frequencies=7, skip_layer=[4], nonlinear="both")
    m["tm"] = modules.TonemapNet()
    return m


def load_state(models, state):
    """state: {slot: state_dict} with the reference's keys."""
    for slot, sd in state.items():
        if sd is not None and models.get(slot) is not None:
            models[slot].load_state_dict(sd)


def mgrid(shape):
    """dataio.get_mgrid(dim=3) flattened to [N*H*W, 3] fp32."""
    n, h, w = shape
    t = np.arange(n, dtype=np.float64) / max(n - 1, 1)
    y = np.arange(h, dtype=np.float64) / (h - 1)
    x = np.arange(w, dtype=np.float64) / (w - 1)
    g = np.stack(np.meshgrid(t, y, x, indexing="ij"), -1)
    g = (g - 0.5) * 2.0
    return torch.from_numpy(g.reshape(-1, 3)).float()


def _crf(x):
    """A fixed gamma-like camera response on linear exposure, [0, inf) -> [0, 1)."""
    return (x / (x + 0.6)) ** (1 / 2.2)


class SyntheticScene:
    """Seeded HDR scene + its N captured LDR frames (different focus / exposure per frame).

    radiance R(y,x): smooth positive field (random 2-D sinusoids) with a few hard edges, range ~[0.05, 5];
    depth: two layers; frame f = CRF(2^{dt_f} * PSF_{|focus_f - depth|} (*) R), values in [0,1].
    """

    def __init__(self, shape, focal_list, exposures=None, seed=0, max_radius=4):
        n, h, w = shape
        g = torch.Generator().manual_seed(seed)
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
        rad = torch.zeros(3, h, w)
        for _ in range(16):
            fy, fx = (torch.rand(2, generator=g) * 14 + 1).tolist()
            ph = (torch.rand(3, generator=g) * 2 * math.pi)
            amp = torch.rand(3, generator=g) * 0.6
            for c in range(3):
                rad[c] += amp[c] * torch.sin(fy * yy + fx * xx + ph[c])
        rad = torch.exp(rad * 0.9)
        for _ in range(4):   # hard edges: bright / dark rectangles
            y0, x0 = int(torch.randint(0, h - 8, (1,), generator=g)), int(torch.randint(0, w - 8, (1,), generator=g))
            hh, ww = int(torch.randint(6, max(h // 3, 8), (1,), generator=g)), int(torch.randint(6, max(w // 3, 8), (1,), generator=g))
            rad[:, y0:y0 + hh, x0:x0 + ww] *= float(torch.rand(1, generator=g) * 3 + 0.2)
        self.radiance = rad.clamp(0.05, 5.0)                       # [3,H,W] GT all-in-focus HDR
        self.depth = torch.where(xx + 0.3 * torch.sin(3 * yy) > 0.1, torch.tensor(1.0), torch.tensor(-1.0))
        self.shape = shape
        self.focal_list = list(focal_list)
        if exposures is None:
            exposures = [((f % 3) - 1) * 1.5 for f in range(n)] if n >= 3 else [0.0] * n
        self.exposures = list(exposures)
        frames = []
        for f in range(n):
            img = torch.zeros_like(self.radiance)
            for layer in (-1.0, 1.0):
                r = int(round(abs(self.focal_list[f] - layer) / 2.0 * max_radius))
                blurred = self.radiance
                if r > 0:
                    k = 2 * r + 1
                    ker = torch.ones(1, 1, k, k) / (k * k)
                    blurred = F.conv2d(F.pad(self.radiance[:, None], (r, r, r, r), mode="replicate"), ker)[:, 0]
                img = torch.where(self.depth == layer, blurred, img)
            frames.append(_crf(img * (2.0 ** self.exposures[f])))
        self.frames = torch.stack(frames).permute(0, 2, 3, 1).contiguous()      # [N,H,W,3] in [0,1]
        self.data = ((self.frames - 0.5) / 0.5).reshape(-1, 3)                   # dataio.py:349
        self.mgrid = mgrid(shape)
        self.gamma_weights = torch.ones(n * h * w, 1)

    def sample(self, B, generator):
        """One batch in the reference's DataLoader layout."""
        idx = torch.randint(0, self.data.shape[0], (B,), generator=generator)
        model_input = {"coords": self.mgrid[idx][None], "coord_idx": idx[None],
                       "gamma_weights": self.gamma_weights[idx][None]}
        return model_input, {"img": self.data[idx][None]}


def psnr(pred, gt):
    """10 log10(1 / mean((gt - pred)^2)), images in [0,1] (utils.py:198)."""
    return 10.0 * torch.log10(1.0 / torch.mean((gt - pred) ** 2))
