"""TEST INFRASTRUCTURE ONLY -- CPU/PyTorch-fp32 restatement of NeuCam's training hot path.

This file is the oracle the CUDA path is checked against.  It is a from-scratch functional
restatement (explicit weight lists, no nn.Module) of what the reference computes, each function
citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import it; the product (neucam_b200/) never does.

Parity pinning: the reference ships NO tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself run in the build container
(oracle/make_golden.py -> tests/golden/*.pt, checked by tests/test_oracle_vs_golden.py, and live, function by
function, by tests/test_oracle_vs_reference.py whenever /root/reference is present).

State layout: `state` maps the reference's model slot names to the reference's own state_dict keys
(utils.py:645-673), e.g. state['atlas_b']['net.net.0.0.weight'], state['blur']['layers.0.weight'],
state['tm']['layers_r.0.weight'], state['exp']['exps'], state['focal']['focus'].
"""
import math
from collections import OrderedDict

import torch

LN2 = math.log(2.0)


# ------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------
def sine_mlp_layers(sd):
    """[(W, b), ...] from a SingleBVPNet state_dict (modules.py:67-83: net.net.{i}.0.{weight,bias})."""
    out = []
    i = 0
    while f"net.net.{i}.0.weight" in sd:
        out.append((sd[f"net.net.{i}.0.weight"], sd[f"net.net.{i}.0.bias"]))
        i += 1
    return out


def sine_mlp(sd, x, return_hidden=False, teacher=None):
    """SingleBVPNet/FCBlock forward, mode='mlp', type='sine', outermost_linear=True.

    h_{l+1} = sin(30 (h_l W_l^T + b_l)) for every layer but the last, which is linear
    (modules.py:17-26 BatchLinear, 33-35 Sine, 67-83 FCBlock layer list, 474 SingleBVPNet).
    `teacher`: optional list of per-layer INPUTS to feed instead of the running activation
    (teacher-forced per-layer parity, SURVEY.md section 7 hard part 1).
    """
    layers = sine_mlp_layers(sd)
    hidden = []
    h = x
    for li, (W, b) in enumerate(layers):
        if teacher is not None:
            h = teacher[li]
        z = h @ W.t() + b
        h = torch.sin(30.0 * z) if li < len(layers) - 1 else z
        hidden.append(h)
    return (h, hidden) if return_hidden else h


def nerf_posenc(x, num_frequencies):
    """PosEncodingNeRF.forward (modules.py:567-580): [c, sin(2^0 pi c_0), cos(2^0 pi c_0), sin(2^0 pi c_1), ...]."""
    feats = [x]
    for i in range(num_frequencies):
        for j in range(x.shape[-1]):
            c = x[..., j:j + 1]
            feats.append(torch.sin((2.0 ** i) * math.pi * c))
            feats.append(torch.cos((2.0 ** i) * math.pi * c))
    return torch.cat(feats, -1)


def simple_mlp(sd, x, last, in_dim, skip_layer=(4,), num_frequencies=None):
    """SimpleMLP.forward (modules.py:282-312): ReLU MLP, optional NeRF encoding, DETACHED skip concat,
    final nonlinearity 'tanh' | 'sigmoid' | 'softmax' | 'both' | 'none'."""
    n = 0
    while f"layers.{n}.weight" in sd:
        n += 1
    if num_frequencies is not None:
        x = nerf_posenc(x.reshape(x.shape[0], -1, in_dim), num_frequencies)
        x = x.reshape(-1, x.shape[-1])
    else:
        x = x.reshape(-1, in_dim)
    skip_in = x.detach()
    for i in range(n):
        if i in skip_layer:
            x = torch.cat((x, skip_in), -1)
        x = x @ sd[f"layers.{i}.weight"].t() + sd[f"layers.{i}.bias"]
        if i != n - 1:
            x = torch.relu(x)
    if last == "sigmoid":
        x = torch.sigmoid(x)
    elif last == "softmax":
        x = torch.softmax(x, -1)
    elif last == "tanh":
        x = torch.tanh(x)
    elif last == "both":
        k = x.shape[-1] // 3
        x = torch.cat([torch.softmax(x[..., :k], -1), torch.tanh(x[..., k:])], -1)
    elif last != "none":
        raise ValueError(last)
    return x


def tonemap(sd, x, exps):
    """TonemapNet.forward, noise=None branch (modules.py:219-247): lnx = x + ln2*dt, per-channel
    1->W->1 ReLU MLP, tanh."""
    lnx = x + LN2 * exps
    outs = []
    for ci, ch in enumerate("rgb"):
        w0, b0 = sd[f"layers_{ch}.0.weight"], sd[f"layers_{ch}.0.bias"]
        w1, b1 = sd[f"layers_{ch}.2.weight"], sd[f"layers_{ch}.2.bias"]
        hcol = torch.relu(lnx[:, ci:ci + 1] @ w0.t() + b0)
        outs.append(hcol @ w1.t() + b1)
    return torch.tanh(torch.cat(outs, -1)), lnx


def tonemap_min_slope(sd, x, exps):
    """`output_grads=True` branch (modules.py:249-251): min over points of d mean(ldr) / d lnx.
    The reference calls autograd.grad without create_graph, so this is a constant w.r.t. parameters."""
    with torch.enable_grad():
        lnx = (x + LN2 * exps).detach().requires_grad_(True)
        outs = []
        for ci, ch in enumerate("rgb"):
            hcol = torch.relu(lnx[:, ci:ci + 1] @ sd[f"layers_{ch}.0.weight"].t() + sd[f"layers_{ch}.0.bias"])
            outs.append(hcol @ sd[f"layers_{ch}.2.weight"].t() + sd[f"layers_{ch}.2.bias"])
        ldr = torch.tanh(torch.cat(outs, -1))
        g = torch.autograd.grad(ldr.mean(), lnx)[0]
    return g.min().detach()


def gamma_map(x, gamma):
    """utils.gamma_map (utils.py:615-619)."""
    x = x / 2.0 + 0.5
    x = torch.log(x * gamma + 1.0) / torch.log(gamma + 1.0)
    return x * 2.0 - 1.0


# ------------------------------------------------------------------------------------------------
# step helpers
# ------------------------------------------------------------------------------------------------
def neighbour_patch(coords, H, W, patch_size):
    """utils.get_input_patch (utils.py:587-612): dy outer, dx inner, steps 2/(H-1), 2/(W-1).
    coords [1,B,3] -> [K,B,3]."""
    r = {9: (-1, 0, 1), 25: (-2, -1, 0, 1, 2)}[patch_size]
    sy, sx = 2 / (H - 1), 2 / (W - 1)
    taps = []
    for dy in r:
        for dx in r:
            taps.append(torch.cat((coords[..., 0:1], coords[..., 1:2] + dy * sy, coords[..., 2:3] + dx * sx), -1))
    return torch.cat(taps, 0)


def per_frame_gather(values, coord_idx, H, W):
    """training.py:104-108 / 162-166: `.view(-1,1,1).repeat(1,H,W).view(-1,1)[coord_idx]` is just a
    gather by frame = coord_idx // (H*W)."""
    return values[torch.div(coord_idx, H * W, rounding_mode="floor")].view(-1, 1)


class StepConfig:
    """The subset of the reference's `opt` namespace (configs.py) the hot path reads."""

    def __init__(self, shape, patch_size=9, offset_size=5, use_random_gamma=False, fixed_value=0.0,
                 lr=1e-4, with_grad_loss=True, uv_mapping_scale=1.0, rigidity_loss_w=None, sparsity_loss_w=None,
                 alpha_loss_w=None, nets=None):
        self.shape = tuple(shape)  # (N, H, W)  == video_shape
        self.patch_size = patch_size
        self.offset_size = offset_size
        self.use_random_gamma = use_random_gamma
        self.fixed_value = fixed_value
        self.lr = lr
        self.with_grad_loss = with_grad_loss
        self.uv_mapping_scale = uv_mapping_scale
        self.rigidity_loss_w = rigidity_loss_w     # None = rigidity loss off (opt.rigidity_loss / steps window)
        self.sparsity_loss_w = sparsity_loss_w     # None = sparsity loss off
        self.alpha_loss_w = alpha_loss_w           # None = alpha loss off
        # per-slot SimpleMLP hyper-parameters that are not recoverable from a state_dict:
        # {'uv_b': dict(last='tanh', skip=(4,), num_frequencies=None), ...}
        self.nets = nets or {}


def _mlp(state, cfg, slot, x, in_dim=3):
    h = cfg.nets.get(slot, {})
    return simple_mlp(state[slot], x, h.get("last", "tanh"), in_dim, tuple(h.get("skip", (4,))), h.get("num_frequencies"))


def rigidity_loss(state, cfg, slot, coords, src_uv, mask=1.0):
    """utils.get_rigidity_loss (utils.py:402-454), derivative_amount = 1: finite-difference Jacobian of the
    (t,y,x) -> (u,v) mapping against the pixel one step up / one step left, pushed towards a rotation.
    NB the reference scales BOTH u differences by (H-1)/2 and BOTH v differences by (W-1)/2 (utils.py:425-428)."""
    N, H, W = cfg.shape
    sy, sx = 2 / (H - 1), 2 / (W - 1)
    t, y, x = coords[:, :, 0:1], coords[:, :, 1:2], coords[:, :, 2:3]
    tyx_p = torch.cat((torch.cat((t, t), 1), torch.cat((y - sy, y), 1), torch.cat((x, x - sx), 1)), -1).squeeze(0)
    uv_p = (_mlp(state, cfg, slot, tyx_p) + tyx_p[:, 1:3]) / 2
    u_p, v_p = uv_p[:, 0].view(2, -1), uv_p[:, 1].view(2, -1)
    src_uv = src_uv.reshape(-1, 2)
    du, dv = src_uv[:, 0].unsqueeze(0) - u_p, src_uv[:, 1].unsqueeze(0) - v_p
    du_dx, du_dy = du[1] * (H - 1) / 2, du[0] * (H - 1) / 2
    dv_dy, dv_dx = dv[0] * (W - 1) / 2, dv[1] * (W - 1) / 2
    J = torch.stack((torch.stack((du_dx, du_dy), -1), torch.stack((dv_dx, dv_dy), -1)), 1) / cfg.uv_mapping_scale
    JtJ = J.transpose(1, 2) @ J
    a, b = JtJ[:, 0, 0] + 0.001, JtJ[:, 0, 1]
    c, d = JtJ[:, 1, 0], JtJ[:, 1, 1] + 0.001
    inv = torch.stack((torch.stack((d, -b), -1), torch.stack((-c, a), -1)), 1) / (a * d - b * c).view(-1, 1, 1)
    loss = (JtJ ** 2).sum((1, 2)).sqrt() + (inv ** 2).sum((1, 2)).sqrt()
    return (loss * mask).mean()


def step_losses(state, batch, cfg, gamma=None, grad_loss_points=None, return_intermediates=False,
                img_weight=1.0, ue_weight=1.0):
    """One forward pass of the 'Main Pipeline' + 'Loss Function' blocks of training.train
    (training.py:95-267): optional blur MLP + focus, optional uv-mapping MLPs / foreground atlas / alpha MLP and
    layer blend, atlas_b, PSF sum, learned or given exposure, CRF, image_mse (+ random gamma / mask), rigidity,
    sparsity and alpha losses, ue_loss, value-only grad_loss -- i.e. every branch the five shipped configs take.

    batch: {'coords' [1,B,3], 'coord_idx' [1,B] int64, 'img' [1,B,3], optional 'gamma_weights' [1,B,1]}
    gamma: the per-step random gamma scalar (injected; the reference draws it from the CPU RNG, 181).
    grad_loss_points: optional (rand_radiance [M,3], rand_exposure [M,1]) (injected; training.py:248-249).
    Returns OrderedDict of scalar losses in the reference's insertion order (+ 'total').
    """
    N, H, W = cfg.shape
    K = cfg.patch_size
    coords = batch["coords"]
    idx = batch["coord_idx"][0]
    inter = {}

    has_blur = "blur" in state and state["blur"] is not None
    if has_blur:
        patch = neighbour_patch(coords, H, W, K)                                   # utils.py:587-612
        focus = per_frame_gather(state["focal"]["focus"], idx, H, W).unsqueeze(0)  # training.py:104-108
        kernel_out = simple_mlp(state["blur"], torch.cat((focus, coords[..., 1:3]), -1), "both", 3)  # 113
        y_off = kernel_out[:, K:2 * K].t()
        x_off = kernel_out[:, 2 * K:3 * K].t()
        py = patch[:, :, 1] + y_off * (2 / (H - 1)) * cfg.offset_size               # 118
        px = patch[:, :, 2] + x_off * (2 / (W - 1)) * cfg.offset_size               # 119
        patch = torch.stack((patch[:, :, 0], py, px), -1)
        centre = torch.zeros(K, 1, 1, dtype=torch.bool, device=coords.device)
        centre[K // 2] = True
        patch = torch.where(centre, coords.expand(K, -1, -1), patch)                # 120 (centre tap reset)
        weights = kernel_out[:, 0:K].t().unsqueeze(-1)                              # 121
        inter["kernel_out"] = kernel_out
    else:
        patch = coords
        K = 1
        weights = None
    src_uv = patch[..., 1:]                                                         # 128-129
    has = lambda slot: state.get(slot) is not None
    uv = src_uv
    if has("uv_b"):                                                                 # 131-133
        uv = (_mlp(state, cfg, "uv_b", patch).view(K, -1, 2) + src_uv) / 2
    inter["uv"] = uv
    rgb = sine_mlp(state["atlas_b"], uv)                                            # 141
    inter["atlas_out"] = rgb
    alpha = None
    if has("uv_f"):                                                                 # 135-153
        uv_f = (_mlp(state, cfg, "uv_f", patch).view(K, -1, 2) + src_uv) / 2
        rgb_f = sine_mlp(state["atlas_f"], uv_f)
        if has("alpha"):
            alpha = _mlp(state, cfg, "alpha", patch).view(K, -1, 1)
        else:
            alpha = torch.sigmoid(rgb_f[..., -1:])
        rgb = (1 - alpha) * rgb + alpha * rgb_f[..., 0:3]
        alpha = alpha[K // 2].unsqueeze(0)
        inter["uv_f"], inter["atlas_f_out"] = uv_f, rgb_f
    if has_blur:
        rgb = torch.sum(weights * rgb, 0, keepdim=True)                             # 157
    inter["hdr"] = rgb

    if has("exp"):
        exps = per_frame_gather(3.0 * torch.tanh(state["exp"]["exps"]), idx, H, W)  # 161-166, modules.py:407-408
    else:
        exps = batch["exps"].view(-1, 1)                                            # 168
    ldr, _ = tonemap(state["tm"], rgb.view(-1, 3), exps)                            # 171
    inter["ldr"] = ldr

    fake, real = ldr, batch["img"]
    if cfg.use_random_gamma:                                                        # 180-183
        gw = batch["gamma_weights"].view(-1, 1)
        g = torch.as_tensor(gamma, dtype=torch.float32, device=coords.device).view(1)
        fake = fake + gamma_map(ldr, g) * gw
        real = real + gamma_map(batch["img"], g) * gw
    losses = OrderedDict()
    if "weights" in batch and batch.get("use_mask", False):
        losses["img_loss"] = (batch["weights"] * (fake - real) ** 2).mean()         # loss_functions.py:10 (after freeze)
    else:
        losses["img_loss"] = ((fake - real) ** 2).mean()                            # loss_functions.py:12

    ref = K // 2
    if cfg.rigidity_loss_w is not None:                                             # 203-217
        mask = batch["weights"] if ("weights" in batch and batch.get("use_mask", False)) else 1.0
        rig = rigidity_loss(state, cfg, "uv_b", coords, uv[ref], mask)
        if has("uv_f"):
            rig = rig + rigidity_loss(state, cfg, "uv_f", coords, inter["uv_f"][ref], mask)
        losses["rigidity_loss"] = rig * cfg.rigidity_loss_w
    if has("uv_f"):                                                                 # 220-234
        if cfg.sparsity_loss_w is not None:
            not_dyn = (1 - alpha) * torch.exp(inter["atlas_f_out"])
            losses["sparsity_loss1"] = (torch.norm(not_dyn, dim=-1) ** 2).mean() * cfg.sparsity_loss_w
            # the reference multiplies by the BOOLEAN flag opt.sparsity_loss, i.e. by 1 (training.py:227)
            losses["sparsity_loss2"] = torch.mean(-1 / (torch.log(alpha + 1e-24) + torch.log(1 - alpha + 1e-24)))
        if cfg.alpha_loss_w is not None:
            losses["alpha_loss"] = torch.mean(alpha) * cfg.alpha_loss_w
    z3 = torch.zeros(1, 3, device=coords.device, dtype=coords.dtype)     # dtype follows the inputs: the oracle also runs
    z1 = torch.zeros(1, 1, device=coords.device, dtype=coords.dtype)     # in fp64 (its own arithmetic-noise reference)
    ue, _ = tonemap(state["tm"], z3, z1)
    losses["ue_loss"] = torch.mean(torch.abs(ue - cfg.fixed_value))                 # 244-245
    if cfg.with_grad_loss and grad_loss_points is not None:
        slope = tonemap_min_slope(state["tm"], grad_loss_points[0], grad_loss_points[1])
        losses["grad_loss"] = torch.relu(-slope) * 1e6                              # 248-252 (value only)
    # img_weight / ue_weight = 1 reproduce the reference; data-parallel ranks use (B_local/B_global, 1/world)
    # so that the per-rank totals SUM to the single-process loss (neucam_b200/dist.py)
    weights = {"img_loss": img_weight, "ue_loss": ue_weight}
    total = 0.0
    for k, v in losses.items():
        total = total + v.mean() * weights.get(k, 1.0)                              # 256-265
    losses["total"] = total
    return (losses, inter) if return_intermediates else losses


def step_grads(state, batch, cfg, gamma=None, grad_loss_points=None, img_weight=1.0, ue_weight=1.0):
    """Losses + gradients of the total loss w.r.t. every tensor in `state` (training.py:269-270)."""
    leaves = {}
    st = {}
    for slot, sd in state.items():
        if sd is None:
            st[slot] = None
            continue
        st[slot] = {}
        for k, v in sd.items():
            t = v.detach().clone().requires_grad_(True)
            st[slot][k] = t
            leaves[(slot, k)] = t
    losses = step_losses(st, batch, cfg, gamma, grad_loss_points, img_weight=img_weight, ue_weight=ue_weight)
    grads = torch.autograd.grad(losses["total"], list(leaves.values()), allow_unused=True)
    out = {}
    for (slot, k), g in zip(leaves.keys(), grads):
        out.setdefault(slot, {})[k] = torch.zeros_like(leaves[(slot, k)]) if g is None else g
    return {k: v.detach() for k, v in losses.items()}, out


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as constructed at training.py:51 (no weight decay, no amsgrad).
    `step` is the 1-based step count AFTER this update.  Returns new (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# ------------------------------------------------------------------------------------------------
# metrics (utils.py:198 LDR PSNR; utils.py:733-736 HDR normalisation)
# ------------------------------------------------------------------------------------------------
def psnr(pred, gt):
    """10 log10(1 / mean((gt - pred)^2)) for images in [0, 1] (utils.py:198)."""
    return 10.0 * torch.log10(1.0 / torch.mean((gt - pred) ** 2))
