"""The training step of the uv-mapping / two-layer configurations (config_me_dynamic, config_video_deblur,
config_video_hdr): reference training.py:95-271 with `models['uv_b']`, and optionally `models['uv_f']` +
`models['atlas_f']` + `models['alpha']`, in play.  SURVEY.md section 8f rows "next" 1 and 2.

What runs where:
  * both scene networks (atlas_b, atlas_f)     fused tcgen05 chains + split-K wgrad kernels, through
                                               modules.SingleBVPNet -> _SineMLPFunction (forward, dgrad incl.
                                               d/d(u,v), weight gradients)
  * uv-mapping / alpha networks                tcgen05 ReLU chains (relu_mlp.py: encoding, forward, input gradient,
                                               weight gradients)
  * blur taps; PSF sum + exposure + CRF +      the per-pixel kernels of the static step (pixel_ops.py: blur_taps,
    image_mse (+gamma, mask) + ue_loss         psf_crf_loss -- forward and gradients in one pass)
  * Adam over every model                      ONE fused launch over the flat parameter buffer (adam.cu)
  * value-only grad_loss                       neucam_tm_min_slope
  * rigidity loss; layer blend + sparsity /    losses.cu through pixel_ops.py (values and analytic gradients)
    alpha losses
  * warp = (mlp + position) / 2, tap           torch elementwise ops on the device under autograd
    assembly, final sums

Nothing here has a CPU path: the scene networks raise on CPU tensors and the optimizer is the CUDA kernel.

Data parallel: like FusedTrainStep, every rank takes a slice of the step's pixels; per-pixel means are
weighted by B_local / B_global, the pixel-free ue_loss by 1 / world, and the flat gradient buffer is summed
once per step.
"""
from collections import OrderedDict

import torch

from . import _gradmode, _lib, ops, pixel_ops
from .step import FlatParameters, StepOptions, flat_adam  # noqa: F401  (StepOptions re-exported)

_LAYER_SLOTS = ("uv_b", "uv_f", "atlas_f", "alpha")


def needs_layered_step(models):
    """True when the model set has any slot FusedTrainStep does not execute."""
    return any(models.get(s) is not None for s in _LAYER_SLOTS) or models.get("blur") is None \
        or models.get("exp") is None


class LayeredTrainStep:
    def __init__(self, models, video_shape, options, device=None, process_group=None, global_pixels=None,
                 _host_logic_only=False):
        # _host_logic_only: CPU unit tests of the glue above the kernels (tests/test_layered_cpu.py swap the scene
        # networks for a stand-in); such an instance cannot run the optimizer and is not a product path
        if not _host_logic_only:
            _lib.check(_lib.lib().neucam_check_device(), "neucam_check_device (need an sm_100 GPU)")
        for slot in ("noise", "depth"):
            if models.get(slot) is not None:
                raise NotImplementedError(f"model slot '{slot}' is not used by any shipped NeuCam config")
        if models.get("atlas_b") is None or models.get("tm") is None:
            raise NotImplementedError("the step needs at least atlas_b and tm (every shipped config has pre_hdr)")
        if models.get("uv_f") is not None and models.get("atlas_f") is None:
            raise ValueError("uv_f without atlas_f (train_video.py:48-58 builds atlas_f whenever split_layer and "
                             "uv_mapping are both set)")
        if (models.get("blur") is None) != (models.get("focal") is None):
            raise ValueError("blur and focal come together (deblur + learn_focal)")
        self.models = models
        self.shape = tuple(int(s) for s in video_shape)
        self.opt = options
        if _host_logic_only:
            self.device = torch.device("cpu")
        else:
            self.device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        self.global_pixels = global_pixels
        self.params = None if _host_logic_only else FlatParameters(models, self.device)
        if self.params is not None:
            ops.check_tm_shapes(list(models["tm"].parameters()))   # the CRF kernels execute TonemapNet(W=128) only
            if self.world > 1:
                self.params.broadcast_from_rank0(self.pg)          # replicas start bit-identical by construction
        n, h, w = self.shape
        self.K = int(options.patch_size) if models.get("blur") is not None else 1
        if self.K not in (1, 9, 25):
            raise ValueError("patch_size must be 9 or 25 (utils.get_input_patch)")
        if not _host_logic_only and self.K > 1:
            blur = models["blur"]
            if not (self.K == 9 and len(blur.layers) == 4 and not blur.use_encoding and blur.layers[0].in_features == 3
                    and blur.layers[0].out_features == 64 and blur.layers[3].out_features == 27
                    and blur.last_nonlinear == "both"):
                raise NotImplementedError("the blur kernels execute SimpleMLP(64, 4 layers, 3 -> 27, 'both') with "
                                          "patch_size 9 (every shipped config with `deblur`)")
        side = {1: (0,), 9: (-1, 0, 1), 25: (-2, -1, 0, 1, 2)}[self.K]
        # tap k = (row dy, column dx) of the neighbourhood, dy outer (utils.py:587-612)
        self.tap_shift = torch.tensor([[0.0, dy * 2 / (h - 1), dx * 2 / (w - 1)] for dy in side for dx in side],
                                      device=self.device).view(self.K, 1, 3)
        # constants of the pipeline live on the device once (no host->device traffic inside a step / a graph)
        self._one_row_up = torch.tensor([0.0, 2 / (h - 1), 0.0], device=self.device)
        self._one_col_left = torch.tensor([0.0, 0.0, 2 / (w - 1)], device=self.device)
        self._off_centre = torch.ones(self.K, 1, 1, device=self.device)
        self._off_centre[self.K // 2] = 0.0      # the centre tap stays on the pixel (training.py:120)
        self._black = (torch.zeros(1, 3, device=self.device), torch.zeros(1, 1, device=self.device))
        self._one = torch.ones(1, device=self.device)
        self.step_index = 0
        self._losses = None

    # -- pieces of the pipeline ------------------------------------------------------------------------------
    def _frame_value(self, per_frame, coord_idx):
        n, h, w = self.shape
        return per_frame.reshape(-1)[torch.div(coord_idx, h * w, rounding_mode="floor")].view(-1, 1)

    def _taps(self, coords, coord_idx):
        """[K,B,3] sample positions and [K,B,1] PSF weights (training.py:99-126)."""
        if self.K == 1:
            return coords, None
        n, h, w = self.shape
        K = self.K
        focus = self._frame_value(self.models["focal"](), coord_idx)
        kernel = self.models["blur"](torch.cat((focus, coords[0, :, 1:3]), -1))            # [B, 3K]
        shift = torch.stack((torch.zeros_like(kernel[:, :K]),
                             kernel[:, K:2 * K] * (2 / (h - 1) * self.opt.offset_size),
                             kernel[:, 2 * K:] * (2 / (w - 1) * self.opt.offset_size)), -1)  # [B,K,3]
        taps = coords + self.tap_shift + shift.permute(1, 0, 2) * self._off_centre
        return taps, kernel[:, :K].t().unsqueeze(-1)

    def _warp(self, slot, taps):
        """(t,y,x) -> atlas (u,v): half-way between the MLP's output and the pixel position (training.py:131-137)."""
        K = taps.shape[0]
        return (self.models[slot](taps).view(K, -1, 2) + taps[..., 1:]) * 0.5

    def _rigidity(self, slot, coords, uv_ref, mask):
        """utils.get_rigidity_loss (utils.py:402-454) with derivative_amount 1, in closed form for the 2x2
        algebra: J = d(u,v)/d(x,y) by backward differences, loss = |J^T J|_F + |(J^T J + 0.001 I)^-1|_F.
        The reference converts BOTH u differences with (H-1)/2 and BOTH v differences with (W-1)/2 (utils.py:425-428)."""
        n, h, w = self.shape
        pts = coords[0]
        moved = torch.cat((pts - self._one_row_up, pts - self._one_col_left), 0)
        uv_moved = ((self.models[slot](moved) + moved[:, 1:3]) * 0.5).view(2, -1, 2)
        if self.params is not None:
            # product path: value and analytic gradient of the 2x2 algebra in one kernel; the mean over pixels (and
            # the reference's mean(loss) * mean(weights), utils.py:452-453) is folded into the device scalar `coef`
            coef = self._one * (1.0 / pts.shape[0])
            if mask is not None:
                coef = coef * mask.mean()
            return pixel_ops.rigidity_loss(uv_ref, uv_moved, coef, self.shape, self.opt.uv_mapping_scale)
        d_up, d_left = uv_ref.reshape(-1, 2) - uv_moved[0], uv_ref.reshape(-1, 2) - uv_moved[1]
        s = 1.0 / self.opt.uv_mapping_scale
        a, b = d_left[:, 0] * ((h - 1) / 2 * s), d_up[:, 0] * ((h - 1) / 2 * s)      # du/dx, du/dy
        c, d = d_left[:, 1] * ((w - 1) / 2 * s), d_up[:, 1] * ((w - 1) / 2 * s)      # dv/dx, dv/dy
        p, q, r = a * a + c * c, a * b + c * d, b * b + d * d                         # J^T J = [[p,q],[q,r]]
        pe, re = p + 0.001, r + 0.001
        det = pe * re - q * q
        loss = torch.sqrt(p * p + 2 * q * q + r * r) + torch.sqrt(pe * pe + 2 * q * q + re * re) / det
        if mask is not None:
            # the reference multiplies the [B] loss by the [1,B,1] weights, which broadcasts to [1,B,B]; the mean
            # of that outer product is mean(loss) * mean(weights) (utils.py:452-453)
            return loss.mean() * mask.mean()
        return loss.mean()

    # -- forward ----------------------------------------------------------------------------------------------
    def forward(self, model_input, gt, gamma=None, use_mask=False):
        """Losses of one step as an OrderedDict of device scalars (this rank's share), graph attached."""
        m, opt = self.models, self.opt
        dev = self.device
        coords = model_input["coords"].to(dev, torch.float32).view(1, -1, 3)
        coord_idx = model_input["coord_idx"].to(dev).view(-1)
        target = gt["img"].to(dev, torch.float32).view(-1, 3)
        B = coords.shape[1]
        K = self.K
        global_pixels = self.global_pixels if self.global_pixels is not None else B * self.world
        share = B / float(global_pixels)

        kernels = self.params is not None      # product path; False only for the CPU host-logic tests
        if kernels and K > 1:
            uv_taps, psf = pixel_ops.blur_taps(coords[0].contiguous(), coord_idx, m["focal"].focus, m["blur"],
                                               self.shape, K, opt.offset_size)
            taps = torch.cat((coords[..., :1].expand(K, -1, -1), uv_taps), -1)
        else:
            taps, psf = self._taps(coords, coord_idx)
        reg_total = reg_values = None
        uv_b = self._warp("uv_b", taps) if m.get("uv_b") is not None else taps[..., 1:]
        radiance = m["atlas_b"]({"coords": uv_b})["model_out"]                             # [K,B,3] log radiance
        alpha = front = uv_f = None
        if m.get("uv_f") is not None:
            uv_f = self._warp("uv_f", taps)
            front = m["atlas_f"]({"coords": uv_f})["model_out"]
        if m.get("atlas_f") is not None:
            if front is None:
                raise ValueError("atlas_f without uv_f: the reference would fail at training.py:149 as well")
            alpha = m["alpha"](taps).view(K, -1, 1) if m.get("alpha") is not None else torch.sigmoid(front[..., -1:])
            if kernels and m.get("alpha") is not None and front.shape[-1] == 3:
                # layer blend + sparsity / alpha regularisers (values and gradients) in one kernel pair
                n_taps = float(K * B)
                c1 = opt.sparsity_loss_w * share / n_taps if opt.sparsity_loss_w is not None else 0.0
                c2 = share / B if opt.sparsity_loss_w is not None else 0.0
                c3 = opt.alpha_loss_w * share / B if opt.alpha_loss_w is not None else 0.0
                radiance, reg_total, reg_values = pixel_ops.layer_blend(radiance, front, alpha, c1, c2, c3)
            else:
                radiance = torch.lerp(radiance, front[..., 0:3], alpha)                    # training.py:152
            alpha = alpha[K // 2].unsqueeze(0)
        mask = model_input["weights"].to(dev, torch.float32).view(-1, 1) if use_mask else None
        losses = OrderedDict()
        if kernels:
            # PSF sum, exposure, CRF, image_mse (+ gamma, mask), ue_loss and all their gradients: one kernel
            given_exp = None if m.get("exp") is not None else model_input["exps"].to(dev, torch.float32).view(-1).contiguous()
            gw = model_input["gamma_weights"].to(dev, torch.float32).view(-1).contiguous() if opt.use_random_gamma else None
            g_dev = torch.as_tensor(gamma, dtype=torch.float32).to(dev).view(1) if opt.use_random_gamma else None
            pixel_total, img_value, ue_value = pixel_ops.psf_crf_loss(
                radiance.contiguous(), psf, m["exp"].exps if m.get("exp") is not None else None, given_exp, m["tm"],
                self.shape, K, opt.fixed_value, 1.0 / self.world, global_pixels, coord_idx, target.contiguous(), gw,
                mask.view(-1).contiguous() if mask is not None else None, g_dev)
            losses["img_loss"] = img_value
        else:
            if psf is not None:
                radiance = (psf * radiance).sum(0, keepdim=True)                           # training.py:157
            if m.get("exp") is not None:
                exposure = self._frame_value(m["exp"](), coord_idx)
            else:
                exposure = model_input["exps"].to(dev, torch.float32).view(-1, 1)
            ldr = m["tm"](radiance.view(-1, 3), exposure, None)
            fake, real = ldr, target
            if opt.use_random_gamma:                                                       # training.py:180-183
                g = torch.as_tensor(gamma, dtype=torch.float32).to(dev).view(1)
                gw = model_input["gamma_weights"].to(dev, torch.float32).view(-1, 1)
                log1pg = torch.log(g + 1.0)
                curve = lambda x: torch.log((x * 0.5 + 0.5) * g + 1.0) / log1pg * 2.0 - 1.0    # utils.gamma_map
                fake, real = fake + curve(ldr) * gw, real + curve(target) * gw
            sq = (fake - real) ** 2
            losses["img_loss"] = (sq * mask if mask is not None else sq).mean() * share    # loss_functions.py:8-12

        if opt.rigidity_loss_w is not None and self.step_index < opt.rigidity_loss_steps:  # training.py:203-217
            rig = self._rigidity("uv_b", coords, uv_b[K // 2], mask)
            if uv_f is not None:
                rig = rig + self._rigidity("uv_f", coords, uv_f[K // 2], mask)
            losses["rigidity_loss"] = rig * (opt.rigidity_loss_w * share)
        if reg_values is not None:                                                         # training.py:220-238
            if opt.sparsity_loss_w is not None:
                losses["sparsity_loss1"], losses["sparsity_loss2"] = reg_values[0], reg_values[1]
            if opt.alpha_loss_w is not None:
                losses["alpha_loss"] = reg_values[2]
        elif uv_f is not None:
            if opt.sparsity_loss_w is not None:
                hidden_front = (1 - alpha) * torch.exp(front)
                losses["sparsity_loss1"] = hidden_front.square().sum(-1).mean() * (opt.sparsity_loss_w * share)
                entropy = torch.log(alpha + 1e-24) + torch.log(1 - alpha + 1e-24)
                losses["sparsity_loss2"] = (-1 / entropy).mean() * share
            if opt.alpha_loss_w is not None:
                losses["alpha_loss"] = alpha.mean() * (opt.alpha_loss_w * share)
        if kernels:
            losses["ue_loss"] = ue_value
            detached = ("img_loss", "ue_loss") + (("sparsity_loss1", "sparsity_loss2", "alpha_loss") if reg_total is not None else ())
            losses["total"] = pixel_total + sum(v for k, v in losses.items() if k not in detached)
            if reg_total is not None:
                losses["total"] = losses["total"] + reg_total
        else:
            black = m["tm"](*self._black)                                                  # training.py:244-245
            losses["ue_loss"] = (black - opt.fixed_value).abs().mean() / self.world
            losses["total"] = sum(losses.values())
        return losses

    # -- backward / optimizer ---------------------------------------------------------------------------------
    def forward_backward(self, model_input, gt, gamma=None, use_mask=False):
        if self.params is None:
            raise _lib.NeucamNativeError("LayeredTrainStep was built for host-logic tests only; no optimizer state")
        self.params.zero_grad()
        # the flat gradient buffer is preallocated and zeroed: the fused networks' weight-gradient kernels add into it
        with _gradmode.accumulate_into_param_grad(), _gradmode.one_step():
            losses = self.forward(model_input, gt, gamma, use_mask)
            losses["total"].backward()
        self._losses = OrderedDict((k, v.detach()) for k, v in losses.items())
        return self._losses

    def reduce_gradients(self):
        if self.world > 1:
            from .dist import allreduce_flat_gradients
            allreduce_flat_gradients(self.params.grad, self.pg)

    def optimizer_step(self, lr=None, reduce=True, lr_groups=None):
        if reduce:
            self.reduce_gradients()
        flat_adam(self.params, self.opt.lr if lr is None else lr, lr_groups)

    def grad_loss(self, rand_radiance=None, rand_exposure=None):
        """Logged-only relu(-min CRF slope) * 1e6 over 10000 random points (training.py:248-252)."""
        if rand_radiance is None:
            rand_radiance = torch.rand(10000, 3) * 5
            rand_exposure = torch.rand(10000, 1) * 6 - 3.0
        slope = ops.tm_min_slope([p.data for p in self.models["tm"].parameters()],
                                 rand_radiance.to(self.device).contiguous(), rand_exposure.to(self.device).contiguous())
        return torch.relu(-slope) * 1e6

    def losses(self):
        return self._losses

    # -- device-resident data ----------------------------------------------------------
    def attach_dataset(self, data, pixels_per_step, gamma_weights=None, mask_weights=None):
        """Makes the video resident in HBM (data [N*H*W,3] = img*2-1, dataio.py:349) so that batches are formed on the
        device by sample_resident() (neucam_sample_batch) instead of the CPU DataLoader."""
        dev = self.device
        B = int(pixels_per_step)
        self.ds_data = data.reshape(-1, 3).to(dev, torch.float32).contiguous()
        self.ds_gamma = None if gamma_weights is None else gamma_weights.reshape(-1).to(dev, torch.float32).contiguous()
        self.ds_mask = None if mask_weights is None else mask_weights.reshape(-1).to(dev, torch.float32).contiguous()
        self._ds = {"coords": torch.zeros(B, 3, device=dev), "coord_idx": torch.zeros(B, dtype=torch.int64, device=dev),
                    "img": torch.zeros(B, 3, device=dev), "gamma_w": torch.zeros(B, device=dev),
                    "mask_w": torch.zeros(B, device=dev), "idx_in": torch.zeros(B, dtype=torch.int64, device=dev)}

    def sample_resident(self, seed=0, step=0, indices=None):
        """Forms a batch on the device, Philox-sampled from (seed, step) or gathered at injected `indices`, and returns
        it in the reference's DataLoader layout (device tensors): (model_input, gt) for step() / graph_step()."""
        d = self._ds
        idx = None
        if indices is not None:
            d["idx_in"].copy_(indices.reshape(-1), non_blocking=True)
            idx = d["idx_in"]
        ops.sample_batch(self.ds_data, self.ds_gamma, self.ds_mask, idx, self.shape, seed, step, d["coords"],
                         d["coord_idx"], d["img"], d["gamma_w"] if self.ds_gamma is not None else None,
                         d["mask_w"] if self.ds_mask is not None else None)
        model_input = {"coords": d["coords"][None], "coord_idx": d["coord_idx"][None]}
        if self.ds_gamma is not None:
            model_input["gamma_weights"] = d["gamma_w"].view(1, -1, 1)
        if self.ds_mask is not None:
            model_input["weights"] = d["mask_w"].view(1, -1, 1)
        return model_input, {"img": d["img"][None]}

    # -- CUDA graph ---------------------------------------------------------------------------------------------
    def capture(self, model_input, gt, gamma=None, use_mask=False, lr_groups=None, adopt=False):
        """Captures forward + autograd backward + gradient all-reduce (NCCL collectives are captured like any other
        stream work) + Adam over STATIC device copies of one batch into ONE CUDA graph; `run_resident()` replays it
        after `load_resident()` refreshed the static buffers.  The rigidity window is decided at capture time.
        adopt=True: the given DEVICE tensors become the static buffers themselves (no copies) -- pass what
        `sample_resident()` returned and every later `sample_resident()` + `run_resident()` is a step on a fresh batch
        with no host involvement and no staging copy (gamma: a device tensor [1] the caller updates in place)."""
        dev = self.device
        if adopt:
            self._static_in = {k: v for k, v in model_input.items() if torch.is_tensor(v)}
            self._static_gt = {"img": gt["img"]}
            self._static_gamma = gamma
            given = list(self._static_in.values()) + [gt["img"]] + ([gamma] if gamma is not None else [])
            if any(t.device.type != dev.type or (dev.index is not None and t.device.index != dev.index) for t in given):
                raise ValueError("capture(adopt=True) needs tensors that already live on the step's device")
        else:
            self._static_in = {k: v.to(dev).clone() for k, v in model_input.items() if torch.is_tensor(v)}
            self._static_gt = {"img": gt["img"].to(dev).clone()}
            self._static_gamma = None if gamma is None else torch.as_tensor(gamma, dtype=torch.float32).to(dev).view(1).clone()
        P = self.params
        saved = [t.clone() for t in (P.flat, P.exp_avg, P.exp_avg_sq, P.adam_state)]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up off the capture stream (kernel attributes, cuBLAS handles)
            for _ in range(2):
                self.forward_backward(self._static_in, self._static_gt, self._static_gamma, use_mask)
                self.optimizer_step(lr_groups=lr_groups)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for dst, src in zip((P.flat, P.exp_avg, P.exp_avg_sq, P.adam_state), saved):
            dst.copy_(src)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self.forward_backward(self._static_in, self._static_gt, self._static_gamma, use_mask)
            self.optimizer_step(reduce=True, lr_groups=lr_groups)
        self._graph = g1
        return self._graph

    def graph_step(self, model_input, gt, gamma=None, use_mask=False, lr_groups=None):
        """`step()` through a CUDA graph: (re)captures whenever something baked into the graph changes -- the mask
        switch, the rigidity window (training.py:203), the optimizer groups (training.py:300-313) or the batch
        size -- then copies the batch into the static buffers and replays."""
        rigid = self.opt.rigidity_loss_w is not None and self.step_index < self.opt.rigidity_loss_steps
        groups = None if lr_groups is None else tuple(sorted((k, v) for k, v in lr_groups.items()))
        key = (bool(use_mask), rigid, groups, int(model_input["coords"].numel()))
        if getattr(self, "_graph_key", None) != key:
            self.capture(model_input, gt, gamma, use_mask, lr_groups)
            self._graph_key = key
        self.load_resident(model_input, gt, gamma)
        return self.run_resident()

    def load_resident(self, model_input, gt, gamma=None, non_blocking=True):
        for k, dst in self._static_in.items():
            dst.copy_(model_input[k], non_blocking=non_blocking)
        self._static_gt["img"].copy_(gt["img"], non_blocking=non_blocking)
        if self._static_gamma is not None:
            self._static_gamma.copy_(torch.as_tensor(gamma, dtype=torch.float32).view(1), non_blocking=non_blocking)

    def run_resident(self):
        self._graph.replay()
        self.step_index += 1
        return self._losses

    def step(self, model_input, gt, gamma=None, use_mask=False, lr_groups=None):
        losses = self.forward_backward(model_input, gt, gamma, use_mask)
        self.optimizer_step(lr_groups=lr_groups)
        self.step_index += 1
        return losses
