"""The fused training step: one forward + backward + Adam update of NeuCam's per-step pipeline
(reference: training.py:95-271) executed by the library's CUDA kernels.

    S2-S5   neucam_blur_fwd        neighbour patch, focus gather, blur-kernel MLP, defocus offsets
    S7      neucam_sine_mlp_fwd    scene sine-MLP on all K*B taps (tcgen05)
    S9-S12  neucam_psf_crf_loss    PSF sum, exposure, CRF, image_mse (+gamma/mask), ue_loss; and their backward
    S14     neucam_sine_mlp_bwd    dgrad chain + layer-0 wgrad (tcgen05)
            neucam_sine_mlp_wgrad  hidden-layer weight gradients (tcgen05 split-K), head wgrad
            neucam_blur_bwd        blur MLP / focus backward
            [all-reduce]           one NCCL sum over the flat gradient buffer when world_size > 1
            neucam_adam_*          fused Adam over the flat buffers + fp16 operand re-pack

All parameters of all models are re-homed as views into ONE flat fp32 buffer (same for gradients and
Adam moments) so the optimizer and the all-reduce see a single tensor; module attribute names and
state_dict keys are untouched, so utils.save_model-style checkpoints keep the reference's layout.
"""
import math
from collections import OrderedDict

import torch

from . import _lib, ops
from .modules import LearnExps, LearnFocal, SimpleMLP, SingleBVPNet, TonemapNet

# order in which training.py:30-49 hands the models to Adam
MODEL_ORDER = ("atlas_b", "tm", "uv_b", "uv_f", "blur", "atlas_f", "alpha", "exp", "focal", "noise", "depth")


class FlatParameters:
    """Flat fp32 parameter / gradient / Adam-moment storage shared by every model of the step."""

    ALIGN = 4  # elements; keeps every tensor 16-byte aligned

    def __init__(self, models, device):
        self.device = torch.device(device)
        self.entries = []  # (slot, name, param, offset, numel)
        self.slot_ranges = OrderedDict()
        off = 0
        for slot in MODEL_ORDER:
            mod = models.get(slot)
            if mod is None:
                continue
            start = off
            for name, p in mod.named_parameters():
                n = p.numel()
                self.entries.append((slot, name, p, off, n))
                off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.slot_ranges[slot] = (start, off)
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        # {step, 1-b1^t, 1-b2^t, -}
        self.adam_state = torch.zeros(4, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for slot, name, p, o, n in self.entries:
                self.flat[o:o + n].copy_(p.detach().reshape(-1).to(self.device, torch.float32))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.grad[o:o + n].view(p.shape)
        # non-parameter buffers of the modules follow their parameters to the device
        for slot in MODEL_ORDER:
            mod = models.get(slot)
            if mod is not None:
                for b in mod.buffers():
                    b.data = b.data.to(self.device)

    def zero_grad(self):
        self.grad.zero_()

    def reset_optimizer(self):
        """A fresh Adam (training.py:300-313 re-creates the optimizer at start_freeze_tm)."""
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.adam_state.zero_()


class StepOptions:
    """The fields of the reference's `opt` namespace (configs.py) that the hot path reads."""

    def __init__(self, patch_size=9, offset_size=5, use_random_gamma=False, fixed_value=0.0, lr=1e-4,
                 uv_mapping_scale=1.0, rigidity_loss_w=None, rigidity_loss_steps=0, sparsity_loss_w=None,
                 alpha_loss_w=None):
        self.patch_size = patch_size
        self.offset_size = offset_size
        self.use_random_gamma = use_random_gamma
        self.fixed_value = fixed_value
        self.lr = lr
        # layered (uv-mapping) configs only; a weight of None means the loss is switched off
        self.uv_mapping_scale = uv_mapping_scale
        self.rigidity_loss_w = rigidity_loss_w
        self.rigidity_loss_steps = rigidity_loss_steps
        self.sparsity_loss_w = sparsity_loss_w
        self.alpha_loss_w = alpha_loss_w

    @classmethod
    def from_opt(cls, opt):
        on = lambda flag, w: float(getattr(opt, w)) if getattr(opt, flag, False) else None
        return cls(patch_size=getattr(opt, "patch_size", 9), offset_size=getattr(opt, "offset_size", 5),
                   use_random_gamma=bool(getattr(opt, "use_random_gamma", False)),
                   fixed_value=float(getattr(opt, "fixed_value", 0.0)), lr=float(getattr(opt, "lr", 1e-4)),
                   uv_mapping_scale=float(getattr(opt, "uv_mapping_scale", 1.0)),
                   rigidity_loss_w=on("rigidity_loss", "rigidity_loss_w"),
                   rigidity_loss_steps=int(getattr(opt, "rigidity_loss_steps", 0) or 0),
                   sparsity_loss_w=on("sparsity_loss", "sparsity_loss_w"),
                   alpha_loss_w=on("alpha_loss", "alpha_loss_w"))


def _unsupported(what):
    raise NotImplementedError(
        f"neucam_b200 fused step: {what} is outside FusedTrainStep, which covers the static multi-focus / "
        "multi-exposure pipeline (atlas_b + blur + focal + exp + tm); the uv-mapping / two-layer configs run "
        "through neucam_b200.layered.LayeredTrainStep (neucam_b200.training.train picks it automatically)")


def flat_adam(P, lr, lr_groups=None):
    """One Adam update of the flat buffers `P` (FlatParameters).  lr_groups: optional {slot: lr or None}; None
    freezes the slot (the post-freeze optimizer of training.py:300-313 drops tm / exp / focal and gives the blur
    net its own rate).  Default: one launch over the whole buffer (all groups identical, training.py:51)."""
    ops.adam_tick(P.adam_state)
    if lr_groups is None:
        ops.adam_step(P.flat, P.grad, P.exp_avg, P.exp_avg_sq, P.adam_state, lr)
        return
    for slot, (lo, hi) in P.slot_ranges.items():
        g_lr = lr_groups.get(slot, lr)
        if g_lr is None or hi == lo:
            continue
        ops.adam_step(P.flat[lo:hi], P.grad[lo:hi], P.exp_avg[lo:hi], P.exp_avg_sq[lo:hi], P.adam_state, g_lr)


class FusedTrainStep:
    """Forward + backward (+ optimizer) of one NeuCam training step on the current CUDA device.

    models: dict with the reference's slot names (train_video.py:157-169).  pixels_per_step = B on THIS
    rank.  With a process_group, gradients are summed across ranks and the loss is normalised by the
    global pixel count, so N ranks with B pixels each take the same step as one rank with N*B pixels.
    """

    def __init__(self, models, video_shape, options, pixels_per_step, device=None, process_group=None,
                 global_pixels=None):
        _lib.check(_lib.lib().neucam_check_device(), "neucam_check_device (need an sm_100 GPU)")
        self.models = models
        self.shape = tuple(int(s) for s in video_shape)
        self.opt = options
        self.B = int(pixels_per_step)
        self.device = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        # pixels of the GLOBAL batch (sum over ranks); differs from B * world only for ragged shards
        self.global_pixels = int(global_pixels) if global_pixels is not None else self.B * self.world

        for slot in ("uv_b", "uv_f", "atlas_f", "alpha", "noise", "depth"):
            if models.get(slot) is not None:
                _unsupported(f"model slot '{slot}'")
        for slot in ("atlas_b", "tm", "exp"):
            if models.get(slot) is None:
                _unsupported(f"a configuration without '{slot}'")
        atlas = models["atlas_b"]
        if not (isinstance(atlas, SingleBVPNet) and atlas.net.fused_eligible()):
            _unsupported("an atlas that is not the sine 2 -> 512 x4 -> 3 SingleBVPNet")
        self.has_blur = models.get("blur") is not None
        if self.has_blur:
            blur = models["blur"]
            ok = (isinstance(blur, SimpleMLP) and len(blur.layers) == 4 and not blur.use_encoding
                  and blur.layers[0].in_features == 3 and blur.layers[0].out_features == 64
                  and blur.layers[3].out_features == 27 and blur.last_nonlinear == "both"
                  and self.opt.patch_size == 9)
            if not ok:
                _unsupported("a blur model other than SimpleMLP(64, 4 layers, 3 -> 27, 'both') with patch_size 9")
            if models.get("focal") is None:
                _unsupported("a blur model without a focal model")
        self.K = 9 if self.has_blur else 1
        self.out_dim = atlas.net.dims[3]
        if self.out_dim != 3:
            _unsupported("an atlas_b with out_features != 3")

        self.params = FlatParameters(models, self.device)
        dev = self.device
        B, K = self.B, self.K
        Q = K * B
        self.Q = Q
        # static step inputs (H2D targets) -- fixed addresses so the step can be graph-captured
        self.coords = torch.zeros(B, 3, device=dev)
        self.coord_idx = torch.zeros(B, dtype=torch.int64, device=dev)
        self.img = torch.zeros(B, 3, device=dev)
        self.gamma_w = torch.zeros(B, device=dev)
        self.mask_w = torch.zeros(B, device=dev)
        self.gamma = torch.ones(1, device=dev)
        # workspaces
        self.uv = torch.empty(K, B, 2, device=dev)
        self.kout = torch.empty(27, B, device=dev)
        self.hid = torch.empty(192, B, device=dev)
        self.y = torch.empty(K, B, self.out_dim, device=dev)
        self.dy = torch.empty(K, B, self.out_dim, device=dev)
        self.dkw = torch.empty(K, B, device=dev)
        self.duv = torch.empty(K, B, 2, device=dev)
        self.act16, self.phase16, self.dz16 = ops.alloc_chain_workspace(Q, dev)
        self.w16 = torch.empty(3 * ops.HID, ops.HID, dtype=torch.float16, device=dev)
        self.wT16 = torch.empty(3 * ops.HID, ops.HID, dtype=torch.float16, device=dev)
        self.w4T16 = torch.empty(ops.HID, 64, dtype=torch.float16, device=dev)
        # scalars: sums[0..3] | absmax bits | scale
        self.scalars = torch.zeros(8, device=dev)
        self.sums = self.scalars[0:4]
        self.absmax_bits = self.scalars[4:5]
        self.scale = torch.ones(1, device=dev)
        self._side = torch.cuda.Stream(device=self.device)
        self._bind()
        self.repack()

    # -- parameter / gradient pointer tables -----------------------------------------------------------
    def _bind(self):
        m = self.models
        a = dict(m["atlas_b"].named_parameters())
        self.aw = [a[f"net.net.{i}.0.weight"] for i in range(5)]
        self.ab = [a[f"net.net.{i}.0.bias"] for i in range(5)]
        # biases of the three hidden layers must be one contiguous [3,512] block for the forward kernel;
        # in the flat buffer they are separated by the weight matrices, so keep a packed copy (repack()).
        self.b_hid = torch.empty(3, ops.HID, device=self.device)
        tm = dict(m["tm"].named_parameters())
        self.tm_p = []
        for ch in "rgb":
            self.tm_p += [tm[f"layers_{ch}.0.weight"], tm[f"layers_{ch}.0.bias"], tm[f"layers_{ch}.2.weight"],
                          tm[f"layers_{ch}.2.bias"]]
        self.tm_g = [p.grad for p in self.tm_p]
        self.exps = m["exp"].exps
        if self.has_blur:
            bl = dict(m["blur"].named_parameters())
            self.blur_p = []
            for i in range(4):
                self.blur_p += [bl[f"layers.{i}.weight"], bl[f"layers.{i}.bias"]]
            self.blur_g = [p.grad for p in self.blur_p]
            self.focus = m["focal"].focus

    def repack(self):
        """fp16 tensor-core operand copies of the atlas hidden weights (+ packed hidden biases)."""
        ops.pack_hidden_weights(self.aw[1].data, self.aw[2].data, self.aw[3].data, self.w16, self.wT16)
        ops.pack_head_weights(self.aw[4].data, self.w4T16)
        for i in range(3):
            self.b_hid[i].copy_(self.ab[i + 1].data)

    # -- data ----------------------------------------------------------------------------------------------
    def load_batch(self, model_input, gt, gamma=None, non_blocking=True):
        """Host (ideally pinned) batch in the reference's DataLoader layout ([1,B,...]) -> device buffers."""
        self.coords.copy_(model_input["coords"].reshape(-1, 3), non_blocking=non_blocking)
        self.coord_idx.copy_(model_input["coord_idx"].reshape(-1), non_blocking=non_blocking)
        self.img.copy_(gt["img"].reshape(-1, 3), non_blocking=non_blocking)
        if self.opt.use_random_gamma:
            self.gamma_w.copy_(model_input["gamma_weights"].reshape(-1), non_blocking=non_blocking)
            self.gamma.copy_(torch.as_tensor(gamma, dtype=torch.float32).reshape(1), non_blocking=non_blocking)
        if "weights" in model_input:
            self.mask_w.copy_(model_input["weights"].reshape(-1), non_blocking=non_blocking)

    def attach_dataset(self, data, gamma_weights=None, mask_weights=None):
        """Makes the video resident in HBM (data [N*H*W,3] = img*2-1, dataio.py:349) so that batches are formed on
        the device by sample_resident() instead of the CPU DataLoader."""
        dev = self.device
        self.ds_data = data.reshape(-1, 3).to(dev, torch.float32).contiguous()
        self.ds_gamma = None if gamma_weights is None else gamma_weights.reshape(-1).to(dev, torch.float32).contiguous()
        self.ds_mask = None if mask_weights is None else mask_weights.reshape(-1).to(dev, torch.float32).contiguous()
        self.idx_in = torch.zeros(self.B, dtype=torch.int64, device=dev)

    def sample_resident(self, seed=0, step=0, indices=None, gamma=None):
        """Forms the step's batch on the device: Philox-sampled (seed, step) or from injected `indices`
        (host or device int64 [B]; only 8 B/pixel cross the bus)."""
        idx = None
        if indices is not None:
            self.idx_in.copy_(indices.reshape(-1), non_blocking=True)
            idx = self.idx_in
        ops.sample_batch(self.ds_data, self.ds_gamma, self.ds_mask, idx, self.shape, seed, step, self.coords,
                         self.coord_idx, self.img, self.gamma_w if self.ds_gamma is not None else None,
                         self.mask_w if self.ds_mask is not None else None)
        if self.opt.use_random_gamma and gamma is not None:
            self.gamma.copy_(torch.as_tensor(gamma, dtype=torch.float32).reshape(1), non_blocking=True)

    def h2d_bytes(self):
        n = self.B * (3 * 4 + 8 + 3 * 4)
        if self.opt.use_random_gamma:
            n += self.B * 4 + 4
        return n

    # -- the step ------------------------------------------------------------------------------------------
    def forward_backward(self, use_mask=False, events=None):
        """Runs S2..S14 on the batch currently in the device buffers; gradients land in params.grad.
        events: optional list; the three tensor-core kernels are then bracketed by CUDA events on the launch stream
        and (name, start, end) tuples appended (bench.py: per-kernel durations inside real steps)."""
        o = self.opt

        def timed(name, fn):
            if events is None:
                return fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            events.append((name, e0, e1))
            return out

        P = self.params
        P.zero_grad()
        self.scalars.zero_()
        K, B = self.K, self.B
        if self.has_blur:
            ops.blur_fwd(self.coords, self.coord_idx, self.focus.data, [p.data for p in self.blur_p], self.shape, K,
                         float(o.offset_size), self.uv, self.kout, self.hid)
            uv = self.uv.view(-1, 2)
        else:
            uv = self.coords[:, 1:3].contiguous()
        timed("sine_chain_fwd", lambda: ops.sine_mlp_fwd(
            uv, self.aw[0].data, self.ab[0].data, self.w16, self.b_hid, self.aw[4].data, self.ab[4].data,
            y=self.y.view(-1, self.out_dim), act16=self.act16, phase16=self.phase16))
        ops.psf_crf_loss(self.y, self.kout if self.has_blur else None, self.coord_idx, self.exps.data, None, self.img,
                         self.gamma_w if o.use_random_gamma else None, self.mask_w if use_mask else None,
                         [p.data for p in self.tm_p], self.tm_g, self.shape, K, self.has_blur,
                         self.gamma if o.use_random_gamma else None, float(o.fixed_value), 1.0 / self.world,
                         3 * self.global_pixels, self.dy, self.dkw if self.has_blur else None, self.exps.grad,
                         self.ab[4].grad, self.sums, self.absmax_bits.view(torch.int32), self.scale)
        dyf = self.dy.view(-1, self.out_dim)
        timed("sine_chain_bwd", lambda: ops.sine_mlp_bwd(
            uv, self.aw[0].data, self.ab[0].data, self.wT16, self.w4T16, dyf, self.phase16, self.scale,
            self.dz16, self.aw[0].grad, self.ab[0].grad, duv=self.duv.view(-1, 2)))
        timed("hidden_wgrad", lambda: ops.sine_mlp_wgrad(
            self.dz16, self.act16, self.scale, [self.aw[i].grad for i in (1, 2, 3)],
            [self.ab[i].grad for i in (1, 2, 3)]))
        # the HBM-bound head wgrad and the FMA-bound blur backward are independent: run them side by side
        main = torch.cuda.current_stream(self.device)
        if self.has_blur:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                ops.blur_bwd(self.coords, self.coord_idx, self.focus.data, [p.data for p in self.blur_p], self.blur_g,
                             self.focus.grad, self.shape, K, float(o.offset_size), self.kout, self.hid, self.dkw,
                             self.duv)
        ops.sine_mlp_head_wgrad(self.act16, dyf, self.aw[4].grad)
        if self.has_blur:
            main.wait_stream(self._side)

    def reduce_gradients(self):
        """The step's only collective: SUM of the flat gradient buffer across the data-parallel ranks."""
        if self.world > 1:
            from .dist import allreduce_flat_gradients
            allreduce_flat_gradients(self.params.grad, self.pg)

    def optimizer_step(self, lr=None, reduce=True, lr_groups=None):
        """Gradient all-reduce (when distributed) + fused Adam + operand re-pack.

        lr_groups: optional {slot: lr or None}; None freezes the slot (the post-freeze optimizer of
        training.py:300-313 drops tm / exp / focal and gives the blur net its own rate).  Default: one launch
        over the whole flat buffer at `lr` (all groups identical, training.py:51)."""
        if reduce:
            self.reduce_gradients()
        flat_adam(self.params, self.opt.lr if lr is None else lr, lr_groups)
        self.repack()

    def grad_loss(self, rand_radiance=None, rand_exposure=None):
        """The reference's logged-only `grad_loss` = relu(-min slope of the CRF) * 1e6 over 10000 random points
        (training.py:248-252; drawn on the host like the reference when not given).  No gradient flows from it."""
        if rand_radiance is None:
            rand_radiance = torch.rand(10000, 3) * 5
            rand_exposure = torch.rand(10000, 1) * 6 - 3.0
        slope = ops.tm_min_slope([p.data for p in self.tm_p], rand_radiance.to(self.device).contiguous(),
                                 rand_exposure.to(self.device).contiguous())
        return torch.relu(-slope) * 1e6

    def losses(self):
        """Device scalars: img_loss (this rank's share of the global mean), ue_loss, total."""
        img = self.sums[0] / float(3 * self.global_pixels)
        ue = self.sums[1]
        return OrderedDict(img_loss=img, ue_loss=ue, total=img + ue)

    # -- CUDA graph ---------------------------------------------------------------------------------------
    def capture(self, use_mask=False, lr=None):
        """Captures forward_backward + optimizer_step into one CUDA graph (all kernel arguments are pointers
        to fixed buffers; the step's inputs are read from the static device buffers filled by load_batch, the
        step counter and loss scale live on the device).  `step()` replays it afterwards.  Re-capture after
        changing the learning rate or use_mask."""
        # one eager step first: sets kernel attributes and warms NCCL outside of capture
        saved = [t.clone() for t in (self.params.flat, self.params.exp_avg, self.params.exp_avg_sq,
                                     self.params.adam_state)]
        self.forward_backward(use_mask=use_mask)
        self.optimizer_step(lr)
        torch.cuda.synchronize(self.device)
        for dst, src in zip((self.params.flat, self.params.exp_avg, self.params.exp_avg_sq, self.params.adam_state),
                            saved):
            dst.copy_(src)
        self.repack()
        # the NCCL all-reduce stays outside the graphs (eager, same stream): graph 1 = forward + backward,
        # graph 2 = Adam + re-pack.  Single GPU: one graph.
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self.forward_backward(use_mask=use_mask)
            if self.world == 1:
                self.optimizer_step(lr, reduce=False)
        g2 = None
        if self.world > 1:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                self.optimizer_step(lr, reduce=False)
        self._graph = (g1, g2)
        self._graph_key = (bool(use_mask), lr)
        return self._graph

    def run_resident(self, use_mask=False):
        """One step on the batch already in the device buffers (graph replay when captured)."""
        g = getattr(self, "_graph", None)
        if g is not None and self._graph_key == (bool(use_mask), None):
            g[0].replay()
            if g[1] is not None:
                self.reduce_gradients()
                g[1].replay()
        else:
            self.forward_backward(use_mask=use_mask)
            self.optimizer_step()

    def step(self, model_input, gt, gamma=None, use_mask=False):
        self.load_batch(model_input, gt, gamma)
        self.run_resident(use_mask=use_mask)
        return self.losses()
