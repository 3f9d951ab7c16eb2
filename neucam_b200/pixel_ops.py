"""autograd bindings of the per-pixel kernels (csrc/pixel.cu) for steps that are assembled with autograd
(neucam_b200.layered.LayeredTrainStep).  FusedTrainStep calls the same kernels directly on its static buffers.

  blur_taps        utils.get_input_patch + focus gather + blur-kernel SimpleMLP + defocus offsets + centre-tap reset
                   (training.py:99-121)  ->  tap positions [K,B,2] and PSF weights [K,B]
  psf_crf_loss     PSF sum, exposure, TonemapNet, image_mse (+ random gamma, mask), ue_loss (training.py:157-188,
                   244-245): the kernel computes the loss AND every gradient in one pass; backward hands them out
"""
import ctypes

import torch

from . import _gradmode, _lib, ops


class _BlurTaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, coord_idx, focus, cfg, *blur_params):
        shape, K, offset_size = cfg
        B = coords.shape[0]
        dev = coords.device
        uv = torch.empty((K, B, 2), device=dev)
        kout = torch.empty((3 * K, B), device=dev)
        hid = torch.empty((192, B), device=dev)
        params = [p.contiguous() for p in blur_params]
        focus_flat = focus.reshape(-1).contiguous()
        ops.blur_fwd(coords, coord_idx, focus_flat, params, shape, K, float(offset_size), uv, kout, hid)
        ctx.save_for_backward(coords, coord_idx, focus_flat, kout, hid, *params)
        ctx.cfg = cfg
        ctx.focus_shape = focus.shape
        ctx.grad_targets = _gradmode.targets_of([focus, *blur_params])
        kw = kout[:K].clone()
        return uv, kw

    @staticmethod
    def backward(ctx, duv, dkw):
        coords, coord_idx, focus_flat, kout, hid, *params = ctx.saved_tensors
        shape, K, offset_size = ctx.cfg
        tg = ctx.grad_targets
        if tg is not None:
            # the step owns preallocated gradient buffers and the kernel ACCUMULATES: add straight into param.grad
            ops.blur_bwd(coords, coord_idx, focus_flat, params, list(tg[1:]), tg[0].view(-1), shape, K,
                         float(offset_size), kout, hid, dkw.contiguous(), duv.contiguous())
            return (None,) * (4 + len(params))
        sizes = [p.numel() for p in params] + [focus_flat.numel()]
        buf = torch.zeros(sum(sizes), device=coords.device)
        parts = torch.split(buf, sizes)
        grads = [g.view_as(p) for g, p in zip(parts[:-1], params)]
        ops.blur_bwd(coords, coord_idx, focus_flat, params, grads, parts[-1], shape, K, float(offset_size), kout, hid,
                     dkw.contiguous(), duv.contiguous())
        return (None, None, parts[-1].view(ctx.focus_shape), None, *grads)


def blur_taps(coords, coord_idx, focus, blur_module, shape, K, offset_size):
    """coords [B,3] fp32, coord_idx [B] int64 -> (tap (y,x) positions [K,B,2], PSF weights [K,B])."""
    params = []
    for layer in blur_module.layers:
        params += [layer.weight, layer.bias]
    return _BlurTaps.apply(coords, coord_idx, focus, (tuple(shape), int(K), float(offset_size)), *params)


class _PsfCrfLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, kw, exps, fixed_exps, cfg, coord_idx, gt, gamma_w, mask_w, gamma_dev, *tm_params):
        shape, K, fixed_value, ue_weight, global_pixels = cfg
        dev = y.device
        has_blur = kw is not None
        tm = [p.contiguous() for p in tm_params]
        n_exp = exps.numel() if exps is not None else 0
        # direct mode (the step owns preallocated gradient buffers, _gradmode): the kernel, which produces every gradient
        # in its one pass and ACCUMULATES, adds the CRF / exposure gradients straight into param.grad.  That bakes in
        # d(total)/d(this loss) = 1, which is how the steps that switch the mode on use it (layered.py: the term enters
        # `total` unweighted); backward then hands dy / dkw on unscaled as well.
        tg = _gradmode.targets_of(list(tm_params) + ([exps] if exps is not None else []))
        sizes = [p.numel() for p in tm] + [max(n_exp, 1), 4, 4, 1]     # tm grads | dexps | db4 | sums | absmax bits
        buf = torch.zeros(sum(sizes), device=dev)
        parts = torch.split(buf, sizes)
        tm_g = [g.view_as(p) for g, p in zip(parts[:len(tm)], tm)]
        dexps, db4, sums, absmax = parts[len(tm):]
        if tg is not None:
            tm_g = list(tg[:len(tm)])
            if exps is not None:
                dexps = tg[len(tm)].view(-1)
        dy = torch.empty_like(y)
        dkw = torch.empty_like(kw) if has_blur else None
        scale = torch.empty(1, device=dev)
        ops.psf_crf_loss(y, kw, coord_idx, exps.reshape(-1) if exps is not None else None, fixed_exps, gt, gamma_w,
                         mask_w, tm, tm_g, shape, K, has_blur, gamma_dev, float(fixed_value), float(ue_weight),
                         3 * int(global_pixels), dy, dkw, dexps if exps is not None else None, db4, sums,
                         absmax.view(torch.int32), scale)
        img = sums[0] / float(3 * int(global_pixels))
        ue = sums[1] * float(ue_weight)
        ctx.save_for_backward(dy, dkw, buf)
        ctx.layout = (sizes, [p.shape for p in tm], exps.shape if exps is not None else None)
        ctx.direct = tg is not None
        ctx.mark_non_differentiable(img, ue)
        return img + ue, img, ue

    @staticmethod
    def backward(ctx, g_total, _g_img, _g_ue):
        dy, dkw, buf = ctx.saved_tensors
        sizes, tm_shapes, exps_shape = ctx.layout
        if ctx.direct:
            return (dy, dkw) + (None,) * (8 + len(tm_shapes))
        parts = torch.split(buf * g_total, sizes)          # one launch scales every small gradient
        tm_g = [g.view(shape) for g, shape in zip(parts, tm_shapes)]
        d_exps = parts[len(tm_shapes)][:exps_shape.numel()].view(exps_shape) if exps_shape is not None else None
        return (dy * g_total, dkw * g_total if dkw is not None else None, d_exps, None, None, None, None, None, None,
                None, *tm_g)


def psf_crf_loss(y, kw, exps, fixed_exps, tm_module, shape, K, fixed_value, ue_weight, global_pixels, coord_idx, gt,
                 gamma_w=None, mask_w=None, gamma_dev=None):
    """y [K,B,3] (log radiance per tap), kw [K,B] PSF weights or None (K = 1).  Returns (img_loss + ue_loss with
    gradient, img_loss value, ue_loss value); img_loss is this rank's share of the global mean, ue_loss is weighted
    by ue_weight (1 / world size)."""
    return _PsfCrfLoss.apply(y, kw, exps, fixed_exps, (tuple(shape), int(K), fixed_value, ue_weight, global_pixels),
                             coord_idx, gt, gamma_w, mask_w, gamma_dev, *tm_module.parameters())


class _Rigidity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uv_ref, uv_moved, coef, ku, kv):
        B = uv_ref.shape[0]
        dev = uv_ref.device
        out = torch.zeros(1, device=dev)
        g_ref = torch.empty((B, 2), device=dev)
        g_moved = torch.empty((2, B, 2), device=dev)
        ws = ops.new_ws(ops.WS_RIGIDITY, B, device=dev)
        rc = _lib.lib().neucam_rigidity_loss(_lib.ptr(uv_ref), _lib.ptr(uv_moved), B, ctypes.c_float(ku),
                                             ctypes.c_float(kv), _lib.ptr(coef), _lib.ptr(out), _lib.ptr(g_ref),
                                             _lib.ptr(g_moved), *ops._ws_args(ws), _lib.cur_stream())
        _lib.check(rc, "neucam_rigidity_loss")
        ctx.save_for_backward(g_ref, g_moved)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        g_ref, g_moved = ctx.saved_tensors
        return g_ref * g, g_moved * g, None, None, None


def rigidity_loss(uv_ref, uv_moved, coef, shape, uv_mapping_scale):
    """coef * sum over the B pixels of |J^T J|_F + |(J^T J + 0.001 I)^-1|_F (utils.py:402-454); uv_ref [B,2] is the
    mapping at the pixels, uv_moved [2,B,2] at the pixels one row up / one column left; coef: device scalar [1]."""
    n, h, w = shape
    ku, kv = (h - 1) / 2.0 / uv_mapping_scale, (w - 1) / 2.0 / uv_mapping_scale
    return _Rigidity.apply(uv_ref.reshape(-1, 2).contiguous(), uv_moved.reshape(2, -1, 2).contiguous(),
                           coef.reshape(1).float().contiguous(), float(ku), float(kv))


class _LayerBlend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, back, front, alpha, coefs):
        K, B = alpha.shape
        c1, c2, c3 = coefs
        dev = back.device
        out = torch.empty_like(back)
        e2 = torch.empty((K, B), device=dev)
        sums = torch.zeros(3, device=dev)
        ws = ops.new_ws(ops.WS_BLEND, K * B, device=dev)
        rc = _lib.lib().neucam_layer_blend_fwd(_lib.ptr(back), _lib.ptr(front), _lib.ptr(alpha), K, B,
                                               ctypes.c_float(c1), ctypes.c_float(c2), ctypes.c_float(c3),
                                               _lib.ptr(out), _lib.ptr(e2), _lib.ptr(sums), *ops._ws_args(ws),
                                               _lib.cur_stream())
        _lib.check(rc, "neucam_layer_blend_fwd")
        ctx.save_for_backward(back, front, alpha, e2)
        ctx.coefs = coefs
        ctx.mark_non_differentiable(sums)
        return out, sums.sum(), sums

    @staticmethod
    def backward(ctx, d_out, g_reg, _):
        back, front, alpha, e2 = ctx.saved_tensors
        K, B = alpha.shape
        c1, c2, c3 = ctx.coefs
        d_back, d_front, d_alpha = torch.empty_like(back), torch.empty_like(front), torch.empty_like(alpha)
        rc = _lib.lib().neucam_layer_blend_bwd(_lib.ptr(d_out.contiguous()), _lib.ptr(back), _lib.ptr(front),
                                               _lib.ptr(alpha), _lib.ptr(e2), _lib.ptr(g_reg.reshape(1).contiguous()),
                                               K, B, ctypes.c_float(c1), ctypes.c_float(c2), ctypes.c_float(c3),
                                               _lib.ptr(d_back), _lib.ptr(d_front), _lib.ptr(d_alpha),
                                               _lib.cur_stream())
        _lib.check(rc, "neucam_layer_blend_bwd")
        return d_back, d_front, d_alpha, None


def layer_blend(back, front, alpha, c1, c2, c3):
    """(1 - alpha) back + alpha front per tap, and the sparsity / alpha regularisers of the centre tap's alpha
    (training.py:143-153, 220-238).  back, front [K,B,3]; alpha [K,B,1].  Returns (blend [K,B,3], the weighted sum
    sparse1 + sparse2 + alpha with gradient, their three values [3] without)."""
    K, B = back.shape[0], back.shape[1]
    return _LayerBlend.apply(back.contiguous(), front.contiguous(), alpha.reshape(K, B).contiguous(),
                             (float(c1), float(c2), float(c3)))
