"""Thin tensor-level wrappers over the C ABI (include/neucam_b200.h).  Every function takes CUDA
torch tensors, allocates outputs with torch, and enqueues on the current CUDA stream.  No fallback."""
import ctypes

import torch

from . import _lib

HID = 512
TILE_M = 128


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise _lib.NeucamNativeError("neucam_b200 ops need contiguous CUDA tensors (there is no CPU path)")


def num_tiles(Q):
    return (Q + TILE_M - 1) // TILE_M


def alloc_chain_workspace(Q, device):
    """(act16 [4,Q,512], phase16 [3, ntiles*128*512], dz16 [3,Q,512]) fp16 scratch for one atlas network."""
    nt = num_tiles(Q)
    act = torch.empty((4, Q, HID), dtype=torch.float16, device=device)
    cos = torch.empty((3, nt * TILE_M * HID), dtype=torch.float16, device=device)
    dz = torch.empty((3, Q, HID), dtype=torch.float16, device=device)
    return act, cos, dz


def sine_mlp_fwd(uv, w0, b0, w_hid16, b_hid, w4, b4, y=None, act16=None, phase16=None):
    """uv [Q,2] fp32 -> y [Q,out_dim] fp32.  w_hid16: fp16 [1536,512] (W1,W2,W3 stacked, [out,in])."""
    _chk_cuda(uv, w0, b0, w_hid16, b_hid, w4, b4, y, act16, phase16)
    Q = uv.shape[0]
    out_dim = w4.shape[0]
    if y is None:
        y = torch.empty((Q, out_dim), dtype=torch.float32, device=uv.device)
    rc = _lib.lib().neucam_sine_mlp_fwd(
        _lib.ptr(uv), ctypes.c_int64(Q), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(w_hid16), _lib.ptr(b_hid),
        _lib.ptr(w4), _lib.ptr(b4), out_dim, _lib.ptr(y), _lib.ptr(act16), _lib.ptr(phase16), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_fwd")
    return y


def pack_head_weights(w4, w4T16):
    """fp32 [out_dim,512] -> fp16 [512,64] zero-padded transposed operand of the head dgrad."""
    _chk_cuda(w4, w4T16)
    rc = _lib.lib().neucam_pack_head_weights(_lib.ptr(w4), w4.shape[0], _lib.ptr(w4T16), _lib.cur_stream())
    _lib.check(rc, "neucam_pack_head_weights")


def sine_mlp_bwd(uv, w0, b0, wT_hid16, w4T16, dy, phase16, scale, dz16, dw0, db0, duv=None):
    """dy [Q,out_dim] -> duv [Q,2]; accumulates dW0/db0; fills dz16 [3,Q,512] with scale * dL/dz_l, l=1..3.
    w4T16: fp16 [512,64] from pack_head_weights."""
    _chk_cuda(uv, w0, b0, wT_hid16, w4T16, dy, phase16, scale, dz16, duv, dw0, db0)
    Q = uv.shape[0]
    if duv is None:
        duv = torch.empty((Q, 2), dtype=torch.float32, device=uv.device)
    rc = _lib.lib().neucam_sine_mlp_bwd(
        _lib.ptr(uv), ctypes.c_int64(Q), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(wT_hid16), _lib.ptr(w4T16),
        dy.shape[1], _lib.ptr(dy), _lib.ptr(phase16), _lib.ptr(scale), _lib.ptr(dz16), _lib.ptr(duv),
        _lib.ptr(dw0), _lib.ptr(db0), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_bwd")
    return duv


def sine_mlp_wgrad(dz16, act16, scale, dws, dbs):
    """Accumulates hidden-layer gradients: dws/dbs = lists of 3 fp32 tensors ([512,512] / [512])."""
    _chk_cuda(dz16, act16, scale, *dws, *dbs)
    Q = act16.shape[1]
    rc = _lib.lib().neucam_sine_mlp_wgrad(
        _lib.ptr(dz16), _lib.ptr(act16), ctypes.c_int64(Q), _lib.ptr(scale), _lib.ptr(dws[0]), _lib.ptr(dbs[0]),
        _lib.ptr(dws[1]), _lib.ptr(dbs[1]), _lib.ptr(dws[2]), _lib.ptr(dbs[2]), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_wgrad")


def sine_mlp_head_wgrad(act16, dy, dw4):
    """Accumulates dW4 [out_dim,512] += dy^T a4."""
    _chk_cuda(act16, dy, dw4)
    Q = act16.shape[1]
    rc = _lib.lib().neucam_sine_mlp_head_wgrad(_lib.ptr(act16[3]), _lib.ptr(dy), dy.shape[1], ctypes.c_int64(Q),
                                               _lib.ptr(dw4), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_head_wgrad")


def pack_hidden_weights(w1, w2, w3, w16, wT16):
    """fp32 [512,512] x3 -> fp16 stacks (plain and transposed) for the forward / dgrad chains."""
    _chk_cuda(w1, w2, w3, w16, wT16)
    rc = _lib.lib().neucam_pack_hidden_weights(_lib.ptr(w1), _lib.ptr(w2), _lib.ptr(w3), _lib.ptr(w16),
                                               _lib.ptr(wT16), _lib.cur_stream())
    _lib.check(rc, "neucam_pack_hidden_weights")


SCALE_TARGET = 16.0


def loss_scale_from(dy):
    """Power-of-two loss scale 2^floor(log2(SCALE_TARGET / max|dy|)) as a device scalar (no host sync)."""
    _chk_cuda(dy)
    dy = dy.contiguous()
    out = torch.empty(2, dtype=torch.float32, device=dy.device)      # [scale, scratch bits]
    rc = _lib.lib().neucam_loss_scale(_lib.ptr(dy), ctypes.c_int64(dy.numel()), ctypes.c_float(SCALE_TARGET),
                                      _lib.ptr(out[1:]), _lib.ptr(out), _lib.cur_stream())
    _lib.check(rc, "neucam_loss_scale")
    return out[:1]


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def blur_fwd(coords, coord_idx, focus, blur_params, shape, K, offset_size, uv, kout, hid):
    """coords [B,3], coord_idx [B] int64 -> uv [K,B,2], kout [3K,B], hid [192,B]."""
    _chk_cuda(coords, coord_idx, focus, uv, kout, hid, *blur_params)
    N, H, W = shape
    rc = _lib.lib().neucam_blur_fwd(_lib.ptr(coords), _lib.ptr(coord_idx), _lib.ptr(focus), _ptr_array(blur_params),
                                    coords.shape[0], K, N, H, W, ctypes.c_float(offset_size), _lib.ptr(uv),
                                    _lib.ptr(kout), _lib.ptr(hid), _lib.cur_stream())
    _lib.check(rc, "neucam_blur_fwd")


def blur_bwd(coords, coord_idx, focus, blur_params, blur_grads, dfocus, shape, K, offset_size, kout, hid, dkw, duv):
    _chk_cuda(coords, coord_idx, focus, dfocus, kout, hid, dkw, duv, *blur_params, *blur_grads)
    N, H, W = shape
    rc = _lib.lib().neucam_blur_bwd(_lib.ptr(coords), _lib.ptr(coord_idx), _lib.ptr(focus), _ptr_array(blur_params),
                                    _ptr_array(blur_grads), _lib.ptr(dfocus), coords.shape[0], K, N, H, W,
                                    ctypes.c_float(offset_size), _lib.ptr(kout), _lib.ptr(hid), _lib.ptr(dkw),
                                    _lib.ptr(duv), _lib.cur_stream())
    _lib.check(rc, "neucam_blur_bwd")


def psf_crf_loss(y, kout, coord_idx, exps, fixed_exps, gt, gamma_w, mask_w, tm_params, tm_grads, shape, K,
                 has_blur, gamma_dev, fixed_value, ue_weight, global_count, dy, dkw, dexps, db4, sums, absmax_bits,
                 scale_out, ldr_out=None):
    """gamma_dev: device scalar tensor with the step's random gamma, or None (no gamma augmentation)."""
    _chk_cuda(y, kout, coord_idx, exps, fixed_exps, gt, gamma_w, mask_w, dy, dkw, dexps, db4, sums, absmax_bits,
              scale_out, ldr_out, gamma_dev, *tm_params, *tm_grads)
    N, H, W = shape
    rc = _lib.lib().neucam_psf_crf_loss(
        _lib.ptr(y), _lib.ptr(kout), _lib.ptr(coord_idx), _lib.ptr(exps), _lib.ptr(fixed_exps), _lib.ptr(gt),
        _lib.ptr(gamma_w), _lib.ptr(mask_w), _ptr_array(tm_params), _ptr_array(tm_grads), gt.shape[0], K, N, H, W,
        int(has_blur), _lib.ptr(gamma_dev), ctypes.c_float(fixed_value),
        ctypes.c_float(ue_weight), ctypes.c_int64(global_count), _lib.ptr(dy), _lib.ptr(dkw), _lib.ptr(dexps),
        _lib.ptr(db4), _lib.ptr(sums), _lib.ptr(absmax_bits), ctypes.c_float(SCALE_TARGET), _lib.ptr(scale_out), _lib.ptr(ldr_out),
        _lib.cur_stream())
    _lib.check(rc, "neucam_psf_crf_loss")


def adam_tick(state, beta1=0.9, beta2=0.999):
    _chk_cuda(state)
    rc = _lib.lib().neucam_adam_tick(_lib.ptr(state), ctypes.c_float(beta1), ctypes.c_float(beta2), _lib.cur_stream())
    _lib.check(rc, "neucam_adam_tick")


def adam_step(p, g, m, v, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    _chk_cuda(p, g, m, v, state)
    rc = _lib.lib().neucam_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), ctypes.c_int64(p.numel()),
                                     _lib.ptr(state), ctypes.c_float(lr), ctypes.c_float(beta1), ctypes.c_float(beta2),
                                     ctypes.c_float(eps), ctypes.c_float(grad_scale), _lib.cur_stream())
    _lib.check(rc, "neucam_adam_step")


def sample_batch(data, gamma_w_all, mask_w_all, idx_in, shape, seed, step, coords, coord_idx, img, gamma_w, mask_w):
    """On-device sampler / gather for a resident video (see neucam_sample_batch in the header)."""
    _chk_cuda(data, gamma_w_all, mask_w_all, idx_in, coords, coord_idx, img, gamma_w, mask_w)
    N, H, W = shape
    rc = _lib.lib().neucam_sample_batch(
        _lib.ptr(data), _lib.ptr(gamma_w_all), _lib.ptr(mask_w_all), _lib.ptr(idx_in), N, H, W, coords.shape[0],
        ctypes.c_uint64(seed), ctypes.c_uint64(step), _lib.ptr(coords), _lib.ptr(coord_idx), _lib.ptr(img),
        _lib.ptr(gamma_w), _lib.ptr(mask_w), _lib.cur_stream())
    _lib.check(rc, "neucam_sample_batch")


def tm_min_slope(tm_params, rad, expo):
    """min over points of d mean(ldr)/d lnx (value-only grad_loss input, training.py:248-252) -> device scalar."""
    _chk_cuda(rad, expo, *tm_params)
    bits = torch.empty(1, dtype=torch.int32, device=rad.device)
    rc = _lib.lib().neucam_tm_min_slope(_ptr_array(tm_params), _lib.ptr(rad), _lib.ptr(expo), rad.shape[0],
                                        _lib.ptr(bits), _lib.cur_stream())
    _lib.check(rc, "neucam_tm_min_slope")
    b = bits.to(torch.int64) & 0xFFFFFFFF
    dec = torch.where((b & 0x80000000) != 0, b & 0x7FFFFFFF, (~b) & 0xFFFFFFFF)
    return dec.to(torch.int32).view(torch.float32).reshape(())
