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


# reduction workspaces (include/neucam_b200.h: NEUCAM_WS_*): every cross-block reduction writes per-block partial
# records there and sums them in block order -- no float atomics, bit-reproducible results
WS_SINE_BWD, WS_WGRAD, WS_HEAD_WGRAD, WS_PSF_CRF, WS_BLUR_BWD, WS_RIGIDITY, WS_BLEND, WS_TM, WS_HEAD_PREP = range(9)


def reduce_ws_bytes(kind, n=0, aux=0):
    out = ctypes.c_int64(0)
    rc = _lib.lib().neucam_reduce_ws_bytes(int(kind), ctypes.c_int64(int(n)), int(aux), ctypes.byref(out))
    _lib.check(rc, "neucam_reduce_ws_bytes")
    return int(out.value)


def new_ws(kind, n=0, aux=0, device=None):
    """A fresh workspace tensor for one call (torch's caching allocator keeps it stream-ordered and graph-safe).
    Steps with static buffers allocate theirs once and pass them in."""
    return torch.empty(max(reduce_ws_bytes(kind, n, aux), 16), dtype=torch.uint8, device=device)


def _ws_args(ws):
    return _lib.ptr(ws), ctypes.c_int64(ws.numel() * ws.element_size())


def launch_count():
    """Kernels launched by the library so far in this process (graph replays re-run captured launches uncounted)."""
    out = ctypes.c_int64(0)
    _lib.check(_lib.lib().neucam_launch_count(ctypes.byref(out)), "neucam_launch_count")
    return int(out.value)


def alloc_chain_workspace(Q, device):
    """(act16 [3,Q,512] = a1..a3, phase16 [3, ntiles*128*512], dz16 [3,Q,512]) fp16 scratch for one atlas network."""
    nt = num_tiles(Q)
    act = torch.empty((3, Q, HID), dtype=torch.float16, device=device)
    cos = torch.empty((3, nt * TILE_M * HID), dtype=torch.float16, device=device)
    dz = torch.empty((3, Q, HID), dtype=torch.float16, device=device)
    return act, cos, dz


def activation_from_phase(phase16, layer, Q):
    """a_{layer+2} = sin(30 z_{layer+1}) as fp32 [Q,512] from the stored 16-bit phases (tile-blocked layout
    [tile][column group of 8][row][8]) -- what the head weight gradient regenerates in-kernel; used by tests."""
    nt = num_tiles(Q)
    p = phase16[layer].view(torch.int16).to(torch.int32) & 0xFFFF
    p = p.view(nt, HID // 8, TILE_M, 8).permute(0, 2, 1, 3).reshape(nt * TILE_M, HID)[:Q]
    return torch.sin(p.float() * (2.0 * 3.141592653589793 / 65536.0))


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


def sine_mlp_bwd(uv, w0, b0, wT_hid16, w4T16, dy, phase16, scale, dz16, dw0, db0, duv=None, ws=None):
    """dy [Q,out_dim] -> duv [Q,2]; accumulates dW0/db0; fills dz16 [3,Q,512] with scale * dL/dz_l, l=1..3.
    w4T16: fp16 [512,64] from pack_head_weights."""
    _chk_cuda(uv, w0, b0, wT_hid16, w4T16, dy, phase16, scale, dz16, duv, dw0, db0, ws)
    Q = uv.shape[0]
    if duv is None:
        duv = torch.empty((Q, 2), dtype=torch.float32, device=uv.device)
    if ws is None:
        ws = new_ws(WS_SINE_BWD, device=uv.device)
    rc = _lib.lib().neucam_sine_mlp_bwd(
        _lib.ptr(uv), ctypes.c_int64(Q), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(wT_hid16), _lib.ptr(w4T16),
        dy.shape[1], _lib.ptr(dy), _lib.ptr(phase16), _lib.ptr(scale), _lib.ptr(dz16), _lib.ptr(duv),
        _lib.ptr(dw0), _lib.ptr(db0), *_ws_args(ws), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_bwd")
    return duv


def sine_mlp_wgrad(dz16, act16, scale, dws, dbs, ws=None):
    """Accumulates hidden-layer gradients: dws/dbs = lists of 3 fp32 tensors ([512,512] / [512])."""
    _chk_cuda(dz16, act16, scale, ws, *dws, *dbs)
    Q = act16.shape[1]
    if ws is None:
        ws = new_ws(WS_WGRAD, device=dz16.device)
    rc = _lib.lib().neucam_sine_mlp_wgrad(
        _lib.ptr(dz16), _lib.ptr(act16), ctypes.c_int64(Q), _lib.ptr(scale), _lib.ptr(dws[0]), _lib.ptr(dbs[0]),
        _lib.ptr(dws[1]), _lib.ptr(dbs[1]), _lib.ptr(dws[2]), _lib.ptr(dbs[2]), *_ws_args(ws), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_wgrad")


def sine_mlp_head_wgrad(phase16, dy, dw4, ws=None):
    """Accumulates dW4 [out_dim,512] += dy^T a4, a4 = sin(30 z3) regenerated from phase16[2] (the forward does not store
    a4)."""
    _chk_cuda(phase16, dy, dw4, ws)
    Q = dy.shape[0]
    if ws is None:
        ws = new_ws(WS_HEAD_WGRAD, device=dy.device)
    rc = _lib.lib().neucam_sine_mlp_head_wgrad(_lib.ptr(phase16[2]), _lib.ptr(dy), dy.shape[1], ctypes.c_int64(Q),
                                               _lib.ptr(dw4), *_ws_args(ws), _lib.cur_stream())
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


def relu_head_prep(dy, y, out_kind, db_last, ws=None):
    """dpre = dy * d(out)/d(pre) (out_kind 0 none / 1 tanh / 2 sigmoid, from the forward output y), db_last += dpre.sum(0)
    and the power-of-two loss scale of max|dpre| in one pass; returns (dpre, scale).  out_kind 0: dpre is dy itself."""
    _chk_cuda(dy, db_last)
    Q, out_dim = dy.shape
    dpre = torch.empty_like(dy) if out_kind != 0 else dy
    out = torch.empty(2, dtype=torch.float32, device=dy.device)       # [scale, scratch bits]
    if ws is None:
        ws = new_ws(WS_HEAD_PREP, device=dy.device)
    rc = _lib.lib().neucam_relu_mlp_head_prep(_lib.ptr(dy), _lib.ptr(y if out_kind != 0 else None), out_dim,
                                              int(out_kind), ctypes.c_int64(Q), ctypes.c_float(SCALE_TARGET),
                                              _lib.ptr(dpre if out_kind != 0 else None), _lib.ptr(db_last),
                                              _lib.ptr(out[1:]), _lib.ptr(out), *_ws_args(ws), _lib.cur_stream())
    _lib.check(rc, "neucam_relu_mlp_head_prep")
    return dpre, out[:1]


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


TM_WIDTH = 128


def check_tm_shapes(tm_params):
    """The CRF kernels are built for TonemapNet(W=128) (modules.py:197-217; no config passes another width): per
    channel 0.weight [128,1], 0.bias [128], 2.weight [1,128], 2.bias [1].  Anything else would read and write out of
    bounds, so it is refused here AND by the C entry point (tm_width)."""
    if len(tm_params) != 12:
        raise NotImplementedError("TonemapNet parameters: expected 3 channels x (0.weight, 0.bias, 2.weight, 2.bias)")
    want = [(TM_WIDTH, 1), (TM_WIDTH,), (1, TM_WIDTH), (1,)]
    for i, p in enumerate(tm_params):
        if tuple(p.shape) != want[i % 4]:
            raise NotImplementedError(
                f"the fused CRF kernels execute TonemapNet(W={TM_WIDTH}) only; parameter {i} has shape {tuple(p.shape)}")
    return TM_WIDTH


def blur_fwd(coords, coord_idx, focus, blur_params, shape, K, offset_size, uv, kout, hid):
    """coords [B,3], coord_idx [B] int64 -> uv [K,B,2], kout [3K,B], hid [192,B]."""
    _chk_cuda(coords, coord_idx, focus, uv, kout, hid, *blur_params)
    N, H, W = shape
    rc = _lib.lib().neucam_blur_fwd(_lib.ptr(coords), _lib.ptr(coord_idx), _lib.ptr(focus), _ptr_array(blur_params),
                                    coords.shape[0], K, N, H, W, ctypes.c_float(offset_size), _lib.ptr(uv),
                                    _lib.ptr(kout), _lib.ptr(hid), _lib.cur_stream())
    _lib.check(rc, "neucam_blur_fwd")


def blur_bwd(coords, coord_idx, focus, blur_params, blur_grads, dfocus, shape, K, offset_size, kout, hid, dkw, duv,
             ws=None):
    _chk_cuda(coords, coord_idx, focus, dfocus, kout, hid, dkw, duv, ws, *blur_params, *blur_grads)
    N, H, W = shape
    if ws is None:
        ws = new_ws(WS_BLUR_BWD, coords.shape[0], N, device=coords.device)
    rc = _lib.lib().neucam_blur_bwd(_lib.ptr(coords), _lib.ptr(coord_idx), _lib.ptr(focus), _ptr_array(blur_params),
                                    _ptr_array(blur_grads), _lib.ptr(dfocus), coords.shape[0], K, N, H, W,
                                    ctypes.c_float(offset_size), _lib.ptr(kout), _lib.ptr(hid), _lib.ptr(dkw),
                                    _lib.ptr(duv), *_ws_args(ws), _lib.cur_stream())
    _lib.check(rc, "neucam_blur_bwd")


def psf_crf_loss(y, kout, coord_idx, exps, fixed_exps, gt, gamma_w, mask_w, tm_params, tm_grads, shape, K,
                 has_blur, gamma_dev, fixed_value, ue_weight, global_count, dy, dkw, dexps, db4, sums, absmax_bits,
                 scale_out, ldr_out=None, ws=None):
    """gamma_dev: device scalar tensor with the step's random gamma, or None (no gamma augmentation)."""
    _chk_cuda(y, kout, coord_idx, exps, fixed_exps, gt, gamma_w, mask_w, dy, dkw, dexps, db4, sums, absmax_bits,
              scale_out, ldr_out, gamma_dev, ws, *tm_params, *tm_grads)
    N, H, W = shape
    tm_width = check_tm_shapes(tm_params)
    if ws is None:
        ws = new_ws(WS_PSF_CRF, gt.shape[0], N, device=y.device)
    rc = _lib.lib().neucam_psf_crf_loss(
        _lib.ptr(y), _lib.ptr(kout), _lib.ptr(coord_idx), _lib.ptr(exps), _lib.ptr(fixed_exps), _lib.ptr(gt),
        _lib.ptr(gamma_w), _lib.ptr(mask_w), _ptr_array(tm_params), _ptr_array(tm_grads), gt.shape[0], K, N, H, W,
        int(has_blur), _lib.ptr(gamma_dev), ctypes.c_float(fixed_value),
        ctypes.c_float(ue_weight), ctypes.c_int64(global_count), _lib.ptr(dy), _lib.ptr(dkw), _lib.ptr(dexps),
        _lib.ptr(db4), _lib.ptr(sums), _lib.ptr(absmax_bits), ctypes.c_float(SCALE_TARGET), _lib.ptr(scale_out), _lib.ptr(ldr_out),
        tm_width, *_ws_args(ws), _lib.cur_stream())
    _lib.check(rc, "neucam_psf_crf_loss")


def tm_fwd(tm_params, x, expo):
    """TonemapNet.forward (modules.py:219-247, noise=None): x [M,3] log radiance, expo [M] stops -> ldr [M,3]."""
    _chk_cuda(x, expo, *tm_params)
    ldr = torch.empty_like(x)
    rc = _lib.lib().neucam_tm_fwd(_ptr_array(tm_params), check_tm_shapes(tm_params), _lib.ptr(x), _lib.ptr(expo),
                                  ctypes.c_int64(x.shape[0]), _lib.ptr(ldr), _lib.cur_stream())
    _lib.check(rc, "neucam_tm_fwd")
    return ldr


def tm_bwd(tm_params, x, expo, dldr, tm_grads, want_dexpo=True):
    """Backward of tm_fwd: returns (dx [M,3], dexpo [M] or None) and ACCUMULATES the 12 parameter gradients."""
    _chk_cuda(x, expo, dldr, *tm_params, *tm_grads)
    M = x.shape[0]
    dx = torch.empty_like(x)
    dexpo = torch.empty_like(expo) if want_dexpo else None
    ws = new_ws(WS_TM, M, device=x.device)
    rc = _lib.lib().neucam_tm_bwd(_ptr_array(tm_params), check_tm_shapes(tm_params), _lib.ptr(x), _lib.ptr(expo),
                                  _lib.ptr(dldr), ctypes.c_int64(M), _lib.ptr(dx), _lib.ptr(dexpo),
                                  _ptr_array(tm_grads), *_ws_args(ws), _lib.cur_stream())
    _lib.check(rc, "neucam_tm_bwd")
    return dx, dexpo


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


# ---- the whole static step from one C call (include/neucam_b200.h: neucam_step_args / neucam_step_fwd_bwd) ----------
_P = ctypes.c_void_p


class StepArgs(ctypes.Structure):
    """Mirror of `struct neucam_step_args` (field order and types must match the header)."""
    _fields_ = [
        ("struct_bytes", ctypes.c_uint32), ("phases", ctypes.c_int),
        ("B", ctypes.c_int), ("K", ctypes.c_int), ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
        ("out_dim", ctypes.c_int), ("tm_width", ctypes.c_int),
        ("offset_size", ctypes.c_float), ("fixed_value", ctypes.c_float), ("ue_weight", ctypes.c_float),
        ("global_count", ctypes.c_int64),
        ("coords", _P), ("coord_idx", _P), ("img", _P), ("gamma_w", _P), ("gamma_dev", _P), ("mask_w", _P),
        ("focus", _P), ("blur_params", _P * 8),
        ("w0", _P), ("b0", _P), ("b_hid", _P), ("w4", _P), ("b4", _P),
        ("w_hid16", _P), ("wT_hid16", _P), ("w4T16", _P),
        ("exps", _P), ("tm_params", _P * 12),
        ("g_focus", _P), ("g_blur", _P * 8),
        ("g_w0", _P), ("g_b0", _P), ("g_w", _P * 3), ("g_b", _P * 3), ("g_w4", _P), ("g_b4", _P),
        ("g_exps", _P), ("g_tm", _P * 12),
        ("zero_ptr", _P), ("zero_bytes", ctypes.c_int64),
        ("uv", _P), ("kout", _P), ("hid", _P), ("y", _P), ("dy", _P), ("dkw", _P), ("duv", _P),
        ("act16", _P), ("phase16", _P), ("dz16", _P),
        ("sums", _P), ("absmax_bits", _P), ("scale", _P),
        ("ws", _P), ("ws_bytes", ctypes.c_int64), ("ws_side", _P), ("ws_side_bytes", ctypes.c_int64),
    ]


def _dp(t):
    return None if t is None else t.data_ptr()


STEP_TO_HIDDEN_WGRAD, STEP_TAILS = 1, 2


def step_fwd_bwd(args, side_stream=None, phases=0):
    """Enqueues the static step (forward + backward, no optimizer) described by `args` (StepArgs) on the current
    stream; `side_stream` (torch.cuda.Stream or None) carries the blur backward next to the head wgrad.  phases: 0 =
    everything, or STEP_TO_HIDDEN_WGRAD / STEP_TAILS to run it in two parts (data-parallel overlap)."""
    side = ctypes.c_void_p(side_stream.cuda_stream) if side_stream is not None else ctypes.c_void_p(0)
    args.phases = int(phases)
    rc = _lib.lib().neucam_step_fwd_bwd(ctypes.byref(args), _lib.cur_stream(), side)
    _lib.check(rc, "neucam_step_fwd_bwd")
