"""The 256-wide ReLU MLPs of the layered configs (uv-mapping and alpha networks) on the tcgen05 chain kernels.

Reference: `SimpleMLP` (modules.py:255-306), called at training.py:131-137 (uv_b, uv_f), 148-150 (alpha) and
utils.py:418-419 (rigidity loss).  `fused_forward(module, x)` is what `neucam_b200.modules.SimpleMLP.forward`
runs for CUDA inputs when the network has the shipped shape; forward, the gradient w.r.t. the (encoded) input,
and every weight / bias gradient come from `csrc/relu_chain.cu` + `csrc/wgrad.cu` through the C ABI.

Encoded-input tile (`feat`, [Q,64], built here with torch elementwise ops so that autograd carries the gradient
back through the positional encoding to the coordinates):
    columns [0,E)    the reference's network input: raw (t,y,x), then the NeRF encoding (modules.py:567-580),
                     rounded to fp16 ("hi" part) ...
    columns [E,2E)   ... and the fp16 rounding residual ("lo" part) of every feature, as E more features that
                     multiply the same weight columns, so the first layer sees its input to ~2^-22
    columns [2E,64)  zero
The forward is split-fp16 throughout (see csrc/relu_chain.cu): weights go in as hi and lo stages too.
"""
import ctypes

import torch

from . import _lib, ops

WID = 256
FEAT = 64
OUT_KINDS = {"none": 0, "tanh": 1, "sigmoid": 2}
_KEEP_LAST_ACTIVATIONS = False      # diagnostics (tools/relu_diag.py): keep the saved activations of each call
_last_activations = []


def eligible(module):
    """The shapes the chain kernels execute (every uv / alpha network of the shipped configs)."""
    L = module.num_layer
    if module.in_dim != 3 or L < 3 or L > 9 or module.last_nonlinear not in OUT_KINDS:
        return False
    E = module.positional_encoding.out_dim if module.use_encoding else module.in_dim
    if 2 * E > FEAT:
        return False
    skips = [i for i in module.skip_layer if i < L]
    if any(i < 1 or i > L - 2 for i in skips):
        return False
    ws = [l.weight for l in module.layers]
    if ws[0].shape != (WID, E) or ws[-1].shape[1] != WID or ws[-1].shape[0] > 4:
        return False
    for i in range(1, L - 1):
        if ws[i].shape != (WID, WID + (E if i in skips else 0)):
            return False
    return True


def _encoded_block(w, E, with_lo=True):
    """[256,E] weight columns that multiply the encoded input -> [256,64] in the feat column layout (the same
    weights multiply the hi and, when with_lo, the lo half of the features)."""
    out = w.new_zeros((w.shape[0], FEAT))
    out[:, :E] = w
    if with_lo:
        out[:, E:2 * E] = w
    return out


def _split(w):
    """fp32 -> (hi, lo) with hi = fp16(w), lo = fp16(w - hi), both returned as fp32 tensors holding fp16 values."""
    hi = w.to(torch.float16).float()
    return hi, (w - hi).to(torch.float16).float()


def _k_blocks(w):
    """[256, 256] (N x K, K-major) -> its four [256,64] K-blocks stacked along rows."""
    return w.reshape(WID, 4, 64).permute(1, 0, 2).reshape(4 * WID, 64)


class _ReluMLPFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, meta, *params):
        skips, out_kind, E = meta
        L = len(params) // 2
        H = L - 1
        ws, bs = params[0::2], params[1::2]
        Q = feat.shape[0]
        dev = feat.device
        feat16 = feat.to(torch.float16).contiguous()
        # stage plan: every weight block as a hi stage (x hi and lo activations) and a lo stage (x hi activations)
        w0_hi, w0_lo = _split(ws[0])
        blocks = [_encoded_block(w0_hi, E), _encoded_block(w0_lo, E, with_lo=False)]
        stage_begin, a_src, a_src2 = [0, 2], [4, 4], [255, 255]
        for l in range(1, H):
            w_hi, w_lo = _split(ws[l])
            kb_hi, kb_lo = _k_blocks(w_hi[:, :WID]).view(4, WID, 64), _k_blocks(w_lo[:, :WID]).view(4, WID, 64)
            for kb in range(4):
                blocks += [kb_hi[kb], kb_lo[kb]]
                a_src += [kb, kb]
                a_src2 += [5 + kb, 255]
            if l in skips:
                blocks += [_encoded_block(w_hi[:, WID:], E), _encoded_block(w_lo[:, WID:], E, with_lo=False)]
                a_src += [4, 4]
                a_src2 += [255, 255]
            stage_begin.append(len(a_src))
        w_stages = torch.cat(blocks, 0).to(torch.float16).contiguous()
        bias = torch.stack(bs[:H]).contiguous()
        w_last, b_last = ws[H].contiguous(), bs[H].contiguous()
        out_dim = w_last.shape[0]
        need_grad = any(ctx.needs_input_grad)
        act16 = torch.empty((2, H, Q, WID), dtype=torch.float16, device=dev) if need_grad else None   # hi, lo
        y = torch.empty((Q, out_dim), dtype=torch.float32, device=dev)
        sb = (ctypes.c_int * len(stage_begin))(*stage_begin)
        src = (ctypes.c_uint8 * len(a_src))(*a_src)
        src2 = (ctypes.c_uint8 * len(a_src2))(*a_src2)
        rc = _lib.lib().neucam_relu_mlp_fwd(_lib.ptr(feat16), ctypes.c_int64(Q), _lib.ptr(w_stages), H, sb, src, src2,
                                            _lib.ptr(bias), _lib.ptr(w_last), _lib.ptr(b_last), out_dim, out_kind,
                                            _lib.ptr(y), _lib.ptr(act16[0] if need_grad else None),
                                            _lib.ptr(act16[1] if need_grad else None), _lib.cur_stream())
        _lib.check(rc, "neucam_relu_mlp_fwd")
        if need_grad:
            ctx.save_for_backward(feat16, act16, y, *ws)
            ctx.meta = meta
            if _KEEP_LAST_ACTIVATIONS:
                _last_activations.append(act16)
        return y

    @staticmethod
    def backward(ctx, dy):
        feat16, act16, y, *ws = ctx.saved_tensors
        skips, out_kind, E = ctx.meta
        H = len(ws) - 1
        Q = feat16.shape[0]
        dev = feat16.device
        dy = dy.contiguous().float()
        if out_kind == 1:
            dpre = dy * (1.0 - y * y)
        elif out_kind == 2:
            dpre = dy * y * (1.0 - y)
        else:
            dpre = dy
        dpre = dpre.contiguous()
        scale = ops.loss_scale_from(dpre)
        want_dfeat = ctx.needs_input_grad[0]
        def hi_lo_stages(wt):       # [256,256] transposed weight -> its K-blocks as (hi, lo) stage pairs
            hi, lo = _split(wt)
            return torch.stack((_k_blocks(hi).view(4, WID, 64), _k_blocks(lo).view(4, WID, 64)), 1).reshape(8 * WID, 64)

        blocks = [hi_lo_stages(ws[l][:, :WID].t()) for l in range(H - 1, 0, -1)]
        if want_dfeat:
            w0t = ws[0].new_zeros((WID, WID))
            w0t[:E] = ws[0].t()     # rows E..2E (the residual features) stay zero: they are constants
            blocks.append(hi_lo_stages(w0t))
        wT_stages = (torch.cat(blocks, 0) if blocks else ws[0].new_zeros((8 * WID, 64))).to(torch.float16).contiguous()
        w_last = ws[H].contiguous()
        out_dim = w_last.shape[0]
        dz16 = torch.empty((2, H, Q, WID), dtype=torch.float16, device=dev)   # hi, lo
        dfeat = torch.empty((Q, FEAT), dtype=torch.float32, device=dev) if want_dfeat else None
        L_ = _lib.lib()
        rc = L_.neucam_relu_mlp_bwd(_lib.ptr(dpre), ctypes.c_int64(Q), _lib.ptr(wT_stages), H, _lib.ptr(w_last),
                                    out_dim, _lib.ptr(act16[0]), _lib.ptr(scale), _lib.ptr(dz16[0]), _lib.ptr(dz16[1]),
                                    _lib.ptr(dfeat), _lib.cur_stream())
        _lib.check(rc, "neucam_relu_mlp_bwd")
        # weight gradients
        dw = [torch.zeros_like(w) for w in ws]
        db = [torch.zeros(w.shape[0], dtype=torch.float32, device=dev) for w in ws]
        enc = torch.zeros((1 + len(skips), WID, FEAT), dtype=torch.float32, device=dev)   # dw_first, dw_skip...
        skip_slot = {l: 1 + i for i, l in enumerate(sorted(skips))}
        arr = lambda items: (ctypes.c_void_p * len(items))(*[0 if t is None else t.data_ptr() for t in items])
        dw_arr = arr([None] + dw[1:H])
        db_arr = arr(db[:H])
        skip_arr = arr([enc[skip_slot[l]] if l in skip_slot else None for l in range(H)])
        ldw = (ctypes.c_int * H)(*[int(w.shape[1]) for w in ws[:H]])
        rc = L_.neucam_relu_mlp_wgrad(_lib.ptr(dz16[0]), _lib.ptr(dz16[1]), _lib.ptr(act16[0]), _lib.ptr(act16[1]),
                                      _lib.ptr(feat16), ctypes.c_int64(Q), H,
                                      _lib.ptr(scale), _lib.ptr(enc[0]), dw_arr, ldw, db_arr, skip_arr,
                                      _lib.cur_stream())
        _lib.check(rc, "neucam_relu_mlp_wgrad")
        rc = L_.neucam_relu_mlp_head_wgrad(_lib.ptr(act16[0, H - 1]), _lib.ptr(act16[1, H - 1]), _lib.ptr(dpre), out_dim,
                                           ctypes.c_int64(Q), _lib.ptr(dw[H]), _lib.cur_stream())
        _lib.check(rc, "neucam_relu_mlp_head_wgrad")
        db[H] = dpre.sum(0)
        fold = lambda e: e[:, :E] + e[:, E:2 * E]      # hi + lo halves of the features
        dw[0] = fold(enc[0])
        for l, slot in skip_slot.items():
            dw[l][:, WID:] = fold(enc[slot])
        grads = []
        for w_g, b_g in zip(dw, db):
            grads += [w_g, b_g]
        return (dfeat, None, *grads)


def fused_forward(module, x):
    """SimpleMLP.forward (modules.py:282-306) for a CUDA input through the chain kernels."""
    if module.use_encoding:
        E = module.positional_encoding.out_dim
        enc = module.positional_encoding(x).view(-1, E)
    else:
        E = module.in_dim
        enc = x.reshape(-1, E)
    enc = enc.float()
    hi = enc.detach().to(torch.float16).float()
    lo = enc.detach() - hi
    enc_hi = enc + (hi - enc.detach())          # value = the fp16-representable part, gradient = identity
    feat = torch.cat((enc_hi, lo, enc.new_zeros((enc.shape[0], FEAT - 2 * E))), 1)
    L = module.num_layer
    skips = tuple(i for i in module.skip_layer if i < L)
    params = []
    for layer in module.layers:
        params += [layer.weight, layer.bias]
    return _ReluMLPFunction.apply(feat, (skips, OUT_KINDS[module.last_nonlinear], E), *params)
