"""The 256-wide ReLU MLPs of the layered configs (uv-mapping and alpha networks) on the tcgen05 chain kernels.

Reference: `SimpleMLP` (modules.py:255-306), called at training.py:131-137 (uv_b, uv_f), 148-150 (alpha) and
utils.py:418-419 (rigidity loss).  `fused_forward(module, x)` is what `neucam_b200.modules.SimpleMLP.forward`
runs for CUDA inputs when the network has the shipped shape; positional encoding, forward, the gradient w.r.t.
the sample position, and every weight / bias gradient come from `csrc/relu_chain.cu`, `csrc/relu_aux.cu` and
`csrc/wgrad.cu` through the C ABI.

Encoded-input tile (`feat16`, [Q,64] fp16, written by neucam_relu_mlp_encode):
    columns [0,E)    the reference's network input: raw (t,y,x), then the NeRF encoding (modules.py:567-580),
                     rounded to fp16 ("hi" part) ...
    columns [E,2E)   ... and the fp16 rounding residual ("lo" part) of every feature, as E more features that
                     multiply the same weight columns, so the first layer sees its input to ~2^-22
    columns [2E,64)  zero
The chain is split-fp16 throughout (see csrc/relu_chain.cu): weights go in as hi and lo stages too.
"""
import ctypes

import torch

from . import _gradmode, _lib, ops

WID = 256
FEAT = 64
OUT_KINDS = {"none": 0, "tanh": 1, "sigmoid": 2}
# Carry dz and the SAVED activations as hi + lo fp16 pairs through the backward chain and the weight-gradient GEMMs.
# Off: their rounding is random per sample and averages out over the Q samples of a weight gradient (measured: it did
# not move the rigidity-cancellation error at all, 1.7308e-2 -> 1.7257e-2), whereas the split of the WEIGHTS (forward
# and transposed) and of the forward's activations is what parity needs -- those stay on unconditionally.
SPLIT_BACKWARD_OPERANDS = False
_KEEP_LAST_ACTIVATIONS = False      # diagnostics (tools/relu_diag.py): keep the saved activations of each call
_last_activations = []


def eligible(module):
    """The shapes the chain kernels execute (every uv / alpha network of the shipped configs)."""
    L = module.num_layer
    if module.in_dim != 3 or L < 3 or L > 8 or module.last_nonlinear not in OUT_KINDS:
        return False
    E = module.positional_encoding.out_dim if module.use_encoding else module.in_dim
    if 2 * E > FEAT or (module.use_encoding and E != 3 + 6 * module.positional_encoding.num_frequencies):
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


def forward_plan(H, skips):
    """Stage plan of the forward chain: every weight K-block as a hi stage (multiplies the hi AND lo activation
    blocks) followed by a lo stage (multiplies the hi block).  Returns per stage (layer, kb | 4 = encoded-input
    block, lo flag), the operand sources a_src / a_src2, and stage_begin per GEMM."""
    st, a_src, a_src2, begin = [(0, 4, 0), (0, 4, 1)], [4, 4], [255, 255], [0, 2]
    for l in range(1, H):
        for kb in range(4):
            st += [(l, kb, 0), (l, kb, 1)]
            a_src += [kb, kb]
            a_src2 += [5 + kb, 255]
        if l in skips:
            st += [(l, 4, 0), (l, 4, 1)]
            a_src += [4, 4]
            a_src2 += [255, 255]
        begin.append(len(st))
    return st, a_src, a_src2, begin


def _u8(values):
    return (ctypes.c_uint8 * len(values))(*values)


class _ReluMLPFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, *params):
        skips, out_kind, nf, E, grad_enabled, ctx.grad_targets = meta
        L = len(params) // 2
        H = L - 1
        ws, bs = params[0::2], params[1::2]
        Q = x.shape[0]
        dev = x.device
        lib = _lib.lib()
        stream = _lib.cur_stream()
        # needs_input_grad ignores torch.no_grad(), and grad mode is always off inside forward(): the caller samples it
        need_grad = grad_enabled and any(ctx.needs_input_grad)
        want_dx = ctx.needs_input_grad[0]

        feat16 = torch.empty((Q, FEAT), dtype=torch.float16, device=dev)
        _lib.check(lib.neucam_relu_mlp_encode(_lib.ptr(x), ctypes.c_int64(Q), nf, _lib.ptr(feat16), stream),
                   "neucam_relu_mlp_encode")
        st, a_src, a_src2, begin = forward_plan(H, skips)
        S = len(st)
        n_bwd = 8 * ((H - 1) + (1 if want_dx else 0)) if need_grad else 0
        # the packed operand stages depend on the parameters only: inside one step (parameters constant) a network that is
        # evaluated again -- the rigidity loss's shifted pixels -- re-uses the first call's
        cache = _gradmode.step_cache()
        # (the backward stages of a call that also needs the input gradient are a superset, in the same order, of those
        #  of a call that does not: whichever was packed with more serves both)
        key = ("relu_pack", tuple(w.data_ptr() for w in ws), tuple(b.data_ptr() for b in bs), S, str(dev))
        packed = cache.get(key) if cache is not None else None
        if packed is not None and packed[3] < n_bwd:
            packed = None
        if packed is None:
            w_stages = torch.empty((S * WID, 64), dtype=torch.float16, device=dev)
            wT_stages = torch.empty((max(n_bwd, 1) * WID, 64), dtype=torch.float16, device=dev) if need_grad else None
            w_ptrs = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
            ld = (ctypes.c_int * L)(*[int(w.stride(0)) for w in ws])
            rc = lib.neucam_relu_mlp_pack(w_ptrs, ld, L, E, S, _u8([s[0] for s in st]), _u8([s[1] for s in st]),
                                          _u8([s[2] for s in st]), int(want_dx), _lib.ptr(w_stages),
                                          _lib.ptr(wT_stages), stream)
            _lib.check(rc, "neucam_relu_mlp_pack")
            bias = torch.stack(bs[:H]).contiguous()
            if cache is not None:
                cache[key] = (w_stages, wT_stages, bias, n_bwd)
        else:
            w_stages, wT_stages, bias = packed[:3]
        w_last, b_last = ws[H].contiguous(), bs[H].contiguous()
        out_dim = w_last.shape[0]
        n_parts = 2 if SPLIT_BACKWARD_OPERANDS else 1
        act16 = torch.empty((n_parts, H, Q, WID), dtype=torch.float16, device=dev) if need_grad else None   # hi[, lo]
        gates = torch.empty((H, Q, 8), dtype=torch.int32, device=dev) if need_grad else None
        y = torch.empty((Q, out_dim), dtype=torch.float32, device=dev)
        rc = lib.neucam_relu_mlp_fwd(_lib.ptr(feat16), ctypes.c_int64(Q), _lib.ptr(w_stages), H,
                                     (ctypes.c_int * len(begin))(*begin), _u8(a_src), _u8(a_src2), _lib.ptr(bias),
                                     _lib.ptr(w_last), _lib.ptr(b_last), out_dim, out_kind, _lib.ptr(y),
                                     _lib.ptr(act16[0] if need_grad else None),
                                     _lib.ptr(act16[1] if need_grad and n_parts == 2 else None), _lib.ptr(gates), stream)
        _lib.check(rc, "neucam_relu_mlp_fwd")
        if need_grad:
            ctx.save_for_backward(x, feat16, act16, y, wT_stages, gates, *ws)
            ctx.meta = meta
            if _KEEP_LAST_ACTIVATIONS:
                _last_activations.append(act16)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, feat16, act16, y, wT_stages, gates, *ws = ctx.saved_tensors
        skips, out_kind, nf, E = ctx.meta[:4]
        H = len(ws) - 1
        Q = feat16.shape[0]
        dev = feat16.device
        lib = _lib.lib()
        stream = _lib.cur_stream()
        dy = dy.contiguous().float()
        # per-block partial records + fixed-order second pass (no float atomics); one workspace serves every call below,
        # which run back to back on this stream
        red_ws = ops.new_ws(ops.WS_WGRAD, device=dev)
        tg = ctx.grad_targets
        db_last = tg[2 * H + 1] if tg is not None else torch.zeros(ws[H].shape[0], dtype=torch.float32, device=dev)
        # dpre = dy * d(out)/d(pre), db_last += dpre.sum(0), loss scale of max|dpre|: one kernel
        dpre, scale = ops.relu_head_prep(dy, y, out_kind, db_last, red_ws)
        want_dx = ctx.needs_input_grad[0]
        w_last = ws[H].contiguous()
        out_dim = w_last.shape[0]
        n_parts = act16.shape[0]
        dz16 = torch.empty((n_parts, H, Q, WID), dtype=torch.float16, device=dev)   # hi[, lo]
        lo = lambda t: t[1] if n_parts == 2 else None
        dfeat = torch.empty((Q, FEAT), dtype=torch.float32, device=dev) if want_dx else None
        rc = lib.neucam_relu_mlp_bwd(_lib.ptr(dpre), ctypes.c_int64(Q), _lib.ptr(wT_stages), H, _lib.ptr(w_last),
                                     out_dim, _lib.ptr(gates), _lib.ptr(scale), _lib.ptr(dz16[0]),
                                     _lib.ptr(lo(dz16)), _lib.ptr(dfeat), stream)
        _lib.check(rc, "neucam_relu_mlp_bwd")
        dx = None
        if want_dx:
            dx = torch.empty_like(x)
            _lib.check(lib.neucam_relu_mlp_encode_bwd(_lib.ptr(x), _lib.ptr(dfeat), ctypes.c_int64(Q), nf,
                                                      _lib.ptr(dx), stream), "neucam_relu_mlp_encode_bwd")
        # weight gradients.  The kernels ACCUMULATE: into param.grad directly when the step owns preallocated gradient
        # buffers (nothing is then returned for the parameters), else into one zero-filled temporary.
        skip_slot = {l: 1 + i for i, l in enumerate(sorted(skips))}
        enc_numel = (1 + len(skips)) * WID * FEAT
        if tg is not None:
            dw, db = list(tg[0::2]), list(tg[1::2])
            enc = torch.zeros(enc_numel, dtype=torch.float32, device=dev).view(1 + len(skips), WID, FEAT)
        else:
            sizes = [w.numel() for w in ws] + [w.shape[0] for w in ws]
            buf = torch.zeros(sum(sizes) + enc_numel, dtype=torch.float32, device=dev)
            parts = list(torch.split(buf, sizes + [enc_numel]))
            dw = [parts[i].view_as(ws[i]) for i in range(H + 1)]
            db = list(parts[H + 1:2 * (H + 1)])
            enc = parts[-1].view(1 + len(skips), WID, FEAT)                    # dw_first, dw_skip...
        arr = lambda items: (ctypes.c_void_p * len(items))(*[0 if t is None else t.data_ptr() for t in items])
        ldw = (ctypes.c_int * H)(*[int(w.shape[1]) for w in ws[:H]])
        rc = lib.neucam_relu_mlp_wgrad(_lib.ptr(dz16[0]), _lib.ptr(lo(dz16)), _lib.ptr(act16[0]), _lib.ptr(lo(act16)),
                                       _lib.ptr(feat16), ctypes.c_int64(Q), H, _lib.ptr(scale), _lib.ptr(enc[0]),
                                       arr([None] + dw[1:H]), ldw, arr(db[:H]),
                                       arr([enc[skip_slot[l]] if l in skip_slot else None for l in range(H)]),
                                       *ops._ws_args(red_ws), stream)
        _lib.check(rc, "neucam_relu_mlp_wgrad")
        rc = lib.neucam_relu_mlp_head_wgrad(_lib.ptr(act16[0, H - 1]),
                                            _lib.ptr(act16[1, H - 1] if n_parts == 2 else None), _lib.ptr(dpre),
                                            out_dim, ctypes.c_int64(Q), _lib.ptr(dw[H]), *ops._ws_args(red_ws), stream)
        _lib.check(rc, "neucam_relu_mlp_head_wgrad")
        fold = lambda e: e[:, :E] + e[:, E:2 * E]      # hi + lo halves of the features
        if tg is not None:
            dw[0].add_(fold(enc[0]))
            for l, slot in skip_slot.items():
                dw[l][:, WID:].add_(fold(enc[slot]))
            return (dx, None) + (None,) * (2 * (H + 1))
        db[H] = db_last
        dw[0] = fold(enc[0])
        for l, slot in skip_slot.items():
            dw[l][:, WID:] = fold(enc[slot])
        grads = []
        for w_g, b_g in zip(dw, db):
            grads += [w_g, b_g]
        return (dx, None, *grads)


def fused_forward(module, x):
    """SimpleMLP.forward (modules.py:282-306) for a CUDA input through the chain kernels."""
    nf = module.positional_encoding.num_frequencies if module.use_encoding else 0
    E = 3 + 6 * nf
    L = module.num_layer
    skips = tuple(i for i in module.skip_layer if i < L)
    params = []
    for layer in module.layers:
        params += [layer.weight, layer.bias]
    x3 = x.reshape(-1, 3).float().contiguous()
    meta = (skips, OUT_KINDS[module.last_nonlinear], nf, E, torch.is_grad_enabled(), _gradmode.targets_of(params))
    return _ReluMLPFunction.apply(x3, meta, *params)
