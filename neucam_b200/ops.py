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
    """(act16 [4,Q,512], cos16 [3, ntiles*128*512], dz16 [3,Q,512]) fp16 scratch for one atlas network."""
    nt = num_tiles(Q)
    act = torch.empty((4, Q, HID), dtype=torch.float16, device=device)
    cos = torch.empty((3, nt * TILE_M * HID), dtype=torch.float16, device=device)
    dz = torch.empty((3, Q, HID), dtype=torch.float16, device=device)
    return act, cos, dz


def sine_mlp_fwd(uv, w0, b0, w_hid16, b_hid, w4, b4, y=None, act16=None, cos16=None):
    """uv [Q,2] fp32 -> y [Q,out_dim] fp32.  w_hid16: fp16 [1536,512] (W1,W2,W3 stacked, [out,in])."""
    _chk_cuda(uv, w0, b0, w_hid16, b_hid, w4, b4, y, act16, cos16)
    Q = uv.shape[0]
    out_dim = w4.shape[0]
    if y is None:
        y = torch.empty((Q, out_dim), dtype=torch.float32, device=uv.device)
    rc = _lib.lib().neucam_sine_mlp_fwd(
        _lib.ptr(uv), ctypes.c_int64(Q), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(w_hid16), _lib.ptr(b_hid),
        _lib.ptr(w4), _lib.ptr(b4), out_dim, _lib.ptr(y), _lib.ptr(act16), _lib.ptr(cos16), _lib.cur_stream())
    _lib.check(rc, "neucam_sine_mlp_fwd")
    return y


def sine_mlp_bwd(uv, w0, b0, wT_hid16, w4, dy, cos16, scale, dz16, dw0, db0, duv=None):
    """dy [Q,out_dim] -> duv [Q,2]; accumulates dW0/db0; fills dz16 [3,Q,512] with scale * dL/dz_l, l=1..3."""
    _chk_cuda(uv, w0, b0, wT_hid16, w4, dy, cos16, scale, dz16, duv, dw0, db0)
    Q = uv.shape[0]
    if duv is None:
        duv = torch.empty((Q, 2), dtype=torch.float32, device=uv.device)
    rc = _lib.lib().neucam_sine_mlp_bwd(
        _lib.ptr(uv), ctypes.c_int64(Q), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(wT_hid16), _lib.ptr(w4),
        w4.shape[0], _lib.ptr(dy), _lib.ptr(cos16), _lib.ptr(scale), _lib.ptr(dz16), _lib.ptr(duv),
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
