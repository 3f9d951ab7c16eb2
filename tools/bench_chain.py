"""Kernel-level timing of the atlas kernels (CUDA events, L2 not flushed: inputs >> L2 at these sizes)."""
import math
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import _lib, ops
if os.environ.get("NEUCAM_CHAIN_DBG", "0") != "0":
    _lib.use_debug_library()   # the experiment switches live in the debug build only

def main():
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 382725
    iters = 10
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    w0 = ((torch.rand(512, 2, generator=g) * 2 - 1) * 0.5).to(dev)
    b0 = ((torch.rand(512, generator=g) * 2 - 1) * 0.7).to(dev)
    wh = ((torch.rand(3, 512, 512, generator=g) * 2 - 1) * math.sqrt(6 / 512) / 30).to(dev)
    bh = ((torch.rand(3, 512, generator=g) * 2 - 1) / math.sqrt(512)).to(dev)
    w4 = ((torch.rand(3, 512, generator=g) * 2 - 1) * math.sqrt(6 / 512) / 30).to(dev)
    b4 = torch.zeros(3, device=dev)
    w16 = wh.half().reshape(1536, 512).contiguous()
    wT16 = (30.0 * wh.transpose(1, 2)).half().reshape(1536, 512).contiguous()
    w4T16 = torch.empty(512, 64, dtype=torch.float16, device=dev)
    ops.pack_head_weights(w4, w4T16)
    uv = (torch.rand(Q, 2, device=dev) * 2 - 1)
    dy = torch.randn(Q, 3, device=dev) * 1e-5
    act, cos, dz = ops.alloc_chain_workspace(Q, dev)
    y = torch.empty(Q, 3, device=dev)
    duv = torch.empty(Q, 2, device=dev)
    scale = torch.tensor([2.0 ** 18], device=dev)
    dw0 = torch.zeros(512, 2, device=dev); db0 = torch.zeros(512, device=dev)
    dws = [torch.zeros(512, 512, device=dev) for _ in range(3)]
    dbs = [torch.zeros(512, device=dev) for _ in range(3)]
    dw4 = torch.zeros(3, 512, device=dev)

    ws = ops.new_ws(ops.WS_WGRAD, device=dev)     # >= every other reduction workspace used below
    fns = {
        "fwd(train)": lambda: ops.sine_mlp_fwd(uv, w0, b0, w16, bh, w4, b4, y=y, act16=act, phase16=cos),
        "fwd(infer)": lambda: ops.sine_mlp_fwd(uv, w0, b0, w16, bh, w4, b4, y=y),
        "bwd": lambda: ops.sine_mlp_bwd(uv, w0, b0, wT16, w4T16, dy, cos, scale, dz, dw0, db0, duv=duv, ws=ws),
        "wgrad": lambda: ops.sine_mlp_wgrad(dz, act, scale, dws, dbs, ws=ws),
        "head_wgrad": lambda: ops.sine_mlp_head_wgrad(cos, dy, dw4, ws=ws),
    }
    flops = {"fwd(train)": 2 * 788992, "fwd(infer)": 2 * 788992, "bwd": 2 * 788992, "wgrad": 2 * 3 * 512 * 512, "head_wgrad": 2 * 3 * 512}
    total = 0.0
    dbg = int(os.environ.get("NEUCAM_CHAIN_DBG", "0"))
    if dbg:
        _lib.lib().neucam_debug_set_chain_flags(dbg)
        print("debug flags", dbg)
    for name, fn in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if name != "fwd(infer)":
            total += ms
        print(f"{name:12s} {ms:8.3f} ms   {Q * flops[name] / ms / 1e9:8.1f} TFLOP/s")
    print(f"total fwd+bwd+wgrad {total:.3f} ms -> {Q / total / 1e3:.1f} M samples/s, "
          f"{Q * 4733952 / total / 1e9:.1f} TFLOP/s algorithmic")

if __name__ == "__main__":
    main()
