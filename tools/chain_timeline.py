"""Decodes the in-kernel timeline of CTA 0 of the chain kernels (neucam_debug_set_chain_trace)."""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import ops, _lib
_lib.use_debug_library()   # diagnostics live in the debug build only

def run(bwd, Q=148 * 128 * 4):
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
    act, ph, dz = ops.alloc_chain_workspace(Q, dev)
    y = torch.empty(Q, 3, device=dev)
    scale = torch.tensor([2.0 ** 18], device=dev)
    dw0 = torch.zeros(512, 2, device=dev); db0 = torch.zeros(512, device=dev)
    trace = torch.zeros(1152, dtype=torch.int64, device=dev)
    for it in range(2):
        ops.sine_mlp_fwd(uv, w0, b0, w16, bh, w4, b4, y=y, act16=act, phase16=ph)
    lib = _lib.lib()
    lib.neucam_debug_set_chain_flags(int(os.environ.get("NEUCAM_CHAIN_DBG", "0")))
    if not bwd:
        lib.neucam_debug_set_chain_trace(ctypes.c_void_p(trace.data_ptr()))
        ops.sine_mlp_fwd(uv, w0, b0, w16, bh, w4, b4, y=y, act16=act, phase16=ph)
    else:
        ops.sine_mlp_bwd(uv, w0, b0, wT16, w4T16, dy, ph, scale, dz, dw0, db0)
        lib.neucam_debug_set_chain_trace(ctypes.c_void_p(trace.data_ptr()))
        ops.sine_mlp_bwd(uv, w0, b0, wT16, w4T16, dy, ph, scale, dz, dw0, db0)
    lib.neucam_debug_set_chain_trace(ctypes.c_void_p(0))
    torch.cuda.synchronize()
    t = trace.cpu().view(6, 2, 3, 8, 4)
    t0 = int(t[t > 0].min())
    rel = lambda x: int(x) - t0 if int(x) > 0 else -1
    print("==== %s (ns since first stamp; CTA pair 0, tile 1 = steady state)" % ("BWD" if bwd else "FWD"))
    for tile_i in (0, 1):
        print(f"-- tile {tile_i}: prologue half0 {rel(t[1,tile_i,0,7,0])}..{rel(t[1,tile_i,0,7,1])}")
        for g in range(3):
            print(f" layer {g}")
            for c in range(4):
                m = t[0, tile_i, g, c]
                print(f"   MMA chunk {c}: start {rel(m[0])} first-A-ready {rel(m[1])} kb7-ready {rel(m[2])} issued {rel(m[3])}")
            print("   MMA waits at kb=0: local a_ready[0..3] " + str([rel(x) for x in t[0, tile_i, g, 4]]) +
                  "  peer[0..3] " + str([rel(x) for x in t[0, tile_i, g, 5]]))
            print("   peer relay of this layer's INPUT k-blocks 0..7: " + str([rel(x) for x in t[3, tile_i, g, 4]] + [rel(x) for x in t[3, tile_i, g, 5]]))
            names = ["A0", "A1", "B0", "B1", "A2", "A3"]
            for cta in (0, 1):
                for h in (0, 1):
                    row = []
                    for st in range(6):
                        e = t[3 * cta + 1 + h, tile_i, g, st]
                        row.append(f"{names[st]}[{rel(e[0])},{rel(e[1])},{rel(e[2])},{rel(e[3])}]")
                    print(f"   cta{cta} epi half{h}: " + " ".join(row))

if __name__ == "__main__":
    run(False)
    run(True)
