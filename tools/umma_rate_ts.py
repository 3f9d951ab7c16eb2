import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import _lib
_lib.use_debug_library()   # diagnostics live in the debug build only
lib = _lib.lib()
grid = 148
for name, extra in (("A in smem", 0), ("A in tmem", 1 << 10)):
    for N in (256, 128):
        out = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
        iters = 2000
        for _ in range(2):
            rc = lib.neucam_umma_rate(N, iters, 0, 4 | extra, grid, ctypes.c_void_p(out.data_ptr()), _lib.cur_stream())
            _lib.check(rc, "umma_rate")
        torch.cuda.synchronize()
        o = out.cpu().view(-1, 2)
        o = o[o[:, 1] > 0]
        print(f"{name} cta_group::1 M128 N{N}: issue {o[:,0].float().mean()/(iters*4):6.1f} cyc/MMA, complete {o[:,1].float().mean()/(iters*4):6.1f} cyc/MMA")
