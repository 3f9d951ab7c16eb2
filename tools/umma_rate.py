import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import _lib
_lib.use_debug_library()   # diagnostics live in the debug build only
lib = _lib.lib()
for grid, ovh in ((148, 0), (148, 1), (148, 2), (148, 3)):
    for two, N in ((0, 256), (1, 256)):
        out = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
        iters = 2000
        for _ in range(2):
            rc = lib.neucam_umma_rate(N, iters, two, 8 | (ovh << 8), grid, ctypes.c_void_p(out.data_ptr()), _lib.cur_stream())
            _lib.check(rc, "umma_rate")
        torch.cuda.synchronize()
        o = out.cpu().view(-1, 2)
        o = o[o[:, 1] > 0]
        n_mma = iters * 4
        M = 256 if two else 128
        ideal = M * N * 16 / (8192 if two else 4096)
        print(f"overhead {ovh} grid {grid:3d} cta_group::{2 if two else 1} M{M} N{N}: issue {o[:,0].float().mean()/n_mma:6.1f} cyc/MMA, "
              f"complete {o[:,1].float().mean()/n_mma:6.1f} cyc/MMA (ideal {ideal:.0f})")
