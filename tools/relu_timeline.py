"""Event timeline of CTA 0 of the ReLU-chain forward (clock64 cycles): where a tile's time goes."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import modules, _lib
_lib.use_debug_library()   # diagnostics live in the debug build only
Q = 165888
net = modules.SimpleMLP(hidden_dim=256, num_layer=4, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3,
                        skip_layer=[4], nonlinear="tanh").cuda()
x = torch.rand(Q, 3, device="cuda") * 2 - 1
lib = _lib.lib()
with torch.no_grad():
    for _ in range(2):
        net(x)
    torch.cuda.synchronize()
    trace = torch.zeros(3, 520, dtype=torch.int64, device="cuda")
    lib.neucam_debug_set_relu_trace(ctypes.c_void_p(trace.data_ptr()))
    net(x)
    torch.cuda.synchronize()
    lib.neucam_debug_set_relu_trace(ctypes.c_void_p(0))
t = trace.cpu()
t0 = min(int(t[r, 2]) for r in range(3) if t[r, 0] > 0)
names = {0: "producer", 1: "mma", 2: "epilogue"}
ev = []
for r in range(3):
    n = int(t[r, 0])
    for i in range(n):
        ev.append((int(t[r, 2 + 2 * i]) - t0, names[r], int(t[r, 1 + 2 * i])))
ev.sort()
for c, who, e in ev[:110]:
    print(f"{c:9d}  {who:9s} {e}")
