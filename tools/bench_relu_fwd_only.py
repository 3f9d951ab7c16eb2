"""Forward-only (no stores) timing of the ReLU chain: isolates MMA + ring + epilogue from the activation stores."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import modules, _lib
_lib.use_debug_library()   # diagnostics live in the debug build only
Q = 165888
net = modules.SimpleMLP(hidden_dim=256, num_layer=4, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3,
                        skip_layer=[4], nonlinear="tanh").cuda()
x = torch.rand(Q, 3, device="cuda") * 2 - 1
from torch.profiler import ProfilerActivity, profile
NAMES = {0: "full", 1: "no operand write-back", 2: "no MMA issue", 4: "no weight loads", 8: "no accumulator drain",
         6: "no MMA, no loads", 9: "no drain, no write-back", 15: "control flow only"}
for flags in (0, 1, 2, 4, 8, 6, 9, 15):
    _lib.lib().neucam_debug_set_relu_flags(flags)
    with torch.no_grad():
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                net(x)
            torch.cuda.synchronize()
    for ev in prof.key_averages():
        if "relu_chain" in ev.key:
            print(f"forward without stores, {NAMES[flags]:28s}: {ev.device_time_total / ev.count:8.1f} us")
_lib.lib().neucam_debug_set_relu_flags(0)
