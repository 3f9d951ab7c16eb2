"""Times the ReLU-MLP kernels alone (CUDA events on the launch stream) at the video_deblur size.
    python tools/bench_relu.py [Q] [layers] [num_freq]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import modules  # noqa: E402

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 165888
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 3
net = modules.SimpleMLP(hidden_dim=256, num_layer=L, in_dim=3, out_dim=2, use_encoding=nf > 0, num_frequencies=nf or None,
                        skip_layer=[4], nonlinear="tanh").cuda()
x = (torch.rand(Q, 3, device="cuda") * 2 - 1).requires_grad_(True)
up = torch.randn(Q, 2, device="cuda") * 1e-4
from torch.profiler import ProfilerActivity, profile  # noqa: E402
for _ in range(3):
    net(x).backward(up)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        net(x).backward(up)
    torch.cuda.synchronize()
H = L - 1
flop_pass = 2 * Q * (256 * 64 + (H - 1) * 256 * 256)          # algorithmic (one product per weight)
for ev in prof.key_averages():
    if any(t in ev.key for t in ("relu_chain", "wgrad", "pack_kernel", "encode")):
        us = ev.device_time_total / ev.count
        extra = ""
        if "relu_chain" in ev.key or "hidden_wgrad" in ev.key:
            issued = 3 if "<false>" in ev.key or "ILb0" in ev.key else (2 if "relu_chain" in ev.key else 1)
            extra = f"  {flop_pass / us / 1e6:7.1f} TFLOP/s algorithmic, x{issued} issued"
        print(f"{ev.key[:70]:70s} {ev.count:3d} x {us:8.1f} us{extra}")
