"""A few no_grad forwards of one uv network (target of `ncu -k regex:relu_chain`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import modules
Q = 165888
net = modules.SimpleMLP(hidden_dim=256, num_layer=4, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3, skip_layer=[4], nonlinear="tanh").cuda()
x = torch.rand(Q, 3, device="cuda") * 2 - 1
with torch.no_grad():
    for _ in range(4):
        net(x)
torch.cuda.synchronize()
