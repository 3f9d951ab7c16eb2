"""Gradient error of the fused 8-layer ReLU MLP against the oracle over a range of seeds (the test draws its seed from
Python's per-process string hash), and bit-reproducibility of repeated calls:  python tools/relu_seed_sweep.py [nseeds]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import modules
from oracle import neucam_oracle as orc

def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
Q = 1000
worst_all = []
for seed in range(n):
    torch.manual_seed(seed)
    net = modules.SimpleMLP(hidden_dim=256, num_layer=8, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3,
                            skip_layer=[4], nonlinear="tanh").cuda()
    x = (torch.rand(Q, 3, device="cuda") * 2 - 1)
    up = torch.randn(Q, 2, device="cuda") * 1e-4
    outs = []
    for rep in range(3):
        xin = x.clone().requires_grad_(True)
        for p in net.parameters():
            p.grad = None
        y = net(xin)
        (y * up).sum().backward()
        outs.append((y.detach().clone(), xin.grad.clone(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    same = all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and
               all(torch.equal(outs[0][2][k], o[2][k]) for k in o[2]) for o in outs[1:])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    xin = x.clone().requires_grad_(True)
    y0 = orc.simple_mlp(sd, xin, "tanh", 3, skip_layer=(4,), num_frequencies=3)
    (y0 * up).sum().backward()
    errs = {"y": rel(outs[0][0], y0.detach()), "dx": rel(outs[0][1], xin.grad)}
    for k, v in sd.items():
        errs[k] = rel(outs[0][2][k], v.grad)
    wk = max((k for k in errs if k != "y"), key=lambda k: errs[k])
    worst_all.append(errs[wk])
    print(f"seed {seed:2d} bit-identical x3: {same}  y {errs['y']:.2e}  worst grad {errs[wk]:.2e} ({wk})")
print("max over seeds", max(worst_all))
