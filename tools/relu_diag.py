"""Replays the two uv_b calls of a layered step (pixels, and the rigidity loss's shifted pixels) through the fused
and the plain-torch SimpleMLP with the SAME upstream gradients; prints per-call and summed gradient errors, the
number of ReLU gates that differ, and the forward error.  Diagnostic; GPU."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_layered_cpu as T  # noqa: E402
import test_layered_gpu as G  # noqa: E402
from neucam_b200 import modules, relu_mlp  # noqa: E402
from neucam_b200.layered import LayeredTrainStep  # noqa: E402
from oracle_replay import rel  # noqa: E402

kind, slot = sys.argv[1], sys.argv[2]
fused_ok = relu_mlp.eligible
relu_mlp.eligible = lambda m: False
modules.FCBlock.forward = T._torch_sine_block
models, state, opts, batch, cfg = G._setup(kind, 700)
step = LayeredTrainStep(models, G.SHAPES[kind], opts, device="cuda")
net = models[slot]
calls = []
orig = modules.SimpleMLP.forward


def fwd(self, x, _o=orig):
    y = _o(self, x)
    if self is net and y.requires_grad:
        rec = {"x": x.detach().clone()}
        y.register_hook(lambda g, rec=rec: rec.__setitem__("dy", g.detach().clone()))
        calls.append(rec)
    return y


modules.SimpleMLP.forward = fwd
step.forward_backward(batch, {"img": batch["img"]})
modules.SimpleMLP.forward = orig
total = {}
for mode in ("torch", "fused"):
    relu_mlp.eligible = fused_ok if mode == "fused" else (lambda m: False)
    total[mode] = []
    for rec in calls:
        for p in net.parameters():
            p.grad = None
        y = net(rec["x"])
        (y * rec["dy"]).sum().backward()
        total[mode].append((y.detach(), {k: p.grad.clone() for k, p in net.named_parameters()}))
# ReLU gates that differ between the fused forward (saved fp16 activations) and a plain fp32 forward, and how
# close to zero the reference pre-activation is there
relu_mlp.eligible = fused_ok
relu_mlp._KEEP_LAST_ACTIVATIONS = True
for i, rec in enumerate(calls):
    relu_mlp._last_activations.clear()
    net(rec["x"]).sum().backward()
    act = relu_mlp._last_activations[0][0]          # [H,Q,256] hi parts
    x = rec["x"].reshape(-1, 3)
    if net.use_encoding:
        x = net.positional_encoding(x).view(-1, net.positional_encoding.out_dim)
    skip_in = x
    for l, layer in enumerate(list(net.layers)[:-1]):
        if l in net.skip_layer:
            x = torch.cat((x, skip_in), -1)
        z64 = torch.nn.functional.linear(x.double(), layer.weight.double(), layer.bias.double())
        z = torch.nn.functional.linear(x, layer.weight, layer.bias)
        diff = (act[l] > 0) != (z > 0)
        diff64 = (z64 > 0) != (z > 0)
        margin = (z64.abs()[diff] / z64.abs().mean()).tolist()
        print(f"call {i} layer {l}: gates differing fused-vs-torch32 {int(diff.sum())} (relative |z| there: "
              f"{', '.join('%.1e' % m for m in margin[:6])});  torch32-vs-fp64 {int(diff64.sum())}")
        x = torch.relu(z)
names = [k for k, _ in net.named_parameters()]
for i, rec in enumerate(calls):
    yt, gt = total["torch"][i]
    yf, gf = total["fused"][i]
    print(f"call {i}: x {tuple(rec['x'].shape)} |dy| max {rec['dy'].abs().max():.2e}  y err {rel(yf, yt):.2e}  ",
          " ".join(f"{k.replace('layers.', 'L')} {rel(gf[k], gt[k]):.1e}" for k in names))
sum_t = {k: sum(g[k] for _, g in total["torch"]) for k in names}
sum_f = {k: sum(g[k] for _, g in total["fused"]) for k in names}
print("sum:   ", " ".join(f"{k.replace('layers.', 'L')} {rel(sum_f[k], sum_t[k]):.1e}" for k in names))
print("cancellation |sum| / sum|.|:", " ".join(
    f"{k.replace('layers.', 'L')} {(sum_t[k].norm() / sum(g[k].norm() for _, g in total['torch'])).item():.1e}" for k in names))
