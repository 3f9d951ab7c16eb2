"""GPU test of the module-level drop-in: neucam_b200.modules.SingleBVPNet called like the reference's
(training.py:141, render_video.py) -- forward values, gradients w.r.t. every parameter AND w.r.t. the input
coordinates (modules.py:460-464 keeps the autograd link for 2-D coords) against the oracle."""
import pytest
import torch

from oracle_replay import orc, rel

from neucam_b200 import modules

pytestmark = pytest.mark.gpu


def test_single_bvp_net_forward_backward_like_reference():
    torch.manual_seed(5)
    net = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=512, num_hidden_layers=3, in_features=2,
                               out_features=3, sidelength=(30, 45), num_frequencies=7)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    K, B = 9, 333
    uv = (torch.rand(K, B, 2) * 2 - 1)
    wts = torch.rand(K, B, 1)
    # ours: coords produced by an upstream op so that the coordinate gradient has somewhere to flow
    src = uv.clone().cuda().requires_grad_(True)
    out = net({"coords": src * 1.0})
    assert out["model_out"].shape == (K, B, 3) and out["model_in"].shape == (K, B, 2)
    loss = (out["model_out"] * wts.cuda()).sum() * 1e-4
    loss.backward()
    torch.cuda.synchronize()
    # oracle
    src_r = uv.clone().requires_grad_(True)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y = orc.sine_mlp(sdr, src_r)
    ((y * wts).sum() * 1e-4).backward()
    assert rel(out["model_out"].detach().cpu(), y.detach()) <= 2e-3
    assert rel(src.grad.cpu(), src_r.grad) <= 1e-2
    for k, p in net.named_parameters():
        assert rel(p.grad.cpu(), sdr[k].grad) <= 1e-2, k


def test_inference_no_grad_and_4_outputs():
    torch.manual_seed(6)
    net = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=512, num_hidden_layers=3, in_features=2,
                               out_features=4, sidelength=(30, 45)).cuda()
    uv = (torch.rand(1, 1000, 2) * 2 - 1)
    with torch.no_grad():
        out = net({"coords": uv.cuda()})["model_out"]
    ref = orc.sine_mlp({k: v.detach().cpu() for k, v in net.state_dict().items()}, uv)
    assert out.shape == (1, 1000, 4)
    assert rel(out.cpu(), ref) <= 2e-3
