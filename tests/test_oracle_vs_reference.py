"""Live pin of the oracle's building blocks to the UNMODIFIED reference modules (runs whenever /root/reference is
present -- the build container; skipped on the GPU box, where the golden fixtures of tests/test_oracle_vs_golden.py
carry the same pin).  Each oracle function is compared on CPU fp32 with the reference class it restates, from the
reference's own state_dict."""
import pytest
import torch

from oracle_replay import orc

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


def test_sine_mlp_equals_SingleBVPNet():
    m = ref_shim.load("modules")
    torch.manual_seed(0)
    net = m.SingleBVPNet(type="sine", mode="mlp", hidden_features=64, num_hidden_layers=3, in_features=2,
                         out_features=3, sidelength=(30, 45), num_frequencies=7)
    x = torch.rand(1, 200, 2) * 2 - 1
    want = net({"coords": x})["model_out"]
    got = orc.sine_mlp(net.state_dict(), x[0])
    assert torch.allclose(got, want[0], atol=1e-6)


@pytest.mark.parametrize("kw,last", [
    (dict(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False, num_frequencies=7, skip_layer=[4],
          nonlinear="both"), "both"),
    (dict(hidden_dim=32, num_layer=6, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3, skip_layer=[4],
          nonlinear="tanh"), "tanh"),
    (dict(hidden_dim=32, num_layer=4, in_dim=3, out_dim=1, use_encoding=True, num_frequencies=4, skip_layer=[4],
          nonlinear="sigmoid"), "sigmoid"),
])
def test_simple_mlp_equals_SimpleMLP(kw, last):
    m = ref_shim.load("modules")
    torch.manual_seed(1)
    net = m.SimpleMLP(**kw)
    x = (torch.rand(150, 3) * 2 - 1).requires_grad_(True)
    want = net(x)
    nf = kw["num_frequencies"] if kw["use_encoding"] else None
    got = orc.simple_mlp(net.state_dict(), x, last, 3, skip_layer=(4,), num_frequencies=nf)
    assert torch.allclose(got, want, atol=1e-6)
    # the DETACHED skip connection (modules.py:287): input gradients agree too
    gw, = torch.autograd.grad(want.sum(), x, retain_graph=True)
    gg, = torch.autograd.grad(got.sum(), x)
    assert torch.allclose(gg, gw, atol=1e-6)


def test_tonemap_and_min_slope_equal_TonemapNet():
    m = ref_shim.load("modules")
    torch.manual_seed(2)
    tm = m.TonemapNet()
    x, e = torch.rand(300, 3) * 4 - 1, torch.rand(300, 1) * 6 - 3
    want = tm(x, e)
    got, _ = orc.tonemap(tm.state_dict(), x, e)
    assert torch.allclose(got, want, atol=1e-6)
    ldr, g = tm(x.clone().requires_grad_(True), e, output_grads=True)      # as training.py:248-251 calls it
    assert torch.allclose(orc.tonemap_min_slope(tm.state_dict(), x, e), g.detach(), atol=1e-9)


def test_neighbour_patch_and_gamma_map_equal_utils():
    u = ref_shim.load("utils")
    coords = torch.rand(1, 50, 3) * 2 - 1
    want = u.get_input_patch({"coords": coords}, (2, 30, 45), 9)["coords"]
    assert torch.allclose(orc.neighbour_patch(coords, 30, 45, 9), want, atol=1e-7)
    x = torch.rand(40, 3) * 2 - 1
    assert torch.allclose(orc.gamma_map(x, torch.tensor([37.0])), u.gamma_map(x, torch.tensor([37.0])), atol=1e-6)
