"""GPU parity of the fused 256-wide ReLU MLP (csrc/relu_chain.cu + the job-table wgrad kernel, through the C ABI and
neucam_b200.relu_mlp) against the ORACLE's restatement of SimpleMLP.forward (oracle.neucam_oracle.simple_mlp, which
follows modules.py:282-312 and is pinned to the unmodified reference by tests/test_oracle_vs_golden.py) in fp32 with
autograd, for every uv-mapping / alpha network shape the shipped configs build
(configs/config_{me_dynamic,video_deblur,video_hdr}.txt).
Tolerances: forward <= 1e-3, gradients <= 1e-2 relative (BASELINE.json north_star)."""
import zlib

import pytest
import torch

from oracle_replay import orc

from neucam_b200 import modules, relu_mlp

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


CASES = {
    "uv_b_plain": dict(num_layer=4, use_encoding=False, num_frequencies=None, nonlinear="tanh", out_dim=2),
    "uv_enc3": dict(num_layer=4, use_encoding=True, num_frequencies=3, nonlinear="tanh", out_dim=2),
    "alpha_enc4": dict(num_layer=4, use_encoding=True, num_frequencies=4, nonlinear="sigmoid", out_dim=1),
    "uv_f_hdr_skip": dict(num_layer=6, use_encoding=True, num_frequencies=3, nonlinear="tanh", out_dim=2),
    "uv_f_me_skip": dict(num_layer=8, use_encoding=True, num_frequencies=3, nonlinear="tanh", out_dim=2),
}


def run(net, x, upstream, fused, monkeypatch):
    """fused=True: the module's CUDA path (the chain kernels).  fused=False: the oracle on the module's state_dict."""
    xin = x.clone().requires_grad_(True)
    if fused:
        for p in net.parameters():
            p.grad = None
        y = net(xin)
        (y * upstream).sum().backward()
        return y.detach(), xin.grad.detach(), {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    nf = net.positional_encoding.num_frequencies if net.use_encoding else None
    y = orc.simple_mlp(sd, xin, net.last_nonlinear, 3, skip_layer=tuple(net.skip_layer), num_frequencies=nf)
    (y * upstream).sum().backward()
    return y.detach(), xin.grad.detach(), {k: v.grad.detach().clone() for k, v in sd.items()}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("Q", [1000, 128 * 300 + 5])
def test_forward_and_gradients_match_oracle(case, Q, monkeypatch):
    torch.manual_seed(zlib.crc32(case.encode()) % 1000)   # (str hash() is salted per process: the draw must not be)
    c = CASES[case]
    net = modules.SimpleMLP(hidden_dim=256, num_layer=c["num_layer"], in_dim=3, out_dim=c["out_dim"],
                            use_encoding=c["use_encoding"], num_frequencies=c["num_frequencies"], skip_layer=[4],
                            nonlinear=c["nonlinear"]).cuda()
    assert relu_mlp.eligible(net)
    x = (torch.rand(Q, 3, device="cuda") * 2 - 1)
    upstream = torch.randn(Q, c["out_dim"], device="cuda") * 1e-4      # small, like d loss / d uv
    y1, dx1, g1 = run(net, x, upstream, True, monkeypatch)
    y0, dx0, g0 = run(net, x, upstream, False, monkeypatch)
    assert rel(y1, y0) < 1e-3, rel(y1, y0)
    assert rel(dx1, dx0) < 1e-2, rel(dx1, dx0)
    for k in g0:
        assert rel(g1[k], g0[k]) < 1e-2, (k, rel(g1[k], g0[k]))


def test_patch_shaped_input_and_no_grad_path():
    """[K,B,3] input as at training.py:131 and a no-grad call (render path: nothing is stored)."""
    torch.manual_seed(1)
    net = modules.SimpleMLP(hidden_dim=256, num_layer=4, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3,
                            skip_layer=[4], nonlinear="tanh").cuda()
    x = torch.rand(9, 333, 3, device="cuda") * 2 - 1
    with torch.no_grad():
        y = net(x)
    assert y.shape == (9 * 333, 2)
    y2 = net(x.reshape(-1, 3))
    assert torch.equal(y, y2.detach())


@pytest.mark.parametrize("Q", [1, 127, 129, 128 * 3])
def test_tiny_and_odd_tile_counts(Q, monkeypatch):
    """The chain runs on CTA pairs: one tile (the peer CTA walks a phantom tile), an odd tile count, partial tiles."""
    torch.manual_seed(3)
    net = modules.SimpleMLP(hidden_dim=256, num_layer=4, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3,
                            skip_layer=[4], nonlinear="tanh").cuda()
    x = (torch.rand(Q, 3, device="cuda") * 2 - 1)
    upstream = torch.randn(Q, 2, device="cuda") * 1e-4
    y1, dx1, g1 = run(net, x, upstream, True, monkeypatch)
    y0, dx0, g0 = run(net, x, upstream, False, monkeypatch)
    assert rel(y1, y0) < 1e-3
    assert rel(dx1, dx0) < 1e-2
    for k in g0:
        assert rel(g1[k], g0[k]) < 1e-2, (k, rel(g1[k], g0[k]))


def test_blur_network_keeps_the_generic_path():
    net = modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                            num_frequencies=7, skip_layer=[4], nonlinear="both")
    assert not relu_mlp.eligible(net)


def test_split_backward_operands_option(monkeypatch):
    """relu_mlp.SPLIT_BACKWARD_OPERANDS carries dz and the saved activations as hi + lo pairs as well (three wgrad GEMMs
    per layer): per-call gradients then agree with fp32 torch to ~1e-6 instead of ~1e-5..6e-4."""
    torch.manual_seed(7)
    net = modules.SimpleMLP(hidden_dim=256, num_layer=6, in_dim=3, out_dim=2, use_encoding=True, num_frequencies=3,
                            skip_layer=[4], nonlinear="tanh").cuda()
    Q = 3000
    x = torch.rand(Q, 3, device="cuda") * 2 - 1
    upstream = torch.randn(Q, 2, device="cuda") * 1e-4
    y0, dx0, g0 = run(net, x, upstream, False, monkeypatch)
    errs = {}
    for split in (False, True):
        monkeypatch.setattr(relu_mlp, "SPLIT_BACKWARD_OPERANDS", split)
        y1, dx1, g1 = run(net, x, upstream, True, monkeypatch)
        errs[split] = max(rel(g1[k], g0[k]) for k in g0)
        assert rel(y1, y0) < 1e-5 and errs[split] < 1e-2 and rel(dx1, dx0) < 1e-2
    assert errs[True] <= errs[False]
