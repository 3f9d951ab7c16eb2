"""The warm-up loops around the step (reference utils.py:265-334, run by train_video.py:126-142) on the package's
kernels, against golden runs of the UNMODIFIED reference on the CPU (tests/golden/warmups.pt, written by
oracle/make_golden_aux.py): same seeds -> same CPU random draws -> same losses per iteration and the same warmed
parameters, up to the kernels' arithmetic (fp32 CRF kernel; split-fp16 ReLU chain)."""
import types

import pytest
import torch

from oracle_replay import load_fixture

from neucam_b200 import modules
from neucam_b200 import utils as nutils

pytestmark = pytest.mark.gpu


def _digest_ok(model, dg):
    for k, v in model.state_dict().items():
        assert float(v.double().sum()) == pytest.approx(dg[k][0], rel=1e-12, abs=1e-12), k
        assert float(v.double().abs().sum()) == pytest.approx(dg[k][1], rel=1e-12), k


def _compare(model, fx, loss_rel, travel, p_mean):
    """Losses per iteration, and the warmed parameters.  Adam's first steps move every weight by ~lr whatever the
    gradient's size, so a weight whose gradient is ~0 can legitimately go the other way: the bound on single entries is
    the distance Adam can travel (`travel` = 2 x steps x lr), while the bulk must agree to a fraction of one step:
    mean <= p_mean, and all but 3e-3 of the entries within 2e-5 -- the share of weights whose gradient is smaller than
    the kernels' ~1e-3 relative gradient error, i.e. whose Adam direction g / |g| is not determined at that accuracy
    (measured: 8e-4 of layers.1.weight of the alpha network, whose sigmoid output starts flat)."""
    got = model.warm_history.cpu()
    assert got.shape == fx["losses"].shape
    assert torch.allclose(got, fx["losses"], rtol=loss_rel, atol=1e-7), (got[:3], fx["losses"][:3])
    for k, want in fx["final"].items():
        mine = model.state_dict()[k].detach().cpu()
        d = (mine - want).abs()
        assert d.max().item() <= travel, (k, d.max().item())
        assert d.mean().item() <= p_mean, (k, d.mean().item())
        n_off = int((d > 2e-5).sum().item())
        assert n_off <= max(1, int(3e-3 * d.numel())), (k, n_off, d.numel())


def test_warm_tm_matches_reference_run():
    fx = load_fixture("warmups")["warm_tm"]
    torch.manual_seed(fx["model_seed"])
    tm = modules.TonemapNet()
    _digest_ok(tm, fx["init_digest"])                 # same initial parameters as the reference run
    torch.manual_seed(fx["rng_seed"])
    nutils.warm_tm(tm, pretrain_iters=fx["iters"], device="cuda")
    # 25 Adam steps of 5e-4: the CRF kernel is fp32, differences are summation order only
    _compare(tm, fx, loss_rel=1e-4, travel=2e-5, p_mean=2e-6)


def test_warm_mapping_matches_reference_run():
    fx = load_fixture("warmups")["warm_mapping_uv_enc_skip"]
    torch.manual_seed(fx["model_seed"])
    net = modules.SimpleMLP(**fx["kw"])
    _digest_ok(net, fx["init_digest"])
    torch.manual_seed(fx["rng_seed"])
    nutils.warm_mapping(net, fx["shape"], uv_mapping_scale=fx["scale"], pretrain_iters=fx["iters"], device="cuda")
    _compare(net, fx, loss_rel=1e-4, travel=2 * 6 * 1e-4, p_mean=2e-6)      # 2 iterations x 3 frames of lr 1e-4


def test_warm_alpha_matches_reference_run():
    fx = load_fixture("warmups")["warm_alpha"]
    torch.manual_seed(fx["model_seed"])
    net = modules.SimpleMLP(**fx["kw"])
    _digest_ok(net, fx["init_digest"])
    ds = types.SimpleNamespace(shape=fx["shape"], vid=fx["vid"].numpy())
    torch.manual_seed(fx["rng_seed"])
    nutils.warm_alpha(net, ds, pretrain_iters=fx["iters"], device="cuda")
    _compare(net, fx, loss_rel=1e-4, travel=2 * 6 * 1e-4, p_mean=2e-6)


def test_tonemap_kernels_match_the_torch_module():
    """neucam_tm_fwd / neucam_tm_bwd against autograd through the same module's torch path."""
    torch.manual_seed(3)
    tm = modules.TonemapNet().cuda()
    x = (torch.rand(1000, 3, device="cuda") * 4 - 1).requires_grad_(True)
    e = (torch.rand(1000, 1, device="cuda") * 6 - 3).requires_grad_(True)
    y = tm(x, e)                                   # kernel path
    w = torch.randn_like(y)
    (y * w).sum().backward()
    got = [x.grad.clone(), e.grad.clone()] + [p.grad.clone() for p in tm.parameters()]
    x.grad = e.grad = None
    for p in tm.parameters():
        p.grad = None
    import math
    lnx = x + math.log(2.0) * e                    # reference formula, modules.py:230-247
    ref = torch.tanh(torch.cat([tm.layers_r(lnx[:, 0:1]), tm.layers_g(lnx[:, 1:2]), tm.layers_b(lnx[:, 2:3])], -1))
    assert torch.allclose(y, ref, atol=2e-6)
    (ref * w).sum().backward()
    want = [x.grad, e.grad] + [p.grad for p in tm.parameters()]
    for a, b in zip(got, want):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * b.abs().max().item() + 1e-7)
