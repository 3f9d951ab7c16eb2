"""CPU checks of the drop-in nn.Module surface: constructor signatures, state_dict keys and shapes,
initialisation ranges (against the reference's own modules when /root/reference is present), flat
parameter re-homing, and the loud failure of the CUDA-only paths."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import harness, modules, _lib  # noqa: E402
from neucam_b200.step import FlatParameters, MODEL_ORDER  # noqa: E402
from oracle import neucam_oracle as orc, ref_shim  # noqa: E402


def test_state_dict_keys_match_reference_layout():
    m = harness.build_static_models((2, 30, 45), [-1.0, 1.0])
    assert list(m["atlas_b"].state_dict().keys()) == [f"net.net.{i}.0.{n}" for i in range(5) for n in ("weight", "bias")]
    assert m["atlas_b"].state_dict()["net.net.0.0.weight"].shape == (512, 2)
    assert m["atlas_b"].state_dict()["net.net.4.0.weight"].shape == (3, 512)
    assert list(m["blur"].state_dict().keys()) == [f"layers.{i}.{n}" for i in range(4) for n in ("weight", "bias")]
    assert m["blur"].state_dict()["layers.3.weight"].shape == (27, 64)
    assert list(m["tm"].state_dict().keys()) == [f"layers_{c}.{i}.{n}" for c in "rgb" for i in (0, 2) for n in ("weight", "bias")]
    assert list(m["exp"].state_dict().keys()) == ["exps"] and list(m["focal"].state_dict().keys()) == ["focus"]
    assert sum(p.numel() for p in m["atlas_b"].parameters()) == 791043
    assert sum(p.numel() for p in m["blur"].parameters()) == 10331
    assert sum(p.numel() for p in m["tm"].parameters()) == 1155


def test_siren_init_ranges():
    torch.manual_seed(0)
    net = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=512, num_hidden_layers=3, in_features=2,
                               out_features=3, sidelength=(30, 45), num_frequencies=7)
    sd = net.state_dict()
    assert sd["net.net.0.0.weight"].abs().max() <= 0.5 and sd["net.net.0.0.weight"].abs().max() > 0.45
    bound = math.sqrt(6 / 512) / 30
    for i in (1, 2, 3, 4):
        w = sd[f"net.net.{i}.0.weight"]
        assert w.abs().max() <= bound and w.abs().max() > 0.95 * bound
        assert sd[f"net.net.{i}.0.bias"].abs().max() <= 1 / math.sqrt(512)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_same_seed_same_parameters_as_reference():
    """Identical constructor order + init calls => identical parameters under the same torch seed."""
    ref_modules = ref_shim.load("modules")
    torch.manual_seed(123)
    ours = harness.build_static_models((3, 20, 30), [-1.0, 0.0, 1.0], seed=123)
    torch.manual_seed(123)
    ref = {
        "atlas_b": ref_modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=512, num_hidden_layers=3,
                                            in_features=2, out_features=3, sidelength=(20, 30), num_frequencies=7),
        "exp": ref_modules.LearnExps(3),
        "focal": ref_modules.LearnFocal([-1.0, 0.0, 1.0], True),
        "blur": ref_modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                      num_frequencies=7, skip_layer=[4], nonlinear="both"),
        "tm": ref_modules.TonemapNet(),
    }
    for slot, rm in ref.items():
        a, b = ours[slot].state_dict(), rm.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert torch.equal(a[k], b[k]), (slot, k)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_small_modules_match_reference_forward():
    ref_modules = ref_shim.load("modules")
    torch.manual_seed(1)
    ours = modules.SimpleMLP(256, 6, 3, 2, True, 3, [4], "tanh")
    ref = ref_modules.SimpleMLP(256, 6, 3, 2, True, 3, [4], "tanh")
    ref.load_state_dict(ours.state_dict())
    x = torch.rand(1, 50, 3) * 2 - 1
    assert torch.allclose(ours(x), ref(x), atol=1e-6)
    pe_o, pe_r = modules.PosEncodingNeRF(3, num_frequencies=3), ref_modules.PosEncodingNeRF(3, num_frequencies=3)
    assert torch.allclose(pe_o(x), pe_r(x), atol=1e-6)
    tm_o, tm_r = modules.TonemapNet(), ref_modules.TonemapNet()
    tm_r.load_state_dict(tm_o.state_dict())
    r, e = torch.rand(40, 3) * 3, torch.rand(40, 1) * 6 - 3
    assert torch.allclose(tm_o(r, e), tm_r(r, e), atol=1e-6)
    # oracle's functional restatements agree too
    assert torch.allclose(orc.tonemap(tm_o.state_dict(), r, e)[0], tm_r(r, e), atol=1e-6)
    assert torch.allclose(orc.simple_mlp(ours.state_dict(), x, "tanh", 3, (4,), 3), ref(x), atol=1e-6)


def test_flat_parameters_alias_and_order():
    m = harness.build_static_models((2, 30, 45), [-1.0, 1.0])
    before = {s: {k: v.clone() for k, v in mod.state_dict().items()} for s, mod in m.items() if mod is not None}
    fp = FlatParameters(m, "cpu")
    assert list(fp.slot_ranges.keys()) == [s for s in MODEL_ORDER if m.get(s) is not None]
    for s, mod in m.items():
        if mod is None:
            continue
        for k, v in mod.state_dict().items():
            assert torch.equal(v, before[s][k])
    # parameters alias the flat buffer; gradients alias the flat gradient
    p = m["tm"].layers_g[0].weight
    p.data.add_(1.0)
    o = [e for e in fp.entries if e[0] == "tm" and e[1] == "layers_g.0.weight"][0][3]
    assert fp.flat[o].item() == pytest.approx(p.view(-1)[0].item())
    fp.grad.fill_(2.0)
    assert p.grad.view(-1)[0].item() == 2.0
    assert fp.numel >= 802533 and all(e[3] % 4 == 0 for e in fp.entries)


def test_cuda_only_paths_fail_loudly_on_cpu():
    net = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=512, num_hidden_layers=3, in_features=2,
                               out_features=3, sidelength=(30, 45))
    with pytest.raises(_lib.NeucamNativeError):
        net({"coords": torch.zeros(1, 4, 2)})
    with pytest.raises(NotImplementedError):
        modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=256, num_hidden_layers=3, in_features=2,
                             out_features=3)({"coords": torch.zeros(1, 4, 2)})


def test_mgrid_matches_reference_formula():
    g = harness.mgrid((2, 5, 7))
    assert g.shape == (70, 3)
    assert g[0].tolist() == [-1.0, -1.0, -1.0] and g[-1].tolist() == [1.0, 1.0, 1.0]
    g1 = harness.mgrid((1, 5, 7))
    assert (g1[:, 0] == -1.0).all()   # single frame: t = 0/max(N-1,1) -> -1
    if ref_shim.available():
        dataio = ref_shim.load("dataio")
        assert torch.equal(g, dataio.get_mgrid((2, 5, 7), dim=3).view(-1, 3))


def test_forward_with_activations_matches_reference():
    """FCBlock / SingleBVPNet.forward_with_activations (modules.py:97-117, 477-481): same entries, same values as the
    reference's (its keys carry the class path, which differs by package name only)."""
    torch.manual_seed(4)
    net = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=32, num_hidden_layers=2, in_features=2,
                               out_features=3, sidelength=(30, 45))
    x = torch.rand(1, 17, 2) * 2 - 1
    out = net.forward_with_activations({"coords": x})
    acts = out["activations"]
    assert list(acts)[0] == "input" and len(acts) == 1 + 2 * 3 + 1 - 1     # input + (linear, sine) x 3; last popped
    key, y = out["model_out"]
    assert "BatchLinear" in key and key.endswith("_3") and y.shape == (1, 17, 3)
    # values: the explicit chain of modules.py:17-35
    h = x
    sd = net.state_dict()
    for i in range(3):
        z = h @ sd[f"net.net.{i}.0.weight"].t() + sd[f"net.net.{i}.0.bias"]
        h = torch.sin(30 * z)
        assert torch.allclose(acts[[k for k in acts if k.endswith(f"Sine'>_{i}")][0]], h, atol=1e-6)
    assert torch.allclose(y, h @ sd["net.net.3.0.weight"].t() + sd["net.net.3.0.bias"], atol=1e-6)
    # retain_grad keeps gradients of the intermediates
    a = net.net.forward_with_activations(x, retain_grad=True)
    list(a.values())[-1].sum().backward()
    assert all(v.grad is not None for k, v in a.items() if k != "input")
    if ref_shim.available():
        ref_modules = ref_shim.load("modules")
        torch.manual_seed(4)
        ref = ref_modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=32, num_hidden_layers=2, in_features=2,
                                       out_features=3, sidelength=(30, 45))
        ro = ref.forward_with_activations({"coords": x})
        assert len(ro["activations"]) == len(acts)
        for (ka, va), (kb, vb) in zip(acts.items(), ro["activations"].items()):
            assert ka.split(".")[-1] == kb.split(".")[-1] and torch.allclose(va, vb, atol=1e-6), (ka, kb)
        assert torch.allclose(ro["model_out"][1], y, atol=1e-6)


def test_from_opt_refuses_options_outside_the_path():
    import argparse
    from neucam_b200.step import StepOptions
    base = dict(patch_size=9, offset_size=5, lr=1e-4)
    StepOptions.from_opt(argparse.Namespace(**base))
    for flag in ("flow_loss", "mask_loss", "use_depth", "patch_sample", "denoise"):
        with pytest.raises(NotImplementedError):
            StepOptions.from_opt(argparse.Namespace(**base, **{flag: True}))
