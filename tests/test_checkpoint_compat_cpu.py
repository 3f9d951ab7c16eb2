"""Checkpoint compatibility with the reference (utils.save_model, utils.py:645-673), both directions:
  (a) tests/golden/ref_checkpoint_small.pth was written by the reference's OWN save_model (oracle/make_golden_aux.py);
      neucam_b200.utils.load_model must load it into this package's modules, every tensor bit-equal;
  (b) a checkpoint written by neucam_b200.utils.save_model must load into the reference's modules with
      load_state_dict(strict=True) (needs /root/reference; its KEY SET is compared with the golden file always)."""
import os

import pytest
import torch

from oracle_replay import GOLDEN_DIR

from neucam_b200 import modules
from neucam_b200 import utils as nutils

H, W, N = 24, 36, 4
CKPT = os.path.join(GOLDEN_DIR, "ref_checkpoint_small.pth")


def build(mod, seed):
    """The narrow two-layer model set of the fixture, from `mod` = this package's modules or the reference's."""
    torch.manual_seed(seed)
    m = {k: None for k in ("atlas_b", "tm", "uv_b", "uv_f", "blur", "atlas_f", "alpha", "exp", "focal", "noise", "depth")}
    m["atlas_b"] = mod.SingleBVPNet(type="sine", mode="mlp", hidden_features=64, num_hidden_layers=3, in_features=2,
                                    out_features=3, sidelength=(H, W), num_frequencies=7)
    m["atlas_f"] = mod.SingleBVPNet(type="sine", mode="mlp", hidden_features=64, in_features=2, out_features=3,
                                    sidelength=(H, W), num_frequencies=7)
    mlp = lambda layers, nf, od, nl: mod.SimpleMLP(hidden_dim=32, num_layer=layers, in_dim=3, out_dim=od,
                                                   use_encoding=True, num_frequencies=nf, skip_layer=[4], nonlinear=nl)
    m["uv_b"], m["uv_f"], m["alpha"] = mlp(4, 3, 2, "tanh"), mlp(6, 3, 2, "tanh"), mlp(4, 4, 1, "sigmoid")
    m["exp"], m["focal"] = mod.LearnExps(N), mod.LearnFocal(N)
    m["blur"] = mod.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False, num_frequencies=7,
                              skip_layer=[4], nonlinear="both")
    m["tm"] = mod.TonemapNet()
    return m


def test_loads_a_checkpoint_written_by_the_reference_save_model():
    ref = torch.load(CKPT, map_location="cpu")
    assert set(ref) == {nutils.CHECKPOINT_KEYS[s] for s in ("atlas_b", "tm", "uv_b", "uv_f", "blur", "atlas_f", "alpha",
                                                            "exp", "focal")}
    models = build(modules, seed=7)          # different seed: everything must come from the file
    nutils.load_model(models, CKPT)
    n = 0
    for slot, key in nutils.CHECKPOINT_KEYS.items():
        if models.get(slot) is None:
            continue
        own = models[slot].state_dict()
        assert set(own) == set(ref[key]), (slot, set(own) ^ set(ref[key]))
        for k, v in ref[key].items():
            assert own[k].shape == v.shape and torch.equal(own[k], v), (slot, k)
            n += 1
    assert n > 40
    # the loaded networks compute what the reference's computed: same seed as the generator -> same init values
    again = build(modules, seed=42)
    assert torch.equal(again["atlas_b"].state_dict()["net.net.1.0.weight"], ref["model_state_dict"]["net.net.1.0.weight"])
    assert torch.equal(models["exp"].exps.detach(), torch.linspace(-0.4, 0.5, N))


def test_own_checkpoint_has_the_reference_layout(tmp_path):
    models = build(modules, seed=3)
    path = str(tmp_path / "mine.pth")
    nutils.save_model(models, path)
    mine, ref = torch.load(path, map_location="cpu"), torch.load(CKPT, map_location="cpu")
    assert set(mine) == set(ref)
    for key in ref:
        assert list(mine[key]) == list(ref[key]), key            # same tensor names, same order
        for k in ref[key]:
            assert mine[key][k].shape == ref[key][k].shape and mine[key][k].dtype == ref[key][k].dtype, (key, k)


def test_reference_modules_load_own_checkpoint_strict(tmp_path):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference is not present on this machine (the key-layout test above still ran)")
    ref_modules = ref_shim.load("modules")
    models = build(modules, seed=5)
    path = str(tmp_path / "mine.pth")
    nutils.save_model(models, path)
    ckpt = torch.load(path, map_location="cpu")
    theirs = build(ref_modules, seed=9)
    for slot, key in nutils.CHECKPOINT_KEYS.items():          # the load sequence of render_video.py:96-122
        if theirs.get(slot) is not None:
            theirs[slot].load_state_dict(ckpt[key], strict=True)
            for k, v in theirs[slot].state_dict().items():
                assert torch.equal(v, models[slot].state_dict()[k]), (slot, k)
    # and the loaded reference networks evaluate like ours (CPU, fp32)
    x = torch.rand(50, 3) * 2 - 1
    assert torch.allclose(theirs["uv_f"](x), models["uv_f"](x), atol=1e-6)
    assert torch.allclose(theirs["tm"](x, torch.zeros(50, 1)), models["tm"](x, torch.zeros(50, 1)), atol=1e-6)
