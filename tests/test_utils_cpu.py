"""CPU checks of checkpoint compatibility (utils.save_model keys, utils.py:645-673) and of the freeze-schedule
learning-rate groups (training.py:300-313)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neucam_b200 import harness, utils as nutils  # noqa: E402
from neucam_b200.training import lr_groups_after_freeze  # noqa: E402
from oracle import ref_shim  # noqa: E402


def test_checkpoint_roundtrip_and_keys(tmp_path):
    m = harness.build_static_models((2, 10, 14), [-1.0, 1.0], hidden=64, seed=0)
    path = os.path.join(str(tmp_path), "model_current.pth")
    nutils.save_model(m, path)
    ck = torch.load(path)
    assert set(ck.keys()) == {"model_state_dict", "model_tm_state_dict", "model_blur_state_dict",
                              "model_exp_state_dict", "model_focal_state_dict"}
    m2 = harness.build_static_models((2, 10, 14), [-1.0, 1.0], hidden=64, seed=7)
    nutils.load_model(m2, path)
    for slot in ("atlas_b", "tm", "blur", "exp", "focal"):
        a, b = m[slot].state_dict(), m2[slot].state_dict()
        assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_checkpoint_is_what_the_reference_writes(tmp_path):
    """Same keys and shapes as the reference's own utils.save_model applied to the reference's modules."""
    ref_modules = ref_shim.load("modules")
    ref_utils = ref_shim.load("utils")
    torch.manual_seed(0)
    ref = {k: None for k in harness.SLOTS}
    ref["atlas_b"] = ref_modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=64, num_hidden_layers=3,
                                              in_features=2, out_features=3, sidelength=(10, 14), num_frequencies=7)
    ref["exp"] = ref_modules.LearnExps(2)
    ref["focal"] = ref_modules.LearnFocal([-1.0, 1.0], True)
    ref["blur"] = ref_modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                        num_frequencies=7, skip_layer=[4], nonlinear="both")
    ref["tm"] = ref_modules.TonemapNet()
    p_ref = os.path.join(str(tmp_path), "ref.pth")
    ref_utils.save_model(ref, p_ref)
    ours = harness.build_static_models((2, 10, 14), [-1.0, 1.0], hidden=64, seed=3)
    p_ours = os.path.join(str(tmp_path), "ours.pth")
    nutils.save_model(ours, p_ours)
    a, b = torch.load(p_ref), torch.load(p_ours)
    assert set(a.keys()) == set(b.keys())
    for key in a:
        assert list(a[key].keys()) == list(b[key].keys())
        assert all(a[key][k].shape == b[key][k].shape for k in a[key])
    # and the reference's checkpoint loads into our modules
    nutils.load_model(ours, p_ref)
    assert torch.equal(ours["tm"].state_dict()["layers_g.2.weight"], ref["tm"].state_dict()["layers_g.2.weight"])


def test_freeze_schedule_groups():
    g = lr_groups_after_freeze()
    assert g["atlas_b"] == 1e-5 and g["blur"] == 1e-6
    assert g["tm"] is None and g["exp"] is None and g["focal"] is None
