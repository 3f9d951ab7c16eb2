"""GPU tests of the SURVEY section 8f rows built so far: on-device sampler / gather (rank 3), checkpoint + full-frame
render (rank 4), the train() loop with the freeze schedule (caller row)."""
import argparse
import os

import pytest
import torch

from oracle_replay import orc, rel

from neucam_b200 import harness, utils as nutils
from neucam_b200.step import FusedTrainStep, StepOptions
from neucam_b200.training import train

pytestmark = pytest.mark.gpu


def _step(shape, focal, B, gamma=False, seed=3):
    models = harness.build_static_models(shape, focal, hidden=512, seed=seed)
    return FusedTrainStep(models, shape, StepOptions(use_random_gamma=gamma), B, device="cuda"), models


def test_device_gather_matches_host_sampler_contract():
    """Injected indices: coords/img formed on the device equal the reference-layout host batch bit for bit."""
    shape, focal, B = (3, 17, 23), [-1.0, 0.0, 1.0], 999
    scene = harness.SyntheticScene(shape, focal, seed=1)
    step, _ = _step(shape, focal, B, gamma=True)
    step.attach_dataset(scene.data, gamma_weights=scene.gamma_weights)
    inp, gt = scene.sample(B, torch.Generator().manual_seed(5))
    step.sample_resident(indices=inp["coord_idx"][0], gamma=torch.tensor([12.5]))
    torch.cuda.synchronize()
    assert torch.equal(step.coord_idx.cpu(), inp["coord_idx"][0])
    assert torch.equal(step.coords.cpu(), inp["coords"][0])            # get_mgrid formula, float64 -> float32
    assert torch.equal(step.img.cpu(), gt["img"][0])
    assert torch.equal(step.gamma_w.cpu(), inp["gamma_weights"][0, :, 0])
    assert step.gamma.item() == 12.5


def test_device_sampler_is_uniform_with_replacement_and_reproducible():
    shape, focal, B = (2, 40, 50), [-1.0, 1.0], 20000
    scene = harness.SyntheticScene(shape, focal, seed=1)
    step, _ = _step(shape, focal, B)
    step.attach_dataset(scene.data)
    step.sample_resident(seed=7, step=3)
    a = step.coord_idx.clone()
    step.sample_resident(seed=7, step=4)
    b = step.coord_idx.clone()
    step.sample_resident(seed=7, step=3)
    torch.cuda.synchronize()
    assert torch.equal(step.coord_idx, a) and not torch.equal(a, b)
    total = 2 * 40 * 50
    assert a.min() >= 0 and a.max() < total
    counts = torch.bincount(a.cpu(), minlength=total).float()
    assert counts.max() > 1                                            # with replacement
    assert abs(counts.mean().item() - B / total) < 1e-6 and counts.std().item() < 2.0 * (B / total) ** 0.5
    assert abs(a.float().mean().item() / total - 0.5) < 0.02
    # gathered data corresponds to the drawn indices
    assert torch.equal(step.img.cpu(), scene.data[a.cpu()])


def test_render_and_checkpoint_roundtrip(tmp_path):
    shape, focal = (2, 16, 20), [-1.0, 1.0]
    models = harness.build_static_models(shape, focal, hidden=512, seed=2)
    for m in models.values():
        if m is not None:
            m.cuda()
    ldr, hdr = nutils.render_frames(models, shape, chunk=300)
    state = {s: {k: v.detach().cpu() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    grid = harness.mgrid(shape)
    with torch.no_grad():
        rgb = orc.sine_mlp(state["atlas_b"], grid[:, 1:])
        frame = torch.arange(grid.shape[0]) // (16 * 20)
        ref_ldr, _ = orc.tonemap(state["tm"], rgb, (3 * torch.tanh(state["exp"]["exps"]))[frame].view(-1, 1))
    assert rel(ldr.cpu().reshape(-1, 3) * 2 - 1, ref_ldr) <= 2e-3
    assert rel(hdr.cpu().reshape(-1, 3), torch.exp(rgb[: 16 * 20])) <= 2e-3
    path = os.path.join(str(tmp_path), "m.pth")
    nutils.save_model(models, path)
    other = harness.build_static_models(shape, focal, hidden=512, seed=9)
    for m in other.values():
        if m is not None:
            m.cuda()
    nutils.load_model(other, path)
    ldr2, _ = nutils.render_frames(other, shape, chunk=1000)
    assert torch.equal(ldr2, ldr)


def test_train_loop_with_freeze_schedule(tmp_path):
    """train() (reference signature): loss goes down; at start_freeze_tm the CRF / exposure / focus stop moving,
    Adam restarts and mask weights are applied (training.py:300-314)."""
    shape, focal, B = (3, 16, 20), [-1.0, 0.0, 1.0], 300
    scene = harness.SyntheticScene(shape, focal, seed=4, max_radius=2)
    gen = torch.Generator().manual_seed(8)

    class Loader:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for _ in range(self.n):
                inp, gt = scene.sample(B, gen)
                inp["weights"] = torch.ones(1, B, 1)
                yield inp, gt

    opt = argparse.Namespace(num_epochs=4, lr=1e-4, steps_til_summary=5, epochs_til_ckpt=2, start_freeze_tm=11,
                             patch_size=9, offset_size=5, use_random_gamma=False, fixed_value=0.0)
    models = harness.build_static_models(shape, focal, hidden=512, seed=6)
    hist = train(models, Loader(5), str(tmp_path), video_shape=shape, opt=opt, log_every=1)
    assert len(hist) == 20
    assert hist[-1][1]["img_loss"] < hist[0][1]["img_loss"]
    assert os.path.exists(os.path.join(str(tmp_path), "checkpoints", "model_final.pth"))
    assert os.path.exists(os.path.join(str(tmp_path), "checkpoints", "model_epoch_0002.pth"))
    # frozen groups: snapshot after the freeze step equals the final value
    ck = torch.load(os.path.join(str(tmp_path), "checkpoints", "model_final.pth"))
    assert "model_tm_state_dict" in ck and "model_exp_state_dict" in ck


def test_frozen_groups_do_not_move():
    shape, focal, B = (2, 16, 20), [-1.0, 1.0], 256
    scene = harness.SyntheticScene(shape, focal, seed=4)
    step, models = _step(shape, focal, B)
    inp, gt = scene.sample(B, torch.Generator().manual_seed(1))
    from neucam_b200.training import lr_groups_after_freeze
    before = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    step.load_batch(inp, gt)
    step.forward_backward()
    step.optimizer_step(lr_groups=lr_groups_after_freeze())
    torch.cuda.synchronize()
    for slot in ("tm", "exp", "focal"):
        for k, v in models[slot].state_dict().items():
            assert torch.equal(v, before[slot][k]), (slot, k)
    moved_atlas = (models["atlas_b"].state_dict()["net.net.2.0.weight"] - before["atlas_b"]["net.net.2.0.weight"]).abs().max()
    moved_blur = (models["blur"].state_dict()["layers.1.weight"] - before["blur"]["layers.1.weight"]).abs().max()
    assert 0.5e-5 < moved_atlas.item() <= 1.01e-5 and 0.5e-6 < moved_blur.item() <= 1.01e-6


def test_parameter_writes_after_the_step_exists_are_picked_up(tmp_path):
    """FusedTrainStep keeps fp16 tensor-core operand copies of the atlas weights.  A checkpoint loaded (utils.load_model)
    or any torch write to the parameters AFTER the step was built must reach those copies before the next launch: the
    step watches the flat buffer's version counter and re-packs (load_model left them stale)."""
    from neucam_b200 import utils as nutils
    shape, focal, B = (2, 24, 36), [-1.0, 1.0], 300
    scene = harness.SyntheticScene(shape, focal, seed=1)
    inp, gt = scene.sample(B, torch.Generator().manual_seed(2))
    donor = harness.build_static_models(shape, focal, hidden=512, seed=11)
    path = str(tmp_path / "ckpt.pth")
    nutils.save_model(donor, path)

    models = harness.build_static_models(shape, focal, hidden=512, seed=3)
    step = FusedTrainStep(models, shape, StepOptions(), B, device="cuda")
    step.load_batch(inp, gt)
    step.forward_backward()
    loss_before = step.losses()["total"].item()
    nutils.load_model(models, path)                    # writes into the flat buffer's views, after the step exists
    step.forward_backward()
    torch.cuda.synchronize()
    got = (step.losses()["total"].item(), step.params.grad.clone())

    fresh_models = harness.build_static_models(shape, focal, hidden=512, seed=11)
    fresh = FusedTrainStep(fresh_models, shape, StepOptions(), B, device="cuda")
    fresh.load_batch(inp, gt)
    fresh.forward_backward()
    torch.cuda.synchronize()
    assert got[0] != loss_before
    assert got[0] == fresh.losses()["total"].item()
    assert torch.equal(got[1], fresh.params.grad)
