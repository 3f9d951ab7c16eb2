"""Host logic of neucam_b200.layered.LayeredTrainStep (everything above the CUDA kernels: taps, uv warps, layer
blend, PSF sum, CRF call, image / rigidity / sparsity / alpha / ue losses) against golden runs of the UNMODIFIED
reference (tests/golden/*.pt, oracle/make_golden.py).  The scene networks have no CPU path, so this test swaps
FCBlock.forward for a plain fp32 torch stand-in -- test scaffolding, defined here, never shipped."""
import pytest
import torch

from neucam_b200 import harness, modules
from neucam_b200.layered import LayeredTrainStep, needs_layered_step
from neucam_b200.step import StepOptions
from oracle_replay import batch_of, load_fixture, rel, replay_rng


def _torch_sine_block(self, coords, params=None, **kw):
    sd = dict(self.named_parameters())
    x = coords
    for i in range(5):
        x = x @ sd[f"net.{i}.0.weight"].t() + sd[f"net.{i}.0.bias"]
        if i < 4:
            x = torch.sin(30.0 * x)
    return x


@pytest.fixture
def cpu_scene_networks(monkeypatch):
    monkeypatch.setattr(modules.FCBlock, "forward", _torch_sine_block)
    monkeypatch.setattr(modules.FCBlock, "fused_eligible", lambda self: True)


def _models_for(fx):
    if "kind" in fx:
        return harness.build_layered_models(fx["kind"], fx["shape"], fx["hidden"], fx["mlp_hidden"])
    return harness.build_static_models(fx["shape"], fx["focal_list"], fx["hidden"])


def _options(fx):
    o = fx["opt"]
    return StepOptions(patch_size=o["patch_size"], offset_size=o["offset_size"], use_random_gamma=fx["use_random_gamma"],
                       fixed_value=o["fixed_value"], lr=o["lr"], uv_mapping_scale=o.get("uv_mapping_scale", 1.0),
                       rigidity_loss_w=o.get("rigidity_loss_w"), rigidity_loss_steps=10 ** 9,
                       sparsity_loss_w=o.get("sparsity_loss_w"), alpha_loss_w=o.get("alpha_loss_w"))


@pytest.mark.parametrize("name", ["mf_static_small", "mfme_gamma_small", "me_dynamic_small", "video_deblur_small",
                                  "video_hdr_small"])
def test_losses_and_gradients_match_reference_run(name, cpu_scene_networks):
    fx = load_fixture(name)
    models = _models_for(fx)
    harness.load_state(models, fx["init_state"])
    step = LayeredTrainStep(models, fx["shape"], _options(fx), _host_logic_only=True)
    b = batch_of(fx, 0)
    gamma, _ = replay_rng(fx, 0)
    losses = step.forward(b, {"img": b["img"]}, gamma)
    ref = fx["losses"][0]
    for k, v in losses.items():
        if k != "total":
            assert v.item() == pytest.approx(ref["loss/" + k], rel=5e-5), k
    # the reference's total also carries the value-only grad_loss
    assert losses["total"].item() + ref["loss/grad_loss"] == pytest.approx(ref["loss/total_train_loss"], rel=1e-4)
    assert set(losses) - {"total"} == {k[5:] for k in ref} - {"grad_loss", "total_train_loss"}
    losses["total"].backward()
    for slot, d in fx["grads"][0].items():
        for k, g_ref in d.items():
            p = dict(models[slot].named_parameters())[k]
            g = torch.zeros_like(p) if p.grad is None else p.grad
            if g_ref.abs().max() == 0:
                assert g.abs().max() == 0, (slot, k)
            else:
                assert rel(g, g_ref) < 3e-4, (slot, k, rel(g, g_ref))


def test_rigidity_mask_follows_reference_broadcast(cpu_scene_networks):
    """After start_freeze_tm the reference multiplies the [B] rigidity loss by the [1,B,1] refine weights, i.e.
    takes mean(loss) * mean(weights) (utils.py:452-453)."""
    fx = load_fixture("me_dynamic_small")
    models = _models_for(fx)
    harness.load_state(models, fx["init_state"])
    step = LayeredTrainStep(models, fx["shape"], _options(fx), _host_logic_only=True)
    b = batch_of(fx, 0)
    gamma, _ = replay_rng(fx, 0)
    plain = step.forward(b, {"img": b["img"]}, gamma)["rigidity_loss"]
    b["weights"] = torch.full((1, fx["B"], 1), 0.2)
    b["weights"][0, ::2] = 1.0
    masked = step.forward(b, {"img": b["img"]}, gamma, use_mask=True)["rigidity_loss"]
    assert masked.item() == pytest.approx(plain.item() * b["weights"].mean().item(), rel=1e-5)


def test_dispatch_rule():
    assert not needs_layered_step(harness.build_static_models((2, 8, 8), [-1.0, 1.0], 64))
    assert needs_layered_step(harness.build_layered_models("me_dynamic", (3, 8, 8), 64, 32))


def test_host_logic_instance_cannot_train(cpu_scene_networks):
    fx = load_fixture("me_dynamic_small")
    step = LayeredTrainStep(_models_for(fx), fx["shape"], _options(fx), _host_logic_only=True)
    b = batch_of(fx, 0)
    with pytest.raises(Exception):
        step.forward_backward(b, {"img": b["img"]}, 2.0)


@pytest.mark.parametrize("name", ["video_deblur_small", "me_dynamic_small"])
def test_pixel_shards_sum_to_the_full_batch(name, cpu_scene_networks):
    """Data parallel contract: two ranks with a (ragged) half of the pixels each, losses
    weighted by B_local / B_global and ue_loss by 1 / world, SUM of the gradients == one rank with all pixels."""
    fx = load_fixture(name)
    b = batch_of(fx, 0)
    gamma, _ = replay_rng(fx, 0)
    B = fx["B"]
    cut = B // 2 - 3

    def grads_of(world, lo, hi):
        models = _models_for(fx)
        harness.load_state(models, fx["init_state"])
        step = LayeredTrainStep(models, fx["shape"], _options(fx), global_pixels=B, _host_logic_only=True)
        step.world = world
        part = {k: (v[:, lo:hi] if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == B else v) for k, v in b.items()}
        losses = step.forward(part, {"img": part["img"]}, gamma)
        losses["total"].backward()
        return ({k: v.item() for k, v in losses.items()},
                {(s, k): (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
                 for s, m in models.items() if m is not None for k, p in m.named_parameters()})

    l_full, g_full = grads_of(1, 0, B)
    l0, g0 = grads_of(2, 0, cut)
    l1, g1 = grads_of(2, cut, B)
    for k in l_full:
        assert l0[k] + l1[k] == pytest.approx(l_full[k], rel=2e-5), k
    for key, g in g_full.items():
        if g.abs().max() > 0:
            assert rel(g0[key] + g1[key], g) < 2e-4, key
