"""GPU parity of the WHOLE fused step (neucam_b200.step.FusedTrainStep, every kernel through the C ABI)
against (a) the oracle on the same inputs and (b) golden vectors recorded from the unmodified reference's
training.train loop (tests/golden/mf_static_h512.pt).  Tolerances (BASELINE.json north_star): forward
quantities <= 1e-3 relative, gradients <= 1e-2 relative."""
import math

import pytest
import torch

from oracle_replay import batch_of, cfg_of, load_fixture, orc, rel, replay_rng

from neucam_b200 import harness
from neucam_b200.step import FusedTrainStep, StepOptions

pytestmark = pytest.mark.gpu


def make_step(fx_or_cfg, state, B, gamma):
    shape = fx_or_cfg["shape"]
    models = harness.build_static_models(shape, fx_or_cfg["focal_list"], hidden=512, seed=0)
    harness.load_state(models, state)
    opts = StepOptions(patch_size=9, offset_size=5, use_random_gamma=gamma, fixed_value=0.0, lr=1e-4)
    return FusedTrainStep(models, shape, opts, B, device="cuda"), models


def grads_of(models):
    return {s: {k: p.grad.detach().float().cpu().clone() for k, p in m.named_parameters()}
            for s, m in models.items() if m is not None}


def test_step_matches_golden_reference_run():
    fx = load_fixture("mf_static_h512")
    step, models = make_step(fx, fx["init_state"], fx["B"], False)
    cfg = cfg_of(fx)
    state = {s: {k: v.clone() for k, v in d.items()} for s, d in fx["init_state"].items()}
    for i in range(fx["n_steps"]):
        inp, gt = fx["batches"][i]
        step.load_batch(inp, gt)
        step.forward_backward()
        torch.cuda.synchronize()
        losses = {k: v.item() for k, v in step.losses().items()}
        ref = fx["losses"][i]
        assert losses["img_loss"] == pytest.approx(ref["loss/img_loss"], rel=1e-3)
        assert losses["ue_loss"] == pytest.approx(ref["loss/ue_loss"], rel=1e-3)
        g = grads_of(models)
        worst = 0.0
        for slot, d in fx["grads_digest"][i].items():
            for k, dg in d.items():
                mine = g[slot][k].reshape(-1)
                if dg["norm"] == 0:
                    assert mine.abs().max() == 0, (slot, k)
                    continue
                e = abs(mine.double().norm().item() - dg["norm"]) / dg["norm"]
                e2 = rel(mine[dg["idx"]], dg["val"])
                worst = max(worst, e, e2)
                # step 0 starts from identical parameters: BASELINE.json's 1e-2 gradient bound applies as is.  Later
                # steps are free-running (each side applied ITS OWN Adam update, parameters differ by up to 2.5e-5
                # per weight), so the gradient difference also contains that divergence: 2e-2 there.
                tol = 1e-2 if i == 0 else 2e-2
                assert e <= tol and e2 <= tol, (i, slot, k, e, e2)
        print(f"step {i}: img_loss {losses['img_loss']:.6f} (ref {ref['loss/img_loss']:.6f}); worst grad err {worst:.2e}")
        before = {(slot, k): dict(models[slot].named_parameters())[k].detach().float().cpu().reshape(-1).clone()
                  for slot, d in fx["params_digest"][i].items() for k in d}
        step.optimizer_step()
        torch.cuda.synchronize()
        for slot, d in fx["params_digest"][i].items():
            for k, dg in d.items():
                mine = dict(models[slot].named_parameters())[k].detach().float().cpu().reshape(-1)
                idx = dg["idx"]
                # (a) the UPDATE Adam applied (~lr = 1e-4 per weight in the first steps) against the reference's own
                # update of the same entries: relative, so an error of the size of the step cannot hide.  Step 0
                # starts from identical parameters and moments; later steps are free-running on both sides.
                ref_prev = (fx["init_state"][slot][k].reshape(-1)[idx] if i == 0
                            else fx["params_digest"][i - 1][slot][k]["val"])
                d_ref = (dg["val"] - ref_prev).double()
                d_mine = (mine[idx] - before[(slot, k)][idx]).double()
                if d_ref.norm() > 0:
                    e_upd = ((d_mine - d_ref).norm() / d_ref.norm()).item()
                    assert e_upd <= (2e-2 if i == 0 else 1e-1), (i, slot, k, e_upd)
                # (b) the parameters themselves: a fraction of one Adam step
                assert (mine[idx] - dg["val"]).abs().max() <= (5e-6 if i == 0 else 2.5e-5), (i, slot, k)


@pytest.mark.parametrize("name,gamma", [("mf", False), ("mfme_gamma", True)])
def test_step_matches_oracle_full_gradients(name, gamma):
    shape = (3, 40, 60) if gamma else (2, 36, 50)
    focal = [-1.0, 0.0, 1.0] if gamma else [-1.0, 1.0]
    B = 1111
    models0 = harness.build_static_models(shape, focal, hidden=512, seed=3)
    with torch.no_grad():
        models0["exp"].exps.copy_(torch.linspace(-0.3, 0.4, shape[0]))
        for p in models0["tm"].parameters():
            p.mul_(1.3)
    state = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models0.items() if m is not None}
    scene = harness.SyntheticScene(shape, focal, seed=5)
    gen = torch.Generator().manual_seed(9)
    inp, gt = scene.sample(B, gen)
    inp["gamma_weights"] = torch.rand(1, B, 1, generator=gen)
    gval = torch.tensor([37.5]) if gamma else None
    step, models = make_step(dict(shape=shape, focal_list=focal), state, B, gamma)
    step.load_batch(inp, gt, gval)
    step.forward_backward()
    torch.cuda.synchronize()

    cfg = orc.StepConfig(shape, use_random_gamma=gamma, with_grad_loss=False)
    batch = dict(inp)
    batch["img"] = gt["img"]
    losses, grads = orc.step_grads(state, batch, cfg, gval)
    mine = {k: v.item() for k, v in step.losses().items()}
    assert mine["img_loss"] == pytest.approx(losses["img_loss"].item(), rel=1e-3)
    assert mine["ue_loss"] == pytest.approx(losses["ue_loss"].item(), rel=1e-3)
    g = grads_of(models)
    report = []
    for slot, d in grads.items():
        for k, gr in d.items():
            if gr.abs().max() == 0:
                assert g[slot][k].abs().max() == 0, (slot, k)
                continue
            e = rel(g[slot][k], gr)
            report.append((e, slot, k))
            assert e <= 1e-2, (slot, k, e)
    report.sort(reverse=True)
    print(f"{name}: img_loss {mine['img_loss']:.6f}; worst grads: " + ", ".join(f"{s}.{k}={e:.1e}" for e, s, k in report[:4]))

    # forward intermediates: tap coordinates and atlas output
    _, inter = orc.step_losses(state, batch, cfg, gval, return_intermediates=True)
    assert rel(step.uv.cpu(), inter["uv"]) <= 1e-5
    assert rel(step.y.cpu(), inter["atlas_out"]) <= 2e-3


def test_cuda_graph_replay_equals_eager_steps():
    """FusedTrainStep.capture(): graph replays take exactly the same steps as eager launches."""
    shape, focal, B = (2, 36, 50), [-1.0, 1.0], 900
    scene = harness.SyntheticScene(shape, focal, seed=5)
    gen = torch.Generator().manual_seed(11)
    batches = [scene.sample(B, gen) for _ in range(3)]
    finals = []
    for use_graph in (False, True):
        models = harness.build_static_models(shape, focal, hidden=512, seed=3)
        step = FusedTrainStep(models, shape, StepOptions(), B, device="cuda")
        if use_graph:
            step.load_batch(*batches[0])
            step.capture()
        losses = []
        for inp, gt in batches:
            losses.append(step.step(inp, gt)["total"].item())
        torch.cuda.synchronize()
        finals.append((losses, step.params.flat.clone(), step.params.adam_state.clone()))
    (l0, p0, s0), (l1, p1, s1) = finals
    assert s0[0].item() == 3.0 and s1[0].item() == 3.0
    # no float atomics anywhere (fixed-order reductions): a graph replay launches the same kernels on the same
    # data as the eager path, so losses and parameters are BIT-identical
    assert l0 == l1
    assert torch.equal(p0, p1)


def test_mask_weights_and_value_only_grad_loss():
    """After start_freeze_tm the reference weights the squared error with per-pixel mask weights
    (loss_functions.py:10); and it logs a value-only grad_loss (training.py:248-252)."""
    shape, focal, B = (3, 30, 40), [-1.0, 0.0, 1.0], 800
    models0 = harness.build_static_models(shape, focal, hidden=512, seed=4)
    with torch.no_grad():
        models0["tm"].layers_g[2].weight.mul_(-1.0)      # make some CRF slopes negative so grad_loss is non-zero
    state = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models0.items() if m is not None}
    scene = harness.SyntheticScene(shape, focal, seed=5)
    gen = torch.Generator().manual_seed(9)
    inp, gt = scene.sample(B, gen)
    inp["weights"] = torch.rand(1, B, 1, generator=gen) + 0.2
    step, models = make_step(dict(shape=shape, focal_list=focal), state, B, False)
    step.load_batch(inp, gt)
    step.forward_backward(use_mask=True)
    torch.cuda.synchronize()
    batch = dict(inp)
    batch["img"] = gt["img"]
    batch["use_mask"] = True
    rad, expo = torch.rand(10000, 3, generator=gen) * 5, torch.rand(10000, 1, generator=gen) * 6 - 3.0
    cfg = orc.StepConfig(shape, with_grad_loss=True)
    losses, grads = orc.step_grads(state, batch, cfg, None, (rad, expo))
    assert step.losses()["img_loss"].item() == pytest.approx(losses["img_loss"].item(), rel=1e-3)
    g = grads_of(models)
    for slot in ("atlas_b", "tm", "blur", "exp"):
        for k, gr in grads[slot].items():
            if gr.abs().max() > 0:
                assert rel(g[slot][k], gr) <= 1e-2, (slot, k)
    gl = step.grad_loss(rad, expo).item()
    assert losses["grad_loss"].item() > 0
    assert gl == pytest.approx(losses["grad_loss"].item(), rel=2e-3)
