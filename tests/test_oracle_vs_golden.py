"""Pins the oracle (oracle/neucam_oracle.py) against golden vectors produced by the UNMODIFIED
reference's training.train loop (oracle/make_golden.py): losses, every parameter gradient, and the
parameters after each Adam update, over several consecutive steps.  CPU only."""
import pytest
import torch

from oracle_replay import batch_of, cfg_of, load_fixture, orc, rel, replay_rng

# the last three take the uv-mapping / foreground-layer / alpha-MLP branches and the rigidity, sparsity and alpha
# losses (configs me_dynamic, video_deblur, video_hdr)
FULL = ["mf_static_small", "mfme_gamma_small", "me_dynamic_small", "video_deblur_small", "video_hdr_small"]


@pytest.mark.parametrize("name", FULL)
def test_losses_grads_and_adam_match_reference(name):
    fx = load_fixture(name)
    cfg = cfg_of(fx)
    state = {s: {k: v.clone() for k, v in d.items()} for s, d in fx["init_state"].items()}
    m = {s: {k: torch.zeros_like(v) for k, v in d.items()} for s, d in state.items()}
    v2 = {s: {k: torch.zeros_like(v) for k, v in d.items()} for s, d in state.items()}
    for step in range(fx["n_steps"]):
        gamma, pts = replay_rng(fx, step)
        losses, grads = orc.step_grads(state, batch_of(fx, step), cfg, gamma, pts)
        ref_l = fx["losses"][step]
        assert losses["img_loss"].item() == pytest.approx(ref_l["loss/img_loss"], rel=2e-5)
        assert losses["ue_loss"].item() == pytest.approx(ref_l["loss/ue_loss"], rel=2e-5)
        assert losses["grad_loss"].item() == pytest.approx(ref_l["loss/grad_loss"], rel=1e-3, abs=1e-3)
        assert losses["total"].item() == pytest.approx(ref_l["loss/total_train_loss"], rel=1e-4)
        for extra in ("rigidity_loss", "sparsity_loss1", "sparsity_loss2", "alpha_loss"):
            assert (extra in losses) == ("loss/" + extra in ref_l), extra
            if extra in losses:
                assert losses[extra].item() == pytest.approx(ref_l["loss/" + extra], rel=5e-5), extra
        for slot, d in fx["grads"][step].items():
            for k, g_ref in d.items():
                g = grads[slot][k]
                if g_ref.abs().max() == 0:
                    assert g.abs().max() == 0, (slot, k)
                else:
                    assert rel(g, g_ref) < 2e-4, (step, slot, k, rel(g, g_ref))
        # Adam (free-running from the oracle's own state: errors would compound)
        for slot, d in state.items():
            for k in d:
                p, mm, vv = orc.adam_update(d[k], grads[slot][k], m[slot][k], v2[slot][k], step + 1, cfg.lr)
                d[k], m[slot][k], v2[slot][k] = p, mm, vv
                p_ref = fx["params"][step][slot][k]
                # the first Adam updates are lr * g / (|g| + eps): where |g| ~ eps = 1e-8 (the layered configs have
                # such entries) a 1e-4 relative gradient difference moves the update by a few % of lr = 1e-4
                assert torch.allclose(p, p_ref, rtol=0, atol=5e-6), (step, slot, k, (p - p_ref).abs().max())


def test_full_width_network_digest():
    """hidden = 512 (the shipped atlas width): losses + sampled gradient entries + norms."""
    fx = load_fixture("mf_static_h512")
    cfg = cfg_of(fx)
    state = {s: {k: v.clone() for k, v in d.items()} for s, d in fx["init_state"].items()}
    gamma, pts = replay_rng(fx, 0)
    losses, grads = orc.step_grads(state, batch_of(fx, 0), cfg, gamma, pts)
    assert losses["img_loss"].item() == pytest.approx(fx["losses"][0]["loss/img_loss"], rel=2e-5)
    for slot, d in fx["grads_digest"][0].items():
        for k, dg in d.items():
            g = grads[slot][k].reshape(-1)
            assert g.double().norm().item() == pytest.approx(dg["norm"], rel=5e-4, abs=1e-12), (slot, k)
            if dg["norm"] > 0:
                assert rel(g[dg["idx"]], dg["val"]) < 5e-4, (slot, k)


def test_centre_tap_offsets_get_no_gradient():
    """training.py:120 overwrites the centre tap with the unshifted coords, so the tanh offsets of tap
    K//2 (outputs 13 and 22 of the blur MLP) receive zero gradient."""
    fx = load_fixture("mf_static_small")
    g = fx["grads"][0]["blur"]
    # last layer rows 13 and 22 (= 9 + 4, 18 + 4)
    w = g["layers.3.weight"]
    assert w[13].abs().max() == 0 and w[22].abs().max() == 0
    assert w[12].abs().max() > 0
