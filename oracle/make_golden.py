"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference
(`/root/reference`, through oracle/ref_shim.py) in the build container.

    python oracle/make_golden.py            # rewrites every fixture

For each case the real `training.train` loop (training.py:18-317) is driven for a few steps over
INJECTED batches (the reference never seeds its RNG, so index/gamma streams are injected rather than
reproduced: SURVEY.md section 7 hard part 7).  Captured per step, without editing the reference:
  * every logged loss           (SummaryWriter.add_scalar recorder)
  * every parameter gradient    (torch.optim.Adam.step wrapper, before the update)
  * every parameter after Adam  (same wrapper, after the update)
The CPU RNG is re-seeded per step with a recorded seed so that the oracle can replay the draws the
step makes (`torch.rand(1)` gamma, then `torch.rand(10000,3)`, `torch.rand(10000,1)`;
training.py:181,248-249).
"""
import argparse
import math
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

SLOTS = ("atlas_b", "tm", "uv_b", "uv_f", "blur", "atlas_f", "alpha", "exp", "focal", "noise", "depth")


def make_opt(**kw):
    base = dict(
        num_epochs=1, lr=1e-4, steps_til_summary=1000000, epochs_til_ckpt=1000000, config_filepath=None,
        patch_size=9, offset_size=5, patch_sample=False, use_depth=False, use_random_gamma=False,
        flow_loss=False, flow_loss_steps=0, rigidity_loss=False, rigidity_loss_steps=0, rigidity_loss_w=None,
        sparsity_loss=False, sparsity_loss_w=None, alpha_loss=False, alpha_loss_w=None, mask_loss=False,
        mask_loss_steps=0, mask_loss_w=None, fixed_value=0.0, start_freeze_tm=-1, uv_mapping_scale=1.0)
    base.update(kw)
    return argparse.Namespace(**base)


def build_reference_models(shape, hidden, focal_list, seed):
    """Model construction exactly as experiment_scripts/train_video.py:27-122 does for the static
    multi-focus / multi-exposure configs (deblur + pre_hdr + learn_exp + learn_focal)."""
    modules = ref_shim.load("modules")
    torch.manual_seed(seed)
    N, H, W = shape
    m = {k: None for k in SLOTS}
    m["atlas_b"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=hidden, num_hidden_layers=3,
                                        in_features=2, out_features=3, sidelength=(H, W), num_frequencies=7)
    m["exp"] = modules.LearnExps(N)
    m["focal"] = modules.LearnFocal(list(focal_list), True)
    m["blur"] = modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                  num_frequencies=7, skip_layer=[4], nonlinear="both")
    m["tm"] = modules.TonemapNet()
    # non-trivial exposure / focus / CRF state so that every gradient path is exercised
    with torch.no_grad():
        m["exp"].exps.copy_(torch.linspace(-0.4, 0.5, N))
        for p in m["tm"].parameters():
            p.mul_(1.5)
    return m


# hyper-parameters of the uv / alpha SimpleMLPs per dynamic config (configs/config_{me_dynamic,video_deblur,
# video_hdr}.txt), with the hidden widths shrunk so that the fixtures stay small
DYNAMIC = {
    "me_dynamic": dict(
        blur=False, split=False, use_random_gamma=True,
        uv_b=dict(num_layer=4, use_encoding=False, num_frequencies=None, skip=[4], last="tanh"),
        opt=dict(uv_mapping_scale=1.0, rigidity_loss=True, rigidity_loss_w=1.0, rigidity_loss_steps=50000)),
    "video_deblur": dict(
        blur=True, split=True, use_random_gamma=False,
        uv_b=dict(num_layer=4, use_encoding=True, num_frequencies=3, skip=[4], last="tanh"),
        uv_f=dict(num_layer=4, use_encoding=True, num_frequencies=3, skip=[4], last="tanh"),
        alpha=dict(num_layer=4, use_encoding=True, num_frequencies=4, skip=[4], last="sigmoid"),
        opt=dict(uv_mapping_scale=1.0, rigidity_loss=True, rigidity_loss_w=1e-2, rigidity_loss_steps=20000,
                 sparsity_loss=True, sparsity_loss_w=1e-3, alpha_loss=True, alpha_loss_w=5e-2)),
    "video_hdr": dict(
        blur=False, split=True, use_random_gamma=False,
        uv_b=dict(num_layer=4, use_encoding=False, num_frequencies=None, skip=[4], last="tanh"),
        uv_f=dict(num_layer=6, use_encoding=True, num_frequencies=3, skip=[4], last="tanh"),
        alpha=dict(num_layer=4, use_encoding=True, num_frequencies=3, skip=[4], last="sigmoid"),
        opt=dict(uv_mapping_scale=0.8, rigidity_loss=True, rigidity_loss_w=1e-2, rigidity_loss_steps=500000,
                 sparsity_loss=True, sparsity_loss_w=5e-2)),
}


def build_reference_models_dynamic(kind, shape, hidden, mlp_hidden, seed):
    """experiment_scripts/train_video.py:27-122 for the uv-mapping configs (me_dynamic, video_deblur, video_hdr)."""
    modules = ref_shim.load("modules")
    spec = DYNAMIC[kind]
    torch.manual_seed(seed)
    N, H, W = shape
    m = {k: None for k in SLOTS}

    def mlp(h, out_dim):
        return modules.SimpleMLP(hidden_dim=mlp_hidden, num_layer=h["num_layer"], in_dim=3, out_dim=out_dim,
                                 use_encoding=h["use_encoding"], num_frequencies=h["num_frequencies"] or 7,
                                 skip_layer=h["skip"], nonlinear=h["last"])

    m["atlas_b"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=hidden, num_hidden_layers=3,
                                        in_features=2, out_features=3, sidelength=(H, W), num_frequencies=7)
    m["uv_b"] = mlp(spec["uv_b"], 2)
    if spec["split"]:
        m["atlas_f"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=hidden, in_features=2,
                                            out_features=3, sidelength=(H, W), num_frequencies=7)
        m["uv_f"] = mlp(spec["uv_f"], 2)
        m["alpha"] = mlp(spec["alpha"], 1)
    m["exp"] = modules.LearnExps(N)
    if spec["blur"]:
        m["focal"] = modules.LearnFocal(N)
        m["blur"] = modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                      num_frequencies=7, skip_layer=[4], nonlinear="both")
    m["tm"] = modules.TonemapNet()
    with torch.no_grad():
        m["exp"].exps.copy_(torch.linspace(-0.4, 0.5, N))
        if m["focal"] is not None:
            m["focal"].focus.copy_(torch.linspace(-0.6, 0.7, N).view_as(m["focal"].focus))
        for p in m["tm"].parameters():
            p.mul_(1.5)
    return m


def make_dynamic_case(name, kind, shape, hidden, mlp_hidden, B, n_steps, seed):
    spec = DYNAMIC[kind]
    opt = make_opt(use_random_gamma=spec["use_random_gamma"], **spec["opt"])
    models = build_reference_models_dynamic(kind, shape, hidden, mlp_hidden, seed)
    init_state = state_of(models)
    batches = make_batches(shape, B, n_steps, seed + 1, spec["use_random_gamma"])
    step_seeds = [1000 + seed * 10 + i for i in range(n_steps)]
    rec = run_reference(models, batches, shape, opt, step_seeds)
    nets = {s: dict(last=spec[s]["last"], skip=tuple(spec[s]["skip"]),
                    num_frequencies=spec[s]["num_frequencies"] if spec[s]["use_encoding"] else None)
            for s in ("uv_b", "uv_f", "alpha") if s in spec}
    fx = {
        "name": name, "kind": kind, "shape": shape, "hidden": hidden, "mlp_hidden": mlp_hidden, "B": B,
        "n_steps": n_steps, "seed": seed, "use_random_gamma": spec["use_random_gamma"], "step_seeds": step_seeds,
        "opt": {"patch_size": opt.patch_size, "offset_size": opt.offset_size, "lr": opt.lr,
                "fixed_value": opt.fixed_value, "uv_mapping_scale": opt.uv_mapping_scale,
                "rigidity_loss_w": opt.rigidity_loss_w if opt.rigidity_loss else None,
                "sparsity_loss_w": opt.sparsity_loss_w if opt.sparsity_loss else None,
                "alpha_loss_w": opt.alpha_loss_w if opt.alpha_loss else None},
        "nets": nets,
        "batches": [({k: v for k, v in b[0].items() if k != "idx"}, b[1]) for b in batches],
        "losses": rec["losses"], "init_state": init_state, "grads": rec["grads"], "params": rec["params"],
        "torch_version": torch.__version__,
    }
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(fx, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, losses step0 = {rec['losses'][0]}")


def make_batches(shape, B, n_steps, seed, with_gamma_weights):
    """Batches in the layout Implicit3DWrapper.__getitem__ + DataLoader(batch_size=1) produce
    (dataio.py:356-393): coords [1,B,3] from get_mgrid, coord_idx [1,B] int64, img [1,B,3] in [-1,1]."""
    dataio = ref_shim.load("dataio")
    N, H, W = shape
    g = torch.Generator().manual_seed(seed)
    mgrid = dataio.get_mgrid(shape, dim=3).view(-1, 3)
    video = torch.rand(N * H * W, 3, generator=g) * 2 - 1
    batches = []
    for _ in range(n_steps):
        idx = torch.randint(0, N * H * W, (B,), generator=g)
        inp = {"idx": torch.zeros(1, dtype=torch.int64), "coords": mgrid[idx][None], "coord_idx": idx[None]}
        if with_gamma_weights:
            inp["gamma_weights"] = torch.rand(B, 1, generator=g)[None]
        batches.append((inp, {"img": video[idx][None]}))
    return batches


def run_reference(models, batches, shape, opt, step_seeds):
    training = ref_shim.load("training")
    records = {"losses": [], "grads": [], "params": []}

    class Recorder:
        def __init__(self, *a, **k):
            self.cur = {}

        def add_scalar(self, name, value, step):
            if len(records["losses"]) <= step:
                records["losses"].append({})
            records["losses"][step][name] = float(value)

    def named_params():
        out = {}
        for slot, mod in models.items():
            if mod is not None:
                out[slot] = {k: v for k, v in mod.named_parameters()}
        return out

    orig_step = torch.optim.Adam.step
    state = {"i": 0}

    def step_wrapper(self, *a, **k):
        records["grads"].append({s: {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
                                     for n, p in d.items()} for s, d in named_params().items()})
        r = orig_step(self, *a, **k)
        records["params"].append({s: {n: p.detach().clone() for n, p in d.items()}
                                  for s, d in named_params().items()})
        state["i"] += 1
        if state["i"] < len(step_seeds):
            torch.manual_seed(step_seeds[state["i"]])
        return r

    training.SummaryWriter = Recorder
    torch.optim.Adam.step = step_wrapper
    tmp = tempfile.mkdtemp(prefix="neucam_golden_")
    try:
        torch.manual_seed(step_seeds[0])
        training.train(models=models, train_dataloader=batches, model_dir=os.path.join(tmp, "run"),
                       loss_fn=ref_shim.load("loss_functions").image_mse, summary_fn=lambda *a, **k: None,
                       video_shape=shape, opt=opt)
    finally:
        torch.optim.Adam.step = orig_step
        shutil.rmtree(tmp, ignore_errors=True)
    return records


def state_of(models):
    return {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}


def digest(t, n=64, seed=0):
    """Norm + n sampled entries of a tensor (for fixtures of full-size networks)."""
    flat = t.reshape(-1)
    g = torch.Generator().manual_seed(seed + flat.numel())
    idx = torch.randint(0, flat.numel(), (min(n, flat.numel()),), generator=g)
    return {"norm": flat.double().norm().item(), "idx": idx, "val": flat[idx].clone(), "shape": tuple(t.shape)}


def make_case(name, shape, hidden, B, n_steps, use_random_gamma, focal_list, seed, full):
    opt = make_opt(use_random_gamma=use_random_gamma)
    models = build_reference_models(shape, hidden, focal_list, seed)
    init_state = state_of(models)
    batches = make_batches(shape, B, n_steps, seed + 1, use_random_gamma)
    step_seeds = [1000 + seed * 10 + i for i in range(n_steps)]
    rec = run_reference(models, batches, shape, opt, step_seeds)
    fx = {
        "name": name, "shape": shape, "hidden": hidden, "B": B, "n_steps": n_steps, "seed": seed,
        "use_random_gamma": use_random_gamma, "focal_list": list(focal_list), "step_seeds": step_seeds,
        "opt": {"patch_size": opt.patch_size, "offset_size": opt.offset_size, "lr": opt.lr,
                "fixed_value": opt.fixed_value},
        "batches": [({k: v for k, v in b[0].items() if k != "idx"}, b[1]) for b in batches],
        "losses": rec["losses"],
        "torch_version": torch.__version__,
    }
    if full:
        fx["init_state"] = init_state
        fx["grads"] = rec["grads"]
        fx["params"] = rec["params"]
    else:
        # full-size network: the initial state is reproducible from `seed` through the reference's own
        # constructors, which tests cannot call on the GPU box -> store it as fp32 anyway for atlas-free
        # slots, and store digests of gradients / updated parameters.
        fx["init_state"] = init_state
        fx["grads_digest"] = [{s: {k: digest(v) for k, v in d.items()} for s, d in g.items()} for g in rec["grads"]]
        fx["params_digest"] = [{s: {k: digest(v) for k, v in d.items()} for s, d in g.items()} for g in rec["params"]]
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(fx, path)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, losses step0 = {rec['losses'][0]}")


def main():
    if not ref_shim.available():
        raise SystemExit("reference tree not present; golden fixtures can only be regenerated in the build container")
    if len(sys.argv) > 1 and sys.argv[1] == "--dynamic-only":
        make_dynamic_case("me_dynamic_small", "me_dynamic", (3, 12, 16), 64, 32, 48, 2, 3)
        make_dynamic_case("video_deblur_small", "video_deblur", (4, 10, 14), 64, 32, 40, 2, 4)
        make_dynamic_case("video_hdr_small", "video_hdr", (4, 10, 14), 64, 32, 48, 2, 5)
        return
    make_case("mf_static_small", (2, 12, 18), 64, 48, 3, False, (-1.0, 1.0), 0, full=True)
    make_case("mfme_gamma_small", (3, 10, 14), 64, 40, 3, True, (-1.0, 0.0, 1.0), 1, full=True)
    make_case("mf_static_h512", (2, 30, 45), 512, 300, 2, False, (-1.0, 1.0), 2, full=False)
    make_dynamic_case("me_dynamic_small", "me_dynamic", (3, 12, 16), 64, 32, 48, 2, 3)
    make_dynamic_case("video_deblur_small", "video_deblur", (4, 10, 14), 64, 32, 40, 2, 4)
    make_dynamic_case("video_hdr_small", "video_hdr", (4, 10, 14), 64, 32, 48, 2, 5)


if __name__ == "__main__":
    main()
