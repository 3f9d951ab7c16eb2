"""TEST INFRASTRUCTURE ONLY -- golden vectors for the rows AROUND the step, produced by
running the UNMODIFIED reference (/root/reference through oracle/ref_shim.py) in the build container:

    python oracle/make_golden_aux.py

  tests/golden/warmups.pt        utils.warm_tm / warm_mapping / warm_alpha (utils.py:265-334) on seeded models and a
                                 seeded CPU RNG: initial state, per-iteration losses, final state
  tests/golden/ref_checkpoint_small.pth  a checkpoint written by the reference's OWN utils.save_model
                                 (utils.py:645-673) for a two-layer model set with every slot the configs use
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def clone_state(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def digest(state):
    """Initial states are re-created from the seed by the tests (the package's modules reproduce the reference's
    initialisation bit for bit: tests/test_modules_cpu.py); the fixture only pins a checksum of them."""
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in state.items()}


def capture_losses(fn):
    """The reference prints its warm-up losses; parse them back."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn()
    losses = []
    for line in buf.getvalue().splitlines():
        if "pre-train loss:" in line:
            losses.append(float(line.split("pre-train loss:")[1].split(",")[0]))
    return out, torch.tensor(losses)


def main():
    modules = ref_shim.load("modules")
    utils = ref_shim.load("utils")
    out = {"torch_version": torch.__version__}

    # ---- warm_tm: TonemapNet, 25 iterations ----
    torch.manual_seed(11)
    tm = modules.TonemapNet()
    init = clone_state(tm)
    torch.manual_seed(12)
    _, losses = capture_losses(lambda: utils.warm_tm(tm, pretrain_iters=25, device="cpu"))
    out["warm_tm"] = dict(model_seed=11, rng_seed=12, iters=25, init_digest=digest(init), final=clone_state(tm), losses=losses)

    # ---- warm_mapping: the uv network of config_video_hdr (uv_f: 6 layers, encoding, skip) and of me_dynamic ----
    for name, kw in (("uv_enc_skip", dict(hidden_dim=256, num_layer=6, in_dim=3, out_dim=2, use_encoding=True,
                                          num_frequencies=3, skip_layer=[4], nonlinear="tanh")),):
        torch.manual_seed(21)
        net = modules.SimpleMLP(**kw)
        init = clone_state(net)
        shape = (3, 20, 30)
        torch.manual_seed(22)
        _, losses = capture_losses(lambda: utils.warm_mapping(net, shape, uv_mapping_scale=0.8, pretrain_iters=2,
                                                              device="cpu"))
        out["warm_mapping_" + name] = dict(model_seed=21, rng_seed=22, iters=2, shape=shape, kw=kw, scale=0.8,
                                           init_digest=digest(init), final=clone_state(net), losses=losses)

    # ---- warm_alpha: alpha network of config_video_deblur against a synthetic mask channel ----
    kw = dict(hidden_dim=256, num_layer=4, in_dim=3, out_dim=1, use_encoding=True, num_frequencies=4, skip_layer=[4],
              nonlinear="sigmoid")
    torch.manual_seed(31)
    net = modules.SimpleMLP(**kw)
    init = clone_state(net)
    shape = (3, 20, 30)
    g = torch.Generator().manual_seed(32)
    vid = torch.rand(shape[0], shape[1], shape[2], 4, generator=g).numpy().astype(np.float32)
    vid[..., 3] = (vid[..., 3] > 0.5).astype(np.float32)
    ds = types.SimpleNamespace(shape=shape, vid=vid)
    torch.manual_seed(33)
    _, losses = capture_losses(lambda: utils.warm_alpha(net, ds, pretrain_iters=2, device="cpu"))
    out["warm_alpha"] = dict(model_seed=31, rng_seed=33, iters=2, shape=shape, kw=kw, vid=torch.from_numpy(vid),
                             init_digest=digest(init), final=clone_state(net), losses=losses)
    torch.save(out, os.path.join(GOLDEN_DIR, "warmups.pt"))
    print("wrote warmups.pt:", {k: (v["losses"][:2].tolist(), v["losses"][-1].item()) for k, v in out.items()
                                if isinstance(v, dict)})

    # ---- a checkpoint written by the reference's own save_model (narrow networks keep the fixture small; every slot
    # the shipped configs use is present) ----
    H, W, N = 24, 36, 4
    slots = ("atlas_b", "tm", "uv_b", "uv_f", "blur", "atlas_f", "alpha", "exp", "focal", "noise", "depth")
    torch.manual_seed(42)
    small = {k: None for k in slots}
    small["atlas_b"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=64, num_hidden_layers=3,
                                            in_features=2, out_features=3, sidelength=(H, W), num_frequencies=7)
    small["atlas_f"] = modules.SingleBVPNet(type="sine", mode="mlp", hidden_features=64, in_features=2, out_features=3,
                                            sidelength=(H, W), num_frequencies=7)
    small["uv_b"] = modules.SimpleMLP(hidden_dim=32, num_layer=4, in_dim=3, out_dim=2, use_encoding=True,
                                      num_frequencies=3, skip_layer=[4], nonlinear="tanh")
    small["uv_f"] = modules.SimpleMLP(hidden_dim=32, num_layer=6, in_dim=3, out_dim=2, use_encoding=True,
                                      num_frequencies=3, skip_layer=[4], nonlinear="tanh")
    small["alpha"] = modules.SimpleMLP(hidden_dim=32, num_layer=4, in_dim=3, out_dim=1, use_encoding=True,
                                       num_frequencies=4, skip_layer=[4], nonlinear="sigmoid")
    small["exp"] = modules.LearnExps(N)
    small["focal"] = modules.LearnFocal(N)
    small["blur"] = modules.SimpleMLP(hidden_dim=64, num_layer=4, in_dim=3, out_dim=27, use_encoding=False,
                                      num_frequencies=7, skip_layer=[4], nonlinear="both")
    small["tm"] = modules.TonemapNet()
    with torch.no_grad():
        small["exp"].exps.copy_(torch.linspace(-0.4, 0.5, N))
        small["focal"].focus.copy_(torch.linspace(-1, 1, N))
    utils.save_model(small, os.path.join(GOLDEN_DIR, "ref_checkpoint_small.pth"))
    sz = os.path.getsize(os.path.join(GOLDEN_DIR, "ref_checkpoint_small.pth"))
    print("wrote ref_checkpoint_small.pth", sz, "bytes")


if __name__ == "__main__":
    main()
