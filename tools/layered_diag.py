"""Per-parameter gradient error of LayeredTrainStep against the oracle with each fused network family swapped
for plain fp32 torch ops in turn (diagnostic; GPU).   python tools/layered_diag.py video_hdr"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_layered_cpu as T  # noqa: E402
import test_layered_gpu as G  # noqa: E402
from neucam_b200 import modules, relu_mlp  # noqa: E402
from neucam_b200.layered import LayeredTrainStep  # noqa: E402
from oracle_replay import orc, rel  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "video_hdr"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 700
fused_sine = modules.FCBlock.forward
fused_ok = relu_mlp.eligible
for mode in ("fused relu + fused sine", "torch relu + fused sine", "fused relu + torch sine", "torch relu + torch sine"):
    relu_mlp.eligible = fused_ok if mode.startswith("fused relu") else (lambda m: False)
    modules.FCBlock.forward = fused_sine if mode.endswith("fused sine") else T._torch_sine_block
    models, state, opts, batch, cfg = G._setup(kind, B)
    step = LayeredTrainStep(models, G.SHAPES[kind], opts, device="cuda")
    gamma = torch.tensor([37.0]) if opts.use_random_gamma else None
    step.forward_backward(batch, {"img": batch["img"]}, gamma)
    dev_state = {s: {k: v.cuda() for k, v in d.items()} for s, d in state.items()}
    dev_batch = {k: v.cuda() for k, v in batch.items()}
    _, ref_g = orc.step_grads(dev_state, dev_batch, cfg, gamma.cuda() if gamma is not None else None)
    errs = []
    for slot, d in ref_g.items():
        for k, g_ref in d.items():
            g = dict(models[slot].named_parameters())[k].grad
            if g_ref.abs().max() > 0:
                errs.append((rel(g, g_ref), f"{slot}.{k}"))
    errs.sort(reverse=True)
    print(mode, " | ".join(f"{n} {e:.2e}" for e, n in errs[:6]))
