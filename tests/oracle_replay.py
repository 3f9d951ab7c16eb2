"""Shared helper: replays a golden fixture's step with the oracle (CPU)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import neucam_oracle as orc  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def replay_rng(fx, step):
    """The CPU RNG draws one reference step makes, in order (training.py:181, 248-249)."""
    torch.manual_seed(fx["step_seeds"][step])
    gamma = None
    if fx["use_random_gamma"]:
        gamma = (torch.rand(1) * 100 + 1.0)
    rad = torch.rand(10000, 3) * 5
    expo = torch.rand(10000, 1) * 6 - 3.0
    return gamma, (rad, expo)


def cfg_of(fx):
    return orc.StepConfig(fx["shape"], patch_size=fx["opt"]["patch_size"], offset_size=fx["opt"]["offset_size"],
                          use_random_gamma=fx["use_random_gamma"], fixed_value=fx["opt"]["fixed_value"],
                          lr=fx["opt"]["lr"], uv_mapping_scale=fx["opt"].get("uv_mapping_scale", 1.0),
                          rigidity_loss_w=fx["opt"].get("rigidity_loss_w"),
                          sparsity_loss_w=fx["opt"].get("sparsity_loss_w"),
                          alpha_loss_w=fx["opt"].get("alpha_loss_w"), nets=fx.get("nets"))


def batch_of(fx, step):
    inp, gt = fx["batches"][step]
    b = dict(inp)
    b["img"] = gt["img"]
    return b


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
