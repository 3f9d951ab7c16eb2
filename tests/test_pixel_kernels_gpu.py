"""GPU parity of the per-pixel kernels that were re-written as tiled SIMT GEMMs / one-pass reductions in round 2, at
the sizes where their tiling has edges (64-pixel groups, persistent grids, partial tiles), against the oracle
(oracle/neucam_oracle.py) under autograd:

  neucam_blur_fwd / neucam_blur_bwd       blur-kernel SimpleMLP + defocus taps (training.py:99-121) and their backward
  neucam_relu_mlp_head_prep               dpre = dy * f'(y), output-bias gradient, loss scale
"""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import neucam_oracle as orc  # noqa: E402  (the checker; tests may import it)

from neucam_b200 import harness, ops, pixel_ops  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _oracle_taps(blur_sd, focus, coords, coord_idx, shape, offset_size):
    """Tap positions [K,B,2] and PSF weights [K,B] exactly as orc.step_losses forms them (training.py:99-121)."""
    N, H, W = shape
    K = 9
    c = coords.view(1, -1, 3)
    patch = orc.neighbour_patch(c, H, W, K)
    f = orc.per_frame_gather(focus, coord_idx, H, W).unsqueeze(0)
    kernel_out = orc.simple_mlp(blur_sd, torch.cat((f, c[..., 1:3]), -1), "both", 3)
    py = patch[:, :, 1] + kernel_out[:, K:2 * K].t() * (2 / (H - 1)) * offset_size
    px = patch[:, :, 2] + kernel_out[:, 2 * K:3 * K].t() * (2 / (W - 1)) * offset_size
    centre = torch.zeros(K, 1, dtype=torch.bool)
    centre[K // 2] = True
    py = torch.where(centre, c[..., 1].expand(K, -1), py)
    px = torch.where(centre, c[..., 2].expand(K, -1), px)
    return torch.stack((py, px), -1), kernel_out[:, :K].t()


@pytest.mark.parametrize("B", [1, 63, 64, 65, 130, 1000, 64 * 148 * 3 + 7])
def test_blur_taps_forward_and_backward_match_oracle(B):
    shape = (3, 40, 56)
    N, H, W = shape
    models = harness.build_static_models(shape, [-1.0, 0.0, 1.0], hidden=512, seed=3)
    blur, focal = models["blur"], models["focal"]
    with torch.no_grad():      # away from the init point: every layer's gradient is exercised
        for p in blur.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=torch.Generator().manual_seed(p.numel())))
    g = torch.Generator().manual_seed(B)
    idx = torch.randint(0, N * H * W, (B,), generator=g)
    grid = harness.mgrid(shape)
    coords = grid[idx].contiguous()
    d_uv = torch.randn(9, B, 2, generator=g)
    d_kw = torch.randn(9, B, generator=g)

    # oracle under autograd (fp32 CPU)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in blur.state_dict().items()}
    foc = focal.focus.detach().clone().requires_grad_(True)
    uv_ref, kw_ref = _oracle_taps(sd, foc.view(-1), coords, idx, shape, 5.0)
    ((uv_ref * d_uv).sum() + (kw_ref * d_kw).sum()).backward()

    blur_c, focal_c = blur.cuda(), focal.cuda()
    uv, kw = pixel_ops.blur_taps(coords.cuda(), idx.cuda(), focal_c.focus, blur_c, shape, 9, 5.0)
    ((uv * d_uv.cuda()).sum() + (kw * d_kw.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert (uv.cpu() - uv_ref).abs().max().item() <= 1e-5            # coordinates in [-1, 1]
    assert (kw.cpu() - kw_ref).abs().max().item() <= 1e-5
    for k, p in blur_c.named_parameters():
        assert rel(p.grad.cpu(), sd[k].grad) <= 1e-4, (k, B)
    assert rel(focal_c.focus.grad.cpu().view(-1), foc.grad.view(-1)) <= 1e-4


def test_blur_backward_is_bit_reproducible_across_launches():
    """One partial record per block of a persistent grid, every entry owned by one thread, groups in order: no atomics."""
    shape = (3, 40, 56)
    B = 5000
    models = harness.build_static_models(shape, [-1.0, 0.0, 1.0], hidden=512, seed=4)
    blur, focal = models["blur"].cuda(), models["focal"].cuda()
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, shape[0] * shape[1] * shape[2], (B,), generator=g)
    coords = harness.mgrid(shape)[idx].contiguous().cuda()
    d_uv, d_kw = torch.randn(9, B, 2, generator=g).cuda(), torch.randn(9, B, generator=g).cuda()
    runs = []
    for _ in range(2):
        for p in list(blur.parameters()) + [focal.focus]:
            p.grad = None
        uv, kw = pixel_ops.blur_taps(coords, idx.cuda(), focal.focus, blur, shape, 9, 5.0)
        ((uv * d_uv).sum() + (kw * d_kw).sum()).backward()
        runs.append([p.grad.clone() for p in list(blur.parameters()) + [focal.focus]])
    for a, b in zip(*runs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("out_dim,out_kind,Q", [(2, 1, 1000), (1, 2, 4097), (3, 0, 129), (4, 1, 300000), (2, 1, 1)])
def test_head_prep_matches_torch(out_dim, out_kind, Q):
    g = torch.Generator().manual_seed(Q + out_dim)
    pre = torch.randn(Q, out_dim, generator=g)
    y = {0: pre, 1: torch.tanh(pre), 2: torch.sigmoid(pre)}[out_kind]
    dy = torch.randn(Q, out_dim, generator=g) * 3e-4
    want = {0: dy, 1: dy * (1 - y * y), 2: dy * y * (1 - y)}[out_kind]
    base = want.double().sum(0).abs().max().item()          # the kernel ACCUMULATES: start from a value of that size
    db = torch.full((out_dim,), base, device="cuda")
    dpre, scale = ops.relu_head_prep(dy.cuda(), y.cuda() if out_kind else None, out_kind, db)
    torch.cuda.synchronize()
    assert (dpre.cpu() - want).abs().max().item() <= 1e-9
    assert (db.cpu().double() - base - want.double().sum(0)).abs().max().item() <= 1e-5 * max(base, 1e-30) + 1e-12
    m = want.abs().max().item()
    assert scale.item() == 2.0 ** math.floor(math.log2(ops.SCALE_TARGET / m))
    # deterministic: fixed-order partial sums
    db2 = torch.full((out_dim,), base, device="cuda")
    ops.relu_head_prep(dy.cuda(), y.cuda() if out_kind else None, out_kind, db2)
    assert torch.equal(db, db2)
