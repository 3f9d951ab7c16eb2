"""GPU unit tests of csrc/losses.cu through the C ABI: the rigidity loss kernel (value + analytic gradient) against
the oracle's autograd restatement of utils.get_rigidity_loss (utils.py:402-454), and the device-side loss scale."""
import math

import pytest
import torch

from oracle_replay import rel

from neucam_b200 import ops, pixel_ops

pytestmark = pytest.mark.gpu


def reference_rigidity(uv_ref, uv_moved, shape, scale):
    """The 2x2 algebra of utils.py:421-451 on given mapping values (matrix form, as the reference writes it)."""
    n, h, w = shape
    du, dv = uv_ref[:, 0][None] - uv_moved[..., 0], uv_ref[:, 1][None] - uv_moved[..., 1]      # [2,B]: up, left
    du_dx, du_dy = du[1] * (h - 1) / 2, du[0] * (h - 1) / 2
    dv_dy, dv_dx = dv[0] * (w - 1) / 2, dv[1] * (w - 1) / 2
    J = torch.stack((torch.stack((du_dx, du_dy), -1), torch.stack((dv_dx, dv_dy), -1)), 1) / scale
    JtJ = J.transpose(1, 2) @ J
    a, b, c, d = JtJ[:, 0, 0] + 0.001, JtJ[:, 0, 1], JtJ[:, 1, 0], JtJ[:, 1, 1] + 0.001
    inv = torch.stack((torch.stack((d, -b), -1), torch.stack((-c, a), -1)), 1) / (a * d - b * c).view(-1, 1, 1)
    return (JtJ ** 2).sum((1, 2)).sqrt() + (inv ** 2).sum((1, 2)).sqrt()


@pytest.mark.parametrize("B,scale", [(1, 1.0), (777, 0.8), (20000, 1.0)])
def test_rigidity_kernel_value_and_gradient(B, scale):
    torch.manual_seed(B)
    shape = (5, 36, 64)
    pos = torch.rand(B, 2, device="cuda", dtype=torch.float64) * 2 - 1
    # a smooth mapping evaluated at the pixel and its upper / left neighbour, plus noise: well-conditioned Jacobians
    step = torch.tensor([[2 / 35, 0.0], [0.0, 2 / 63]], device="cuda", dtype=torch.float64)
    f = lambda p: 0.5 * p + 0.2 * torch.sin(3 * p.flip(-1)) + 0.1 * p * p
    uv_ref = f(pos).requires_grad_(True)
    uv_moved = torch.stack((f(pos - step[0]), f(pos - step[1]))).requires_grad_(True)
    want = reference_rigidity(uv_ref, uv_moved, shape, scale)
    coef = 0.37 / B
    (want.sum() * coef).backward()

    r32 = uv_ref.detach().float().requires_grad_(True)
    m32 = uv_moved.detach().float().requires_grad_(True)
    got = pixel_ops.rigidity_loss(r32, m32, torch.full((1,), coef, device="cuda"), shape, scale)
    got.backward()
    assert got.item() == pytest.approx((want.sum() * coef).item(), rel=2e-5)
    assert rel(r32.grad, uv_ref.grad) < 1e-4
    assert rel(m32.grad, uv_moved.grad) < 1e-4


def test_loss_scale_rule():
    for mx in (3e-7, 1.0, 15.9, 16.0, 16.1, 5e4):
        x = torch.zeros(1000, device="cuda")
        x[123] = -mx
        s = ops.loss_scale_from(x).item()
        assert s == 2.0 ** math.floor(math.log2(ops.SCALE_TARGET / mx)) or s in (2.0 ** -20, 2.0 ** 64)
        assert ops.SCALE_TARGET / 2 <= mx * s < ops.SCALE_TARGET * (1 + 1e-6) or s in (2.0 ** -20, 2.0 ** 64)
    assert ops.loss_scale_from(torch.zeros(10, device="cuda")).item() == 1.0
    assert ops.loss_scale_from(torch.full((10,), float("inf"), device="cuda")).item() == 1.0
