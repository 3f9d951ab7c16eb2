"""GPU unit test of the tcgen05/TMA building blocks (neucam_umma_probe): one 128xN tile GEMM in every
operand-layout / dtype combination the production kernels use, against torch.matmul on the same
16-bit-rounded inputs (fp32 accumulate both sides -> only summation order differs)."""
import ctypes

import pytest
import torch

from neucam_b200 import _lib

pytestmark = pytest.mark.gpu

CASES = [
    # (N, K, flags, id)
    (256, 256, 0, "kmajor_f16"),
    (128, 64, 0, "kmajor_f16_small"),
    (256, 128, 4 | 8, "kmajor_bf16"),
    (256, 256, 16, "a_manual_swizzle"),
    (256, 128, 1 | 2 | 4 | 8, "mnmajor_bf16"),
    (128, 256, 1 | 2, "mnmajor_f16"),
    (256, 128, 2 | 4 | 8, "a_k_b_mn_bf16"),
    (256, 128, 1 | 4 | 8, "a_mn_b_k_bf16"),
    (64, 64, 0, "n64"),
    (16, 64, 0, "n16"),
    (128, 64, 32, "a_in_tmem_small"),      # A operand read from tensor memory (tcgen05.st-written rows)
    (256, 256, 32, "a_in_tmem"),
]


def _run(N, K, flags):
    lib = _lib.debug_lib()      # the probes exist only in the debug build (include/neucam_b200_debug.h)
    g = torch.Generator(device="cpu").manual_seed(1234 + N + K + flags)
    a_dt = torch.bfloat16 if flags & 4 else torch.float16
    b_dt = torch.bfloat16 if flags & 8 else torch.float16
    A = torch.randn(128, K, generator=g).to(a_dt).cuda()
    B = torch.randn(N, K, generator=g).to(b_dt).cuda()
    a_in = A.t().contiguous() if flags & 1 else A.contiguous()
    b_in = B.t().contiguous() if flags & 2 else B.contiguous()
    D = torch.full((128, N), float("nan"), device="cuda", dtype=torch.float32)
    rc = lib.neucam_umma_probe(_lib.ptr(a_in), _lib.ptr(b_in), _lib.ptr(D), N, K, flags, _lib.cur_stream())
    _lib.check(rc, "neucam_umma_probe")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    return D, ref


@pytest.mark.parametrize("N,K,flags,name", CASES, ids=[c[3] for c in CASES])
def test_probe(N, K, flags, name):
    D, ref = _run(N, K, flags)
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-4 * scale, f"{name}: max abs err {err} (scale {scale})"


# NOTE (measured on B200, round 1): mixing operand formats inside kind::f16 (A fp16 x B bf16) raises
# cudaErrorIllegalInstruction, so every GEMM in this library uses ONE 16-bit format for both operands.
