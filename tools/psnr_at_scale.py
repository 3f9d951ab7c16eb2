"""PSNR after the same number of iterations on the same batches, at BASELINE config shapes (the measurement behind
tests/test_psnr_parity.py).

    python tools/psnr_at_scale.py mf_static [iters]     # 2 x 300 x 450, B = 8100, K = 9 (config_mf_static.txt)
    python tools/psnr_at_scale.py me_dynamic [iters]    # 3 x 448 x 680, B = 54 835, uv_b + rigidity + gamma

Runs: the fused step (twice: must be bit-identical), the oracle (reference semantics, torch fp32 on the GPU, TF32 off)
on the same injected index stream, and the oracle again with the pixels of every batch PERMUTED (same set, other fp32
summation order) -- the reference's own sensitivity to arithmetic order.  Prints LDR PSNR (utils.py:198) over all
frames and the all-in-focus HDR PSNR with evaluate_hdr's normalisation (utils.py:733-736).
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from psnr_harness import gammas, index_stream, make_case, metrics, run_fused, run_oracle  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "mf_static"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    n_perm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    case = make_case(kind)
    n_total = case["shape"][0] * case["shape"][1] * case["shape"][2]
    idxs, gams = index_stream(n_total, case["B"], iters, 4), gammas(iters, 5)
    out = {"kind": kind, "iters": iters, "B": case["B"], "shape": list(case["shape"])}
    t0 = time.time()
    a = run_fused(case, idxs, gams)
    b = run_fused(case, idxs, gams)
    out["fused_bit_identical"] = all(torch.equal(a[s][k], b[s][k]) for s in a for k in a[s])
    out["fused"] = metrics(case, a)
    out["fused_seconds"] = (time.time() - t0) / 2
    t0 = time.time()
    out["oracle"] = metrics(case, run_oracle(case, idxs, gams))
    out["oracle_seconds"] = time.time() - t0
    out["oracle_permuted"] = [metrics(case, run_oracle(case, idxs, gams, permute_seed=100 + j)) for j in range(n_perm)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
