"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding, global loss normalisation,
the single all-reduce over the flat gradient buffer (neucam_b200.step.FlatParameters + neucam_b200.dist).
Per-rank gradients come from the oracle (the CUDA kernels cannot run here); the check is that the
all-reduced flat buffer equals the single-process gradient of the union batch, in flat-buffer order."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import neucam_oracle as orc
    from neucam_b200 import harness
    from neucam_b200.dist import allreduce_flat_gradients, loss_weights, shard_batch
    from neucam_b200.step import FlatParameters

    shape, focal, B = (3, 10, 14), [-1.0, 0.0, 1.0], 41          # odd B: ragged shards
    models = harness.build_static_models(shape, focal, hidden=64, seed=0)
    with torch.no_grad():
        models["exp"].exps.copy_(torch.linspace(-0.3, 0.4, 3))
    state = {s: {k: v.detach().clone() for k, v in m.state_dict().items()} for s, m in models.items() if m is not None}
    scene = harness.SyntheticScene(shape, focal, seed=2)
    inp, gt = scene.sample(B, torch.Generator().manual_seed(3))
    gamma = torch.tensor([17.0])
    cfg = orc.StepConfig(shape, use_random_gamma=True, with_grad_loss=False)

    fp = FlatParameters(models, "cpu")
    mi, g = shard_batch(inp, gt, rank, world)
    b_local = mi["coords"].shape[1]
    iw, uw = loss_weights(b_local, B, world)
    batch = dict(mi)
    batch["img"] = g["img"]
    losses, grads = orc.step_grads(state, batch, cfg, gamma, img_weight=iw, ue_weight=uw)
    for slot, name, p, off, n in fp.entries:
        fp.grad[off:off + n].copy_(grads[slot][name].reshape(-1))
    allreduce_flat_gradients(fp.grad)
    total = losses["total"].detach().clone()
    dist.all_reduce(total)

    if rank == 0:
        full = dict(inp)
        full["img"] = gt["img"]
        l1, g1 = orc.step_grads(state, full, cfg, gamma)
        ref = torch.zeros_like(fp.grad)
        for slot, name, p, off, n in fp.entries:
            ref[off:off + n].copy_(g1[slot][name].reshape(-1))
        torch.save({"err": (fp.grad - ref).abs().max().item(), "scale": ref.abs().max().item(),
                    "loss_sum": total.item(), "loss_ref": l1["total"].item(), "b_local": b_local},
                   os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_union_batch_gradient(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    assert r["b_local"] == 21                                   # 41 px -> 21 + 20
    assert r["err"] <= 2e-5 * r["scale"] + 1e-9
    assert r["loss_sum"] == pytest.approx(r["loss_ref"], rel=1e-5)


def test_shard_range_partitions_exactly():
    from neucam_b200.dist import shard_range
    for n in (1, 7, 41, 42525):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
