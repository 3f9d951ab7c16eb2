"""Run under torchrun on >= 2 GPUs: the data-parallel fused step (pixel shards + one NCCL all-reduce of the
flat gradient) must take the same step as one GPU on the union batch."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import harness
from neucam_b200.dist import shard_batch
from neucam_b200.step import FusedTrainStep, StepOptions

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    shape, focal = (3, 40, 60), [-1.0, 0.0, 1.0]
    B = 1024 * world
    scene = harness.SyntheticScene(shape, focal, seed=5)
    gen = torch.Generator().manual_seed(9)
    batches = [scene.sample(B, gen) for _ in range(3)]
    gammas = [torch.tensor([10.0 + 7 * i]) for i in range(3)]
    opts = StepOptions(use_random_gamma=True)
    models = harness.build_static_models(shape, focal, hidden=512, seed=3)
    step = FusedTrainStep(models, shape, opts, B // world, device=dev, process_group=dist.group.WORLD)
    losses = []
    for (inp, gt), g in zip(batches, gammas):
        mi, gg = shard_batch(inp, gt, rank, world)
        l = step.step(mi, gg, gamma=g)
        t = l["img_loss"].detach().clone()
        dist.all_reduce(t)
        losses.append(t.item() + l["ue_loss"].item())
    torch.cuda.synchronize()
    flat = step.params.flat.clone()
    # every rank must hold identical parameters (no broadcast needed)
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool((ref == flat).all().item())
    if rank == 0:
        models1 = harness.build_static_models(shape, focal, hidden=512, seed=3)
        single = FusedTrainStep(models1, shape, opts, B, device=dev)
        l1 = []
        for (inp, gt), g in zip(batches, gammas):
            l1.append(single.step(inp, gt, gamma=g)["total"].item())
        torch.cuda.synchronize()
        d = (single.params.flat - flat).abs().max().item()
        print(f"world={world}: losses dp {losses} vs single {l1}; max |param diff| after 3 steps {d:.3e}; "
              f"ranks identical: {same}")
        assert all(abs(a - b) <= 2e-3 * abs(b) for a, b in zip(losses, l1)), (losses, l1)
        assert d <= 3e-4, d
    assert same
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
