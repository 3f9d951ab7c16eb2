"""Run under torchrun on >= 2 GPUs: the data-parallel LAYERED step (pixel shards, one NCCL all-reduce of the flat
gradient between two CUDA graphs) must take the same steps as one GPU on the union batch, and leave identical
parameters on every rank."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neucam_b200 import harness
from neucam_b200.layered import LayeredTrainStep
from neucam_b200.step import StepOptions


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    kind, shape = "video_deblur", (6, 36, 64)
    spec = harness.LAYERED_CONFIGS[kind]
    B = 512 * world
    g = torch.Generator().manual_seed(9)
    grid = harness.mgrid(shape)
    batches = []
    for _ in range(3):
        idx = torch.randint(0, grid.shape[0], (B,), generator=g)
        batches.append({"coords": grid[idx][None], "coord_idx": idx[None], "img": torch.rand(1, B, 3, generator=g) * 2 - 1})
    opts = StepOptions(lr=1e-4, **spec["options"])
    models = harness.build_layered_models(kind, shape, seed=3)
    step = LayeredTrainStep(models, shape, opts, device=dev, process_group=dist.group.WORLD, global_pixels=B)
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    for b in batches:
        part = {k: v[:, lo:hi].contiguous() for k, v in b.items()}
        step.graph_step(part, {"img": part["img"]})          # two graphs around the eager all-reduce
    torch.cuda.synchronize()
    flat = step.params.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool((ref == flat).all().item())
    gathered = [None] * world
    dist.all_gather_object(gathered, same)
    if rank == 0:
        models1 = harness.build_layered_models(kind, shape, seed=3)
        single = LayeredTrainStep(models1, shape, opts, device=dev)
        for b in batches:
            single.step(b, {"img": b["img"]})
        torch.cuda.synchronize()
        d = (single.params.flat - flat).abs()
        print(f"ranks identical: {gathered}; vs single GPU on the union batch: max |dp| {d.max().item():.2e}, "
              f"mean {d.mean().item():.2e} after 3 Adam steps of lr 1e-4")
        assert all(gathered) and d.mean().item() < 2e-6 and d.max().item() < 3e-4
        print("dp_check_layered ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
