"""Data-parallel plumbing: the pixel batch of a step shards across ranks (all K taps of a pixel stay on one
rank), every rank normalises its loss by the GLOBAL pixel count, gradients are summed with ONE all-reduce
of the flat gradient buffer, and the (identical) fused Adam runs on every rank -- no parameter broadcast.

The reference has no distributed code; this is the one collective the build adds."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """[lo, hi) of the items rank `rank` owns out of n (contiguous, remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(model_input, gt, rank, world):
    """Slices a step's batch (reference DataLoader layout, [1,B,...]) to this rank's pixels."""
    B = model_input["coords"].shape[1]
    lo, hi = shard_range(B, rank, world)
    mi = {k: (v[:, lo:hi] if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == B else v)
          for k, v in model_input.items()}
    g = {k: (v[:, lo:hi] if torch.is_tensor(v) and v.dim() >= 2 and v.shape[1] == B else v) for k, v in gt.items()}
    return mi, g


def loss_weights(local_pixels, global_pixels, world):
    """(weight of the local image-loss mean, weight of pixel-independent terms) that make the SUM over ranks of
    the per-rank losses equal the single-process loss on the union batch."""
    return local_pixels / float(global_pixels), 1.0 / float(world)


def allreduce_flat_gradients(flat_grad, group=None):
    """The step's only collective: in-place SUM of the flat gradient buffer across ranks."""
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad
