"""`train()` with the reference's signature (training.py:18) driving the fused CUDA step.

Mirrors the cadence of the reference loop: per-step pipeline (training.py:95-271) = FusedTrainStep for the static
multi-focus / multi-exposure model set, LayeredTrainStep for the uv-mapping / two-layer ones; the freeze
schedule at `opt.start_freeze_tm` (training.py:300-314: a FRESH Adam with lr 1e-5, blur lr 1e-6, tm / exp /
focal no longer updated, mask weights switched on); checkpoints under the reference's keys every
`epochs_til_ckpt` epochs and at the end (training.py:86-87, 317).  What it does not reproduce: TensorBoard
writes and the summary renderer's image dumps (offline visualisation) -- `summary_fn`
is still called at the reference's cadence if given.  Losses stay on the device; `.item()` only happens when a
step is logged (the reference syncs once per loss per step, training.py:256-267).

Both model sets run through a captured CUDA graph (one graph per (mask, learning-rate groups) combination; the freeze
switch re-captures once).

Data parallel (`process_group`): the reference has no distributed code; here every rank must be handed ITS OWN
shard of each step's pixel batch by its DataLoader (neucam_b200.dist.shard_batch slices a global batch; B may differ
by one between ranks -- pass the global count through `opt.global_pixels` then).  The step constructors broadcast rank
0's parameters and optimizer state, the random gamma of a step is drawn on rank 0 and broadcast, and only rank 0
writes checkpoints / calls summary_fn.
"""
import os

import torch

from . import utils as nutils
from .layered import LayeredTrainStep, needs_layered_step
from .step import FusedTrainStep, StepOptions


def lr_groups_after_freeze():
    """training.py:300-313: atlas / uv / alpha at 1e-5, blur at 1e-6, tm / exp / focal frozen."""
    return {"atlas_b": 1e-5, "uv_b": 1e-5, "uv_f": 1e-5, "atlas_f": 1e-5, "alpha": 1e-5, "blur": 1e-6,
            "tm": None, "exp": None, "focal": None, "noise": None, "depth": None}


def train(models, train_dataloader, model_dir, loss_fn=None, summary_fn=None, val_dataloader=None,
          loss_schedules=None, video_shape=None, opt=None, log_every=None, process_group=None):
    if loss_schedules is not None or val_dataloader is not None:
        raise NotImplementedError("loss_schedules / val_dataloader are unused by every NeuCam config")
    epochs = opt.num_epochs
    steps_til_summary = getattr(opt, "steps_til_summary", 1000)
    epochs_til_ckpt = getattr(opt, "epochs_til_ckpt", 1000)
    start_freeze = getattr(opt, "start_freeze_tm", -1)
    log_every = steps_til_summary if log_every is None else log_every

    rank = 0
    if process_group is not None:
        import torch.distributed as dist
        rank = dist.get_rank(process_group)
    is_writer = rank == 0
    if is_writer:
        os.makedirs(os.path.join(model_dir, "checkpoints"), exist_ok=True)
    first_in, _ = next(iter(train_dataloader))
    B = first_in["coords"].reshape(-1, 3).shape[0]
    global_pixels = getattr(opt, "global_pixels", None)
    layered = needs_layered_step(models)   # uv-mapping / two-layer configs (me_dynamic, video_deblur, video_hdr)
    if layered:
        step = LayeredTrainStep(models, video_shape, StepOptions.from_opt(opt), process_group=process_group,
                                global_pixels=global_pixels)
    else:
        step = FusedTrainStep(models, video_shape, StepOptions.from_opt(opt), B, process_group=process_group,
                              global_pixels=global_pixels)

    def draw_gamma():
        if not step.opt.use_random_gamma:
            return None
        g = torch.rand(1) * 100 + 1.0                                                      # training.py:181
        if process_group is not None:      # one gamma per step for the whole (sharded) batch, like the reference
            g = g.to(step.device)
            dist.broadcast(g, src=dist.get_global_rank(process_group, 0), group=process_group)
        return g

    use_mask = False
    lr_groups = None
    total_steps = 0
    history = []
    for epoch in range(epochs):
        if epoch and not epoch % epochs_til_ckpt and is_writer:
            nutils.save_model(models, os.path.join(model_dir, "checkpoints", "model_epoch_%04d.pth" % epoch))
        for model_input, gt in train_dataloader:
            gamma = draw_gamma()
            step.graph_step(model_input, gt, gamma, use_mask=use_mask, lr_groups=lr_groups)
            if not total_steps % log_every:
                losses = {k: float(v) for k, v in step.losses().items()}
                history.append((total_steps, losses))
                if not total_steps % steps_til_summary and is_writer:
                    nutils.save_model(models, os.path.join(model_dir, "checkpoints", "model_current.pth"))
                    if summary_fn is not None:
                        summary_fn(model_dir, models, gt, None, None, total_steps, opt)
            if total_steps == start_freeze:
                step.params.reset_optimizer()
                lr_groups = lr_groups_after_freeze()
                use_mask = True
            total_steps += 1
    if is_writer:
        nutils.save_model(models, os.path.join(model_dir, "checkpoints", "model_final.pth"))
    return history
