"""Switch used by the training steps that own a flat gradient buffer (FlatParameters): while it is on, the fused
autograd Functions let their weight-gradient kernels accumulate straight into `param.grad` (the kernels accumulate
anyway) and return None for those inputs, which saves a zero-filled temporary and an AccumulateGrad add per
parameter per call.  Off (the default, e.g. modules used under a stock torch optimizer) they return gradients the
ordinary way."""
import contextlib

_direct = False
_step_cache = None     # dict that lives for ONE forward + backward of a step (parameters are constant inside it), or None


def direct():
    return _direct


@contextlib.contextmanager
def accumulate_into_param_grad():
    global _direct
    prev, _direct = _direct, True
    try:
        yield
    finally:
        _direct = prev


def targets_of(params):
    """The `.grad` tensors of `params` when direct accumulation is on and every one is a preallocated, contiguous
    fp32 buffer of the parameter's shape; else None."""
    if not _direct:
        return None
    out = []
    for p in params:
        g = getattr(p, "grad", None)
        if g is None or g.shape != p.shape or not g.is_contiguous() or g.dtype != p.dtype or not p.requires_grad:
            return None
        out.append(g)
    return tuple(out)


@contextlib.contextmanager
def one_step():
    """Scope of one forward + backward of a training step.  Parameters do not change inside it, so operand copies derived
    from them (the hi / lo fp16 weight stages of a ReLU MLP that the step evaluates twice: on the taps and on the
    rigidity loss's shifted pixels) are packed once and shared through `step_cache()`.  The step's optimizer kernel writes
    the parameters without bumping torch's version counters, which is why the cache must not outlive the scope."""
    global _step_cache
    prev, _step_cache = _step_cache, {}
    try:
        yield
    finally:
        _step_cache = prev


def step_cache():
    return _step_cache
