"""Multi-GPU sampling: one process per GPU (torch.distributed, NCCL), the sample batch sharded over
ranks, no per-step communication, ONE gather of the results at the end.

The reference's only multi-GPU mechanism is ``torch.nn.DataParallel`` around each model
(ccsd/src/utils/loader.py:134-135, 649-650), with the solver state on cuda:0.  Samples are independent
in every network and predictor; the only cross-sample term is the Langevin / S4 step size, a batch MEAN
of per-sample norms (ccsd/src/solver.py:695-699, 763-767, 1300-1311).  Each rank uses the mean over
its own shard -- exactly the semantics of the reference's ``divide_batch`` (independent sub-batches,
ccsd/src/sampler.py:224-232) -- so parity is checked against the oracle run at the shard's batch
size.  Philox noise is keyed by the GLOBAL sample index (``sample_offset``), so with no corrector the
samples do not depend on the number of ranks at all (tests/test_shard_gloo.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(t: torch.Tensor, sizes: Sequence[int], group=None) -> torch.Tensor:
    """All-gather row shards (padded to the largest when their lengths differ) and concatenate.  One collective writing
    straight into the result tensor (``all_gather_into_tensor``); the list API is only the fallback for backends
    without it."""
    world = len(sizes)
    if world == 1:
        return t
    mx = max(sizes)
    pad = t
    if t.shape[0] < mx:
        pad = torch.cat([t, t.new_zeros((mx - t.shape[0],) + tuple(t.shape[1:]))])
    pad = pad.contiguous()
    out = pad.new_empty((world * mx,) + tuple(pad.shape[1:]))
    try:
        dist.all_gather_into_tensor(out, pad, group=group)
    except (RuntimeError, NotImplementedError):
        dist.all_gather(list(out.view((world, mx) + tuple(pad.shape[1:])).unbind(0)), pad, group=group)
    if all(n == mx for n in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + n] for r, n in enumerate(sizes)])


def sharded_sample(
    make_sampler: Callable[[int], Callable], models: Sequence, init_flags: torch.Tensor, *, seed: Optional[int] = None,
    quantize_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, group=None, **kw,
) -> List[torch.Tensor]:
    """Run ``make_sampler(local_batch)(*models, flags_shard, sample_offset=lo, seed=seed)`` on this
    rank's shard of ``init_flags`` (the full batch, identical on every rank) and gather the results.

    make_sampler: e.g. ``lambda b: get_pc_sampler(sde_x, sde_adj, (b, N, F), (b, N, N), ...)``.
    quantize_fn: applied to adj / rank2 before the gather (the reference quantises right after
    sampling, sampler.py:531-543; gathering uint8 instead of fp32 cuts the payload 4x).
    Returns [x, adj(, rank2)] for the FULL batch on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total = init_flags.shape[0]
    lo, hi = shard_bounds(total, world, rank)
    if seed is None:
        s = torch.randint(0, 2 ** 62, (1,))
        if world > 1:
            # the collective's device: CUDA for NCCL (whatever device the flags live on), CPU for gloo
            backend = dist.get_backend(group)
            s = s.to(torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else "cpu")
            src = dist.get_global_rank(group, 0) if group is not None else 0   # `src` is a GLOBAL rank
            dist.broadcast(s, src=src, group=group)
        seed = int(s.item())
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    if hi > lo:
        fn = make_sampler(hi - lo)
        out = fn(*models, init_flags[lo:hi], seed=seed, sample_offset=lo, **kw)
        n_obj = len(out) - 2
        res = list(out[:n_obj])
        if quantize_fn is not None:
            res = [res[0]] + [quantize_fn(t) for t in res[1:]]
    else:
        # more ranks than samples: this rank has nothing to sample but still takes part in the gather.  It needs the
        # per-object shapes / dtypes, which it gets from a one-sample dry construction of the sampler's outputs.
        res = _empty_results(make_sampler, models, init_flags, quantize_fn, **kw)
    return [gather_rows(t, sizes, group) for t in res]


def _empty_results(make_sampler, models, init_flags, quantize_fn, **kw) -> List[torch.Tensor]:
    """Zero-row tensors with the shapes / dtypes / device a non-empty shard would return (one sample, zero steps)."""
    kw = dict(kw)
    kw["max_steps"] = 0
    kw["record_traj"] = False
    out = make_sampler(1)(*models, init_flags[:1], seed=0, sample_offset=0, **kw)
    res = list(out[: len(out) - 2])
    if quantize_fn is not None:
        res = [res[0]] + [quantize_fn(t) for t in res[1:]]
    return [t[:0] for t in res]
