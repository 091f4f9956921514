"""Instance sharding of posterior sampling and log-prob evaluation over the GPUs of one node.

Rows of the flow are independent and the per-instance work (feature network, condition projection)
is local, so the path shards by conditioning instance with NO collective on the data path
(SURVEY.md section 8e): rank r owns a contiguous block of instances and produces every sample for them;
weights are replicated.  An optional all_gather assembles the blocks.  One process per GPU
(torchrun); the functions below only need an initialised ``torch.distributed`` group (NCCL on the
GPUs, gloo in the CPU tests of the host-side logic).
"""
from __future__ import annotations

from typing import Any, Callable, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_conditions", "gather_instance_blocks", "sample_sharded", "log_prob_sharded"]


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the instances owned by ``rank``: contiguous, sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"invalid rank {rank} for world size {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _group_info(group: Any = None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_conditions(conditions: Sequence[torch.Tensor], group: Any = None) -> tuple[list[torch.Tensor], int, int]:
    """This rank's block of every condition tensor, plus its [lo, hi) bounds."""
    n = conditions[0].shape[0]
    if any(c.shape[0] != n for c in conditions):
        raise ValueError(f"All conditions must have the same number of samples (dim = 0). Got {[c.shape for c in conditions]}.")
    rank, world = _group_info(group)
    lo, hi = shard_bounds(n, rank, world)
    return [c[lo:hi] for c in conditions], lo, hi


def _collective_device(block: torch.Tensor, group: Any = None) -> torch.device:
    """Device the backend of ``group`` moves tensors on: NCCL needs CUDA tensors, gloo takes CPU ones."""
    backend = str(dist.get_backend(group)).lower()
    if "nccl" in backend and block.device.type != "cuda":
        return torch.device("cuda", torch.cuda.current_device())
    return block.device


def gather_instance_blocks(block: torch.Tensor, n_total: int, dim: int, group: Any = None) -> torch.Tensor:
    """all_gather blocks of unequal length along ``dim`` (instance axis) into the full tensor.

    A CPU block under the NCCL backend is moved to this rank's GPU for the collective and the result returned on the
    block's own device (prefer ``output_device=<the model's device>`` in that case: ``sample_sharded`` does)."""
    rank, world = _group_info(group)
    if world == 1:
        return block
    home = block.device
    cdev = _collective_device(block, group)
    if cdev != home:
        return gather_instance_blocks(block.to(cdev), n_total, dim, group).to(home)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    moved = block.movedim(dim, 0).contiguous()
    pad = torch.zeros((longest,) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
    pad[: moved.shape[0]] = moved
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    return full.movedim(0, dim)


def sample_sharded(sample_fn: Callable[..., torch.Tensor], n_samples: int, *conditions: torch.Tensor,
                   gather: bool = False, group: Any = None, **kwargs: Any) -> torch.Tensor:
    """``sample_fn(n_samples, *conditions_block, outer=True, **kwargs)`` on this rank's instances.

    ``sample_fn`` is ``model.sample``.  Returns (n_samples, n_local, D), or the full
    (n_samples, N, D) on every rank when ``gather=True``.
    """
    mine, lo, hi = shard_conditions(conditions, group)
    want = kwargs.get("output_device", None)
    if gather and want is None and dist.is_available() and dist.is_initialized() \
            and "nccl" in str(dist.get_backend(group)).lower():
        # model.sample defaults to output_device="cpu": gathering that over NCCL would round-trip the samples through
        # host memory (and all_gather of a CPU tensor raises).  Keep the block on the GPU for the collective.
        kwargs["output_device"] = torch.device("cuda", torch.cuda.current_device())
    out = sample_fn(n_samples, *mine, outer=True, **kwargs)
    if gather:
        out = gather_instance_blocks(out, conditions[0].shape[0], dim=1, group=group)
    return out


def log_prob_sharded(log_prob_fn: Callable[..., torch.Tensor], y: torch.Tensor, *conditions: torch.Tensor,
                     gather: bool = False, group: Any = None, **kwargs: Any) -> torch.Tensor:
    """``log_prob_fn(y_block, *conditions_block)`` on this rank's rows; optional gather to (N,)."""
    mine, lo, hi = shard_conditions(conditions, group)
    out = log_prob_fn(y[lo:hi], *mine, **kwargs)
    if gather:
        out = gather_instance_blocks(out, conditions[0].shape[0], dim=0, group=group)
    return out
