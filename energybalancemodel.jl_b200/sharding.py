"""Member sharding across GPUs (one process per GPU) and the gather of ensemble diagnostics.

Members are independent (no reference code path couples them), so the data path has NO collective: each rank
integrates its own members (contiguous blocks, or 32-member packets dealt round-robin: ``member_deal``).  The only exchange is the final gather of the per-member-year L0
diagnostics to rank 0 -- ``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

__all__ = ["member_block", "member_deal", "gather_member_rows"]


def member_block(nmem_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block (offset, count) of ``rank``; the first ``nmem_total % world`` ranks take one extra member."""
    if not (0 <= rank < world) or nmem_total < 0:
        raise ValueError("need 0 <= rank < world and nmem_total >= 0")
    base, extra = divmod(nmem_total, world)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def member_deal(nmem_total: int, world: int, rank: int, key=None, group: int = 32):
    """Global member indices of ``rank`` when the ensemble is dealt in ``group``-member packets, round-robin over the
    ranks, after a stable sort by ``key`` (e.g. the regime of the initial state: what a member costs).

    A contiguous cut (``member_block``) of an ensemble ordered by a physical parameter hands one rank all the
    expensive members (round 1: the rank holding F = -20... ran 10 % longer than the others and set the step time).
    Dealing packets of 32 -- the granularity at which the classic kernel wants members of one regime side by side --
    gives every rank the same mix.  Ranks differ by at most one packet."""
    if not (0 <= rank < world) or nmem_total < 0 or group < 1:
        raise ValueError("need 0 <= rank < world, nmem_total >= 0, group >= 1")
    order = np.arange(nmem_total, dtype=np.int64) if key is None else np.argsort(np.asarray(key), kind="stable").astype(np.int64)
    ngroups = (nmem_total + group - 1) // group
    mine = [order[g * group:(g + 1) * group] for g in range(rank, ngroups, world)]
    return np.concatenate(mine) if mine else np.empty(0, dtype=np.int64)


def gather_member_rows(local, nmem_total: int, dst: int = 0, group=None, index=None):
    """Gather per-member rows ``local[count_r, ...]`` of every rank into ``[nmem_total, ...]`` on ``dst`` (member
    order = global member index).  Returns the gathered tensor on ``dst`` and ``None`` elsewhere.  Ragged blocks
    are padded to the largest block for the collective and trimmed afterwards.  ``index``: list (one entry per
    rank) of the global member indices each rank holds (``member_deal``); default: contiguous ``member_block``s."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != nmem_total:
            raise ValueError("single process: local block must hold every member")
        if index is not None:
            out = torch.empty_like(local)
            out[torch.as_tensor(np.asarray(index[0]), device=local.device, dtype=torch.long)] = local
            return out
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [member_block(nmem_total, world, r)[1] for r in range(world)] if index is None else [len(ix) for ix in index]
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} members, expected {counts[rank]}")
    cmax = max(counts)
    send = local
    if counts[rank] < cmax:
        pad = torch.zeros((cmax - counts[rank],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send = torch.cat([local, pad], dim=0)
    send = send.contiguous()
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    if index is None:
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    out = torch.empty((nmem_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for b, c, ix in zip(bufs, counts, index):
        out[torch.as_tensor(np.asarray(ix), device=local.device, dtype=torch.long)] = b[:c]
    return out
