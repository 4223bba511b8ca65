"""Member sharding across GPUs (one process per GPU) and the gather of ensemble diagnostics.

Members are independent (no reference code path couples them), so the data path has NO collective: each rank
integrates a contiguous block of members.  The only exchange is the final gather of the per-member-year L0
diagnostics to rank 0 -- ``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

__all__ = ["member_block", "gather_member_rows"]


def member_block(nmem_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block (offset, count) of ``rank``; the first ``nmem_total % world`` ranks take one extra member."""
    if not (0 <= rank < world) or nmem_total < 0:
        raise ValueError("need 0 <= rank < world and nmem_total >= 0")
    base, extra = divmod(nmem_total, world)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def gather_member_rows(local, nmem_total: int, dst: int = 0, group=None):
    """Gather per-member rows ``local[count_r, ...]`` of every rank into ``[nmem_total, ...]`` on ``dst`` (member
    order = global member index).  Returns the gathered tensor on ``dst`` and ``None`` elsewhere.  Ragged blocks
    are padded to the largest block for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != nmem_total:
            raise ValueError("single process: local block must hold every member")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [member_block(nmem_total, world, r)[1] for r in range(world)]
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
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
