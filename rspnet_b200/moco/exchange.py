"""Shuffle-BN as a permutation exchange (replaces the reference's all_gather + index,
moco/builder_diffspeed_diffloss.py:361-406).

The reference gathers every rank's key clips on every rank (W x the batch) and then keeps B rows.  Here each rank
receives only the B rows it will encode: the rows a rank owes to each peer are packed by a gather kernel, moved with
one ``all_to_all_single`` over NCCL/NVLink, and put in slice order by a second gather.  The permutation itself is
still drawn exactly as the reference does (``torch.randperm(B*W)`` on every rank's CPU generator, rank 0's wins).

``plan_exchange`` is pure index arithmetic on host tensors so that it can be tested without GPUs.
"""
from typing import List, NamedTuple

import torch
import torch.distributed as dist


class ExchangePlan(NamedTuple):
    send_index: torch.Tensor     # int64 [n_send]  local row ids, grouped by destination rank
    send_counts: List[int]       # rows sent to each rank
    recv_counts: List[int]       # rows received from each rank
    unpack_index: torch.Tensor   # int64 [B]: shuffled[p] = received[unpack_index[p]]


def plan_exchange(idx_shuffle: torch.Tensor, rank: int, world: int) -> ExchangePlan:
    """idx_shuffle: int64 [B*W] on the host. Rank r must end up with rows ``idx_shuffle.view(W,-1)[r]`` (builder:383-387)."""
    idx = idx_shuffle.view(world, -1)
    batch = idx.shape[1]
    owner = idx // batch                                    # rank that holds each requested row
    send_parts, send_counts = [], []
    for dst in range(world):
        mine = owner[dst] == rank                           # positions of dst's slice that I own, ascending
        send_parts.append(idx[dst][mine] - rank * batch)
        send_counts.append(int(mine.sum()))
    need_owner = owner[rank]
    recv_counts = [int((need_owner == src).sum()) for src in range(world)]
    # data arrives grouped by source rank, each group in ascending slice position
    arrival_pos = torch.cat([torch.nonzero(need_owner == src).flatten() for src in range(world)])
    unpack_index = torch.argsort(arrival_pos)
    return ExchangePlan(torch.cat(send_parts), send_counts, recv_counts, unpack_index)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


_cpu_group = None


def _host_group():
    """A gloo side group for tiny host-side broadcasts (keeps the permutation off the GPU stream: no device sync)."""
    global _cpu_group
    if _cpu_group is None:
        _cpu_group = dist.new_group(backend="gloo")
    return _cpu_group


def broadcast_permutation(idx_cpu: torch.Tensor) -> torch.Tensor:
    """Rank 0's permutation on every rank's host (the reference broadcasts the CUDA copy, builder:375-378)."""
    _, world = world_info()
    if world == 1:
        return idx_cpu
    if dist.get_backend() == "gloo":
        dist.broadcast(idx_cpu, src=0)
    else:
        dist.broadcast(idx_cpu, src=0, group=_host_group())
    return idx_cpu


def exchange_rows(x: torch.Tensor, idx_shuffle_cpu: torch.Tensor, gather_rows) -> torch.Tensor:
    """Returns ``concat_all_gather(x)[idx_shuffle.view(W,-1)[rank]]`` without materialising the gather.

    ``gather_rows(src, index)`` is the device row-gather (rspnet_b200.ops.gather_rows on CUDA)."""
    rank, world = world_info()
    if world == 1:
        return gather_rows(x, idx_shuffle_cpu.to(x.device, non_blocking=True))
    plan = plan_exchange(idx_shuffle_cpu, rank, world)
    packed = gather_rows(x, plan.send_index.to(x.device, non_blocking=True))
    recv = torch.empty((sum(plan.recv_counts),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv, packed, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts)
    return gather_rows(recv, plan.unpack_index.to(x.device, non_blocking=True))


def all_gather_rows(x: torch.Tensor) -> torch.Tensor:
    """concat_all_gather without the list + cat copies (builder:249-260)."""
    _, world = world_info()
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    return out
