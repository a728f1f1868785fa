"""Shuffle-BN as a permutation exchange (replaces the reference's all_gather + index,
moco/builder_diffspeed_diffloss.py:361-406).

The reference gathers every rank's key clips on every rank (W x the batch) and then keeps B rows.  Here each rank
receives only the B rows it will encode: the rows a rank owes to each peer are packed by a gather kernel, moved with
one ``all_to_all_single`` over NCCL/NVLink, and put in slice order by a second gather.  The permutation itself is
still drawn exactly as the reference does (``torch.randperm(B*W)`` on every rank's CPU generator, rank 0's wins).

``plan_exchange`` is pure index arithmetic on host tensors so that it can be tested without GPUs.

``ShuffleExchange`` is the transport the builder uses.  On NCCL ranks of one node it keeps every rank's key clips in
buffers that all peers have mapped over NVLink (CUDA IPC, csrc/peer.cu) and each rank PULLS the rows the permutation
assigns to it with one kernel: no pack / all_to_all / unpack passes, no host-side split sizes, and the permutation only
has to exist on the device — it travels in one tiny all-reduce per step that doubles as the "key clips are written"
barrier, so no rank's host ever blocks on another rank's host.  Without peer mapping (gloo, IPC refused) the
``all_to_all_single`` path below is used.
"""
import ctypes as C
import logging
import os
from typing import List, NamedTuple, Optional

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)


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


# ----------------------------------------------------------------------------------------------------------------
# peer-memory transport
# ----------------------------------------------------------------------------------------------------------------
class _RawDeviceBuffer:
    """A device allocation exposed through ``__cuda_array_interface__`` (torch.as_tensor keeps this object alive)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


class _PinnedRing:
    """Pinned host staging for the per-step permutations.  The host runs ahead of the device, so a buffer is only
    rewritten after the copy that read it has completed (an event per slot; practically never waits)."""

    def __init__(self, shape, slots: int = 8):
        self.bufs = [torch.empty(shape, dtype=torch.int64).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0

    def next(self):
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.events[i] is not None:
            self.events[i].synchronize()
        return i, self.bufs[i]

    def mark(self, i):
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class ShuffleExchange:
    """Key-clip buffers of one model + the shuffle-BN row exchange (builder:361-387) for ``rows`` clips per rank.

    Per step:  ``k_neg_buf, k_buf = begin_step()`` (the producer writes the two key-clip batches there),
    ``idx = draw(n)`` (reference-order randperm draws, rank 0's values), ``publish(idx)`` on the stream that wrote the
    buffers, then ``pull(slot, idx[slot])`` -> this rank's shuffled batch."""
    SLOT_KNEG, SLOT_K = 0, 1

    def __init__(self, rows: int, row_shape, dtype: torch.dtype, device: torch.device, force_mode: Optional[str] = None):
        self.rank, self.world = world_info()
        self.rows, self.row_shape, self.dtype, self.device = rows, tuple(row_shape), dtype, device
        row_elems = 1
        for v in self.row_shape:
            row_elems *= v
        self.slot_bytes = rows * row_elems * torch.empty((), dtype=dtype).element_size()
        self.slot_bytes = (self.slot_bytes + 255) // 256 * 256
        self.sets = 2 if self.world > 1 else 1     # peers may still read step n while step n+1 is being written
        self._set = 0
        self._ring = None
        self._peer_ptr = None
        self._mapped = []
        mode = force_mode or os.environ.get("RSP_SHUFFLE_EXCHANGE")
        if self.world == 1:
            mode = "local"
        elif mode is None:
            mode = "peer" if (device.type == "cuda" and dist.get_backend() == "nccl") else "a2a"
        if mode == "peer" and not self._setup_peer():
            mode = "a2a"
        self.mode = mode
        if mode != "peer":
            flat = torch.empty((self.sets * 2 * self.slot_bytes,), dtype=torch.uint8, device=device)
            self._carve(flat)

    # ---- buffers ------------------------------------------------------------------------------------------------
    def _carve(self, flat_u8: torch.Tensor):
        self._flat = flat_u8
        n_elem = self.rows
        for v in self.row_shape:
            n_elem *= v
        self._bufs = []
        for s in range(self.sets):
            slots = []
            for k in range(2):
                off = (s * 2 + k) * self.slot_bytes
                t = flat_u8[off:off + self.slot_bytes].view(self.dtype)[:n_elem].view((self.rows,) + self.row_shape)
                slots.append(t)
            self._bufs.append(slots)

    def _setup_peer(self) -> bool:
        """cudaMalloc + IPC export / open on every rank; all ranks fall back together when any of them fails."""
        from .. import _lib
        total = self.sets * 2 * self.slot_bytes
        ok, err = True, ""
        try:
            _lib._ensure_device()
            p = C.c_void_p()
            _lib.call("rsp_peer_alloc", total, C.byref(p))
            self._peer_ptr = p.value
            handle = C.create_string_buffer(64)
            _lib.call("rsp_peer_export", p, handle)
            mine = handle.raw
        except RuntimeError as e:   # keep the collective below symmetric
            ok, err, mine = False, str(e), b""
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        bases = []
        if ok and all(len(h) == 64 for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        bases.append(self._peer_ptr)
                    else:
                        m = C.c_void_p()
                        _lib.call("rsp_peer_open", h, C.byref(m))
                        self._mapped.append(m.value)
                        bases.append(m.value)
            except RuntimeError as e:
                ok, err = False, str(e)
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) != 1:
            if self.rank == 0:
                logger.warning("rspnet_b200: peer-memory shuffle exchange unavailable (%s); using all_to_all", err)
            self.close()
            return False
        self._carve(torch.as_tensor(_RawDeviceBuffer(self._peer_ptr, total), device=self.device))
        # device tables of the W base pointers of every (set, slot)
        self._tables = [[torch.tensor([b + (s * 2 + k) * self.slot_bytes for b in bases], dtype=torch.int64,
                                      device=self.device) for k in range(2)] for s in range(self.sets)]
        return True

    def close(self):
        from .. import _lib
        for m in self._mapped:
            try:
                _lib.call("rsp_peer_close", C.c_void_p(m))
            except RuntimeError:
                pass
        self._mapped = []
        if self._peer_ptr is not None:
            try:
                _lib.call("rsp_peer_free", C.c_void_p(self._peer_ptr))
            except RuntimeError:
                pass
            self._peer_ptr = None

    def begin_step(self):
        """Buffers (k_neg, k) this step's key clips go to."""
        self._set = (self._set + 1) % self.sets
        return self._bufs[self._set][0], self._bufs[self._set][1]

    def holds(self, x: torch.Tensor, slot: int) -> bool:
        b = self._bufs[self._set][slot]
        return x.data_ptr() == b.data_ptr() and tuple(x.shape) == tuple(b.shape) and x.dtype == b.dtype

    # ---- permutations -------------------------------------------------------------------------------------------
    def draw(self, n_all: int, count: int = 2, out: Optional[torch.Tensor] = None):
        """``count`` x the reference's ``torch.randperm(batch_size_all)`` (builder:375) on this rank's CPU generator, in
        order.  Returns (device int64 [count, n_all], host copy or None).  Every rank consumes its generator like the
        reference; ranks other than 0 contribute zeros so that the sum all-reduce in publish() yields rank 0's draw.
        ``out``: a persistent device tensor to fill instead of a fresh one (CUDA-graph replays read it)."""
        perms = [torch.randperm(n_all) for _ in range(count)]
        if self.mode == "a2a":
            host = torch.stack(perms)
            broadcast_permutation(host)                    # host-side plan needs the values: blocking side-group bcast
            return host.to(self.device, non_blocking=True) if self.device.type == "cuda" else host, host
        if self.device.type != "cuda":
            return torch.stack(perms), None
        if self._ring is None or tuple(self._ring.bufs[0].shape) != (count, n_all):
            self._ring = _PinnedRing((count, n_all))
        i, host = self._ring.next()
        if self.rank == 0:
            for j, p in enumerate(perms):
                host[j].copy_(p)
        else:
            host.zero_()
        dev = out if out is not None else torch.empty((count, n_all), dtype=torch.int64, device=self.device)
        dev.copy_(host, non_blocking=True)
        self._ring.mark(i)
        return dev, None

    def publish(self, idx_dev: torch.Tensor):
        """Call on the stream that wrote this step's key clips, after them: one small sum all-reduce makes rank 0's
        permutations visible everywhere AND orders every peer's pull after every rank's writes."""
        if self.mode == "peer":
            dist.all_reduce(idx_dev)

    # ---- rows -----------------------------------------------------------------------------------------------------
    def pull(self, slot: int, idx_all: torch.Tensor, idx_host: Optional[torch.Tensor], gather_rows):
        """This rank's shuffled batch ``concat_all_gather(x)[idx_all.view(W, -1)[rank]]`` of the clips in ``slot``."""
        mine = idx_all.view(self.world, -1)[self.rank]
        src = self._bufs[self._set][slot]
        if self.mode == "local":
            return gather_rows(src, mine)
        if self.mode == "peer":
            from .. import ops
            return ops.gather_rows_peer(self._tables[self._set][slot], mine, self.rows, self.row_shape, self.dtype)
        return exchange_rows(src, idx_host, gather_rows)
