"""Data-parallel plumbing for the hot path: batch sharding, one contiguous gradient bucket,
one NCCL all-reduce after backward, SGD with grad_scale = 1/world.  `torch.distributed` is plumbing only.

The layout logic is pure Python (tested on CPU with gloo, world size 2, tests/test_dp_cpu.py); the device part
wraps the bucket's HBM as a torch tensor and calls `all_reduce` on the same stream the kernels run on."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of a batch of n owned by `rank` (replica r gets rows [r*n/R, (r+1)*n/R))."""
    if n % world:
        raise ValueError(f"batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


class BucketLayout:
    """Offsets of every parameter's gradient inside one flat buffer (element units)."""

    def __init__(self, shapes: Sequence[Sequence[int]]):
        self.shapes = [tuple(int(d) for d in s) for s in shapes]
        self.offsets: List[int] = []
        off = 0
        for s in self.shapes:
            self.offsets.append(off)
            off += int(np.prod(s)) if len(s) else 1
        self.total = off

    def views(self, flat):
        """Slices of a flat numpy array / torch tensor shaped like the parameters."""
        out = []
        for s, o in zip(self.shapes, self.offsets):
            n = int(np.prod(s)) if len(s) else 1
            out.append(flat[o:o + n].reshape(s))
        return out


class GradientBucket:
    """One contiguous device buffer holding every leaf gradient; `views[i]` is handed to
    `Var.requires_grad(grad_array=...)` so backward writes straight into the bucket."""

    def __init__(self, device, shapes, dtype, ptr=None):
        from .device import CuArray
        self.layout = BucketLayout(shapes)
        self.array = CuArray(device, (self.layout.total,), dtype, ptr=ptr)   # ptr: caller-owned (peer-mapped) memory
        self.views = [self.array.slice_flat(o, s) for s, o in zip(self.layout.shapes, self.layout.offsets)]
        self._torch = None

    def as_torch(self):
        import torch
        if self._torch is None:
            class _CAI:
                pass
            holder = _CAI()
            holder.__cuda_array_interface__ = self.array.cuda_array_interface()
            t = torch.as_tensor(holder, device=f"cuda:{self.array.device.index}")
            self._torch = t.view(torch.bfloat16) if self.array.dtype == 1 else t
        return self._torch

    def all_reduce(self, stream=None):
        """Sum the bucket over all ranks (NCCL), ordered on `stream` with the kernels that produced it."""
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        t = self.as_torch()
        if stream is not None:
            with torch.cuda.stream(stream):
                dist.all_reduce(t)
        else:
            dist.all_reduce(t)


class ReadyRanges:
    """Bookkeeping of which element ranges of a gradient bucket are final during one backward pass.

    `add(lo, hi)` returns the ranges to exchange NOW: a range of at least `min_elems` goes out at once, widened by
    any adjacent small pieces that were waiting (a bias next to the last block of its weight); small pieces wait for
    such a neighbour.  `flush()` returns what is still waiting, merged where contiguous."""

    def __init__(self, min_elems: int):
        self.min_elems = int(min_elems)
        self.held: List[Tuple[int, int]] = []

    def add(self, lo: int, hi: int) -> List[Tuple[int, int]]:
        if hi - lo < self.min_elems:
            self.held.append((lo, hi))
            return []
        grown = True
        while grown:
            grown = False
            for p in self.held:
                if p[0] == hi:
                    hi = p[1]
                elif p[1] == lo:
                    lo = p[0]
                else:
                    continue
                self.held.remove(p)
                grown = True
                break
        return [(lo, hi)]

    def flush(self) -> List[Tuple[int, int]]:
        out: List[Tuple[int, int]] = []
        for lo, hi in sorted(self.held):
            if out and out[-1][1] == lo:
                out[-1] = (out[-1][0], hi)
            else:
                out.append((lo, hi))
        self.held = []
        return out


class OverlappedAllReduce:
    """All-reduce leaf gradients (slices of a GradientBucket) on a side stream as soon as backward has produced them,
    so the exchange overlaps the remaining backward kernels.  Large matrices are delivered by the graph in row blocks
    (`row_chunks`), so even a single layer's exchange overlaps its own dW GEMMs.

        sync = OverlappedAllReduce(bucket, compute_stream, params)
        loss.backward(1.0)        # hooks fire inside; the exchanges are in flight when this returns
        sync.wait()               # flush what is left; compute stream waits for the exchanges; then optimizer.step()
    """

    def __init__(self, bucket: GradientBucket, compute_stream, params, chunk_elems: int = 4 << 20,
                 max_chunks: int = 8):
        import torch
        self.torch = torch
        self.bucket, self.compute, self.params = bucket, compute_stream, params
        self.comm = torch.cuda.Stream(device=bucket.array.device.index)
        self._events = [torch.cuda.Event() for _ in range(4 * len(params) + 8)]
        self._ev_i = 0
        self._pending = []
        self.flat = bucket.as_torch()
        self.ranges = ReadyRanges(min_elems=chunk_elems // 4)
        self.launched = 0
        lay = bucket.layout
        for pi, p in enumerate(params):
            n = int(np.prod(lay.shapes[pi])) if lay.shapes[pi] else 1
            chunks = max(1, min(max_chunks, n // chunk_elems))
            p.set_grad_hook(lambda b, e, off=lay.offsets[pi]: self._ready(off + b, off + e), row_chunks=chunks)

    # The hook runs INSIDE backward, between two kernel launches of the compute stream: anything slow here (an NCCL
    # enqueue costs tens of microseconds of host time) delays the next backward kernel.  So the hook only records an
    # event; the exchange itself is enqueued at the next hook or in wait(), by which time later kernels are queued.
    def _flush_pending(self) -> None:
        import torch.distributed as dist
        if not self._pending:
            return
        pending, self._pending = self._pending, []
        for ev, pieces in pending:
            self.comm.wait_event(ev)
            with self.torch.cuda.stream(self.comm):
                for lo, hi in pieces:
                    dist.all_reduce(self.flat[lo:hi])
                    self.launched += 1

    def _exchange(self, pieces) -> None:
        self._flush_pending()
        if not pieces:
            return
        ev = self._events[self._ev_i % len(self._events)]
        self._ev_i += 1
        ev.record(self.compute)           # everything launched so far (incl. the kernels that wrote the grads)
        self._pending.append((ev, pieces))

    def _ready(self, lo: int, hi: int) -> None:
        self._exchange(self.ranges.add(lo, hi))

    def wait(self) -> None:
        self._exchange(self.ranges.flush())
        self._flush_pending()
        self.compute.wait_stream(self.comm)


class PeerMemory:
    """Device buffers every rank can address (nk_ipc_alloc + CUDA IPC handles exchanged through the process group):
    `ptrs[r]` is rank r's buffer as seen from THIS process (ptrs[rank] is the local allocation)."""

    def __init__(self, device, nbytes: int, world: int, rank: int):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib as L
        self.device, self.world, self.rank, self.nbytes = device, world, rank, int(nbytes)
        # every step below is followed by a collective agreement, so that a failure on one rank (no peer access, IPC
        # disabled in the container, out of memory) raises on ALL ranks instead of leaving the others in a collective
        self.local, handle, err = 0, None, None
        try:
            p = C.c_void_p()
            L.check(L.lib.nk_ipc_alloc(device.ctx, self.nbytes, C.byref(p)), device.ctx)
            self.local = int(p.value)
            h = C.create_string_buffer(64)
            L.check(L.lib.nk_ipc_export(device.ctx, C.c_void_p(self.local), h), device.ctx)
            handle = bytes(h.raw)
        except Exception as e:      # noqa: BLE001 -- reported below, on every rank
            err = repr(e)
        handles = [None] * world
        dist.all_gather_object(handles, (handle, err))
        bad = [(r, e) for r, (hd, e) in enumerate(handles) if hd is None]
        if bad:
            self._release()
            raise RuntimeError(f"peer memory: allocation / export failed on rank(s) {bad}")
        self.ptrs: List[int] = []
        self._opened: List[int] = []
        err = None
        try:
            for r in range(world):
                if r == rank:
                    self.ptrs.append(self.local)
                else:
                    q = C.c_void_p()
                    L.check(L.lib.nk_ipc_open(device.ctx, C.create_string_buffer(handles[r][0], 64), C.byref(q)),
                            device.ctx)
                    self.ptrs.append(int(q.value))
                    self._opened.append(int(q.value))
        except Exception as e:      # noqa: BLE001
            err = repr(e)
        errs = [None] * world
        dist.all_gather_object(errs, err)
        bad = [(r, e) for r, e in enumerate(errs) if e is not None]
        if bad:
            self._release()
            raise RuntimeError(f"peer memory: opening the peers' handles failed on rank(s) {bad}")
        self.table = (C.c_void_p * world)(*self.ptrs)

    def _release(self) -> None:
        import ctypes as C
        from . import _lib as L
        for q in getattr(self, "_opened", []):
            L.lib.nk_ipc_close(self.device.ctx, C.c_void_p(q))
        self._opened = []
        if self.local:
            L.lib.nk_ipc_free(self.device.ctx, C.c_void_p(self.local))
            self.local = 0

    def offset_table(self, byte_offset: int):
        import ctypes as C
        return (C.c_void_p * self.world)(*[p + int(byte_offset) for p in self.ptrs])


def rs_eligible(shape, world: int, min_elems: int = 1 << 20) -> bool:
    """A parameter takes the fused exchange when it is a matrix whose rows split into world shards of 128-row GEMM
    tiles (nk_gemm_rs) and it is big enough for the exchange to matter; the rest goes through the all-reduce."""
    if len(shape) != 2 or world < 2:
        return False
    rows, cols = int(shape[0]), int(shape[1])
    # nk_gemm_rs runs the 128x256 tcgen05 tile only (cols > 128) on TMA-addressable operands (cols % 8 == 0)
    return (rows % (world * 128) == 0 and rows * cols >= min_elems and (rows // world * cols) % 4 == 0
            and cols > 128 and cols % 8 == 0)


class FusedGradientExchange:
    """Data-parallel gradient exchange over NVLink peer memory, no library collective on the path:

      * large matrices: the reduce-scatter is fused into the dW GEMM epilogue (nk_gemm_rs), then ONE kernel per matrix on
        a side stream does barrier -> owner reduce -> broadcast into every replica's bucket -> barrier
        (nk_reduce_exchange, device-resident epoch);
      * everything else (biases, small or oddly shaped matrices): one single-CTA kernel per contiguous bucket range
        writes the local values into every peer's receive slots and sums the `world` copies in rank order
        (nk_peer_allreduce_small); ranges above 2^20 elements go through the context-owned NCCL communicator
        (nk_allreduce_sum).

    Every launch is identical from step to step, so the whole step -- exchange included -- can be captured in a CUDA
    graph (Device.capture).  Every replica receives bit-identical gradients.

        ex = FusedGradientExchange(device, compute_stream, shapes, world, rank)     # owns the gradient bucket
        params = [from_ndarray(...).requires_grad(F32, ex.bucket.views[i]) for i ...]
        ex.attach(params)
        loss.backward(1.0)      # dW epilogues push shards to their owners; the exchange kernels run on a side stream
        ex.wait()               # all gradients (summed over ranks) are in the bucket; then optimizer.step()
    """

    SMALL_MAX = 1 << 20

    def __init__(self, device, compute_stream, shapes, world: int, rank: int, reduce_ctas: int = 0,
                 min_elems: int = 1 << 20):
        import torch
        from .device import F32, CuArray
        self.torch, self.device, self.compute = torch, device, compute_stream
        self.world, self.rank, self.reduce_ctas = world, rank, reduce_ctas
        self.layout = BucketLayout(shapes)
        self.bucket_mem = PeerMemory(device, self.layout.total * 4, world, rank)
        self.bucket = GradientBucket(device, shapes, F32, ptr=self.bucket_mem.local)
        self.fused = [i for i, s in enumerate(self.layout.shapes) if rs_eligible(s, world, min_elems)]
        self.slot_off = {}
        off = 0
        for i in self.fused:
            self.slot_off[i] = off
            off += int(np.prod(self.layout.shapes[i]))
        self.slots = PeerMemory(device, max(off, 4) * 4, world, rank)
        # receive slots of the small all-reduce: world copies of every non-fused element
        small_total = self.layout.total - off
        # every non-fused parameter has its own receive region (offsets in floats, 16-byte aligned), so that its all-reduce
        # can be issued the moment its gradient is final instead of in one bulk call at the end of backward
        self.small_off = {}
        cur = 0
        for i, sh in enumerate(self.layout.shapes):
            if i not in self.fused:
                self.small_off[i] = cur
                cur += (int(np.prod(sh)) + 3) // 4 * 4
        small_total = max(small_total, cur)
        self.small_slots = PeerMemory(device, max(small_total, 4) * 4 * world, world, rank)
        self.flags = PeerMemory(device, 1024, world, rank)          # words [0,16): exchange, [64,80): small all-reduce
        self.state = CuArray(device, (16,), F32)                    # zero-filled local words: ExState x 2
        self.comm = torch.cuda.Stream(device=device.index)
        self._events = [torch.cuda.Event() for _ in range(4 * len(shapes) + 8)]
        self._ev_i = 0
        self._pending = []
        self.ranges = ReadyRanges(min_elems=1 << 62)      # everything that is not fused is exchanged in bulk
        self.flat = self.bucket.as_torch()
        self.pushed = 0
        self.comm_device = None
        self._params = []
        self._nccl_ready = False
        import os
        self._debug_skip_reduce = bool(os.environ.get("NK_DP_DEBUG_SKIP_REDUCE"))

    def attach(self, params) -> None:
        from .device import Device
        # the side stream gets its own context handle bound to the same device: kernels are enqueued on self.comm
        if self.comm_device is None:
            self.comm_device = Device(self.device.index, stream=self.comm.cuda_stream)
        # the exchange kernels run beside the dX GEMMs: keep the GEMM's balanced grid (20 SMs free), never the tail split
        self.device.gemm_tail_split(False)
        self._params = list(params)
        lay = self.layout
        for pi, p in enumerate(params):
            if pi in self.fused:
                table = self.slots.offset_table(self.slot_off[pi] * 4)
                p.set_grad_rs(self.world, self.rank, [int(v) for v in table],
                              lambda pushed, pi=pi: self._pushed(pi, pushed))
            else:
                p.set_grad_hook(lambda b, e, pi=pi: self._small_ready(pi, b, e))

    def _small_ready(self, pi: int, b: int, e: int) -> None:
        """gradient-ready hook of a non-fused parameter (called inside backward between two launches): record an event;
        the all-reduce itself is enqueued at the next hook / wait(), on the side stream, where it overlaps the rest of
        backward instead of queueing behind the big matrices' exchanges at the end"""
        self._flush_pending()
        lay = self.layout
        n = int(np.prod(lay.shapes[pi]))
        if (b, e) != (0, n) or n > self.SMALL_MAX:          # partial delivery or too big for the peer-memory kernel
            self.ranges.add(lay.offsets[pi] + b, lay.offsets[pi] + e)
            return
        ev = self._events[self._ev_i % len(self._events)]
        self._ev_i += 1
        ev.record(self.compute)
        self._pending.append((ev, pi))

    def detach(self) -> None:
        """Remove the exchange plans and hooks from the parameters (gradients stay local afterwards)."""
        for pi, p in enumerate(self._params):
            if pi in self.fused:
                p.set_grad_rs(0, 0, None, None)
            else:
                p.set_grad_hook(None)

    def _pushed(self, pi: int, pushed: int) -> None:
        # called inside backward between two kernel launches: record an event only, enqueue the exchange later
        # (next hook or wait()) so that the host time of the enqueue does not delay the next backward kernel
        self._flush_pending()
        lay = self.layout
        n = int(np.prod(lay.shapes[pi]))
        if not pushed:                                  # computed locally: exchanged like the small tensors
            self.ranges.held.append((lay.offsets[pi], lay.offsets[pi] + n))
            return
        self.pushed += 1
        if self._debug_skip_reduce:      # development knob: time the pushing GEMM alone (results are then wrong)
            return
        ev = self._events[self._ev_i % len(self._events)]
        self._ev_i += 1
        ev.record(self.compute)                         # the dW GEMM (and its pushes) launched so far
        self._pending.append((ev, pi))

    def _flush_pending(self) -> None:
        from . import _lib as L
        import ctypes as C
        if not self._pending:
            return
        pending, self._pending = self._pending, []
        lay = self.layout
        for ev, pi in pending:
            self.comm.wait_event(ev)
            if pi not in self.fused:                        # small tensor: single-CTA all-reduce through peer memory
                n = int(np.prod(lay.shapes[pi]))
                L.check(L.lib.nk_peer_allreduce_small(self.comm_device.ctx, C.c_void_p(self.bucket_mem.local + lay.offsets[pi] * 4),
                                                      self.small_slots.offset_table(self.small_off[pi] * self.world * 4),
                                                      self.flags.offset_table(256), self.world, self.rank, n,
                                                      C.c_void_p(self.state.ptr.value + 16)), self.comm_device.ctx)
                continue
            shard = int(np.prod(lay.shapes[pi])) // self.world
            grads = self.bucket_mem.offset_table(lay.offsets[pi] * 4)
            slots_local = self.slots.local + self.slot_off[pi] * 4
            L.check(L.lib.nk_reduce_exchange(self.comm_device.ctx, C.c_void_p(slots_local), grads, self.flags.table,
                                             self.world, self.rank, shard, self.state.ptr, self.reduce_ctas),
                    self.comm_device.ctx)

    def _nccl(self) -> None:
        """Context-owned NCCL communicator on the side stream (nk_comm_init_rank); the id travels over the process group."""
        import ctypes as C
        import torch.distributed as dist
        from . import _lib as L
        if self._nccl_ready:
            return
        box = [None]
        if self.rank == 0:
            buf = C.create_string_buffer(128)
            L.check(L.lib.nk_comm_unique_id(self.comm_device.ctx, buf), self.comm_device.ctx)
            box[0] = bytes(buf.raw)
        dist.broadcast_object_list(box, src=0)
        L.check(L.lib.nk_comm_init_rank(self.comm_device.ctx, self.world, self.rank, C.create_string_buffer(box[0], 128)),
                self.comm_device.ctx)
        self._nccl_ready = True

    def _exchange_rest(self) -> None:
        from . import _lib as L
        import ctypes as C
        rest = self.ranges.flush()
        if not rest:
            return
        ev = self._events[self._ev_i % len(self._events)]
        self._ev_i += 1
        ev.record(self.compute)
        self.comm.wait_event(ev)
        slot_cursor = 0
        for lo, hi in rest:
            n = hi - lo
            grad = self.bucket_mem.local + lo * 4
            if n <= self.SMALL_MAX and (slot_cursor + n) * self.world * 4 <= self.small_slots.nbytes:
                slots = self.small_slots.offset_table(slot_cursor * self.world * 4)
                L.check(L.lib.nk_peer_allreduce_small(self.comm_device.ctx, C.c_void_p(grad), slots,
                                                      self.flags.offset_table(256), self.world, self.rank, n,
                                                      C.c_void_p(self.state.ptr.value + 16)), self.comm_device.ctx)
                slot_cursor += n
            else:
                self._nccl()
                L.check(L.lib.nk_allreduce_sum(self.comm_device.ctx, C.c_void_p(grad), n, 0), self.comm_device.ctx)

    @property
    def launches(self) -> int:
        return self.comm_device.launches if self.comm_device is not None else 0

    def wait(self) -> None:
        self._flush_pending()
        if self.ranges.held:
            self._exchange_rest()
        self.compute.wait_stream(self.comm)
