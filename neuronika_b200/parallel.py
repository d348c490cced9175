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

    def __init__(self, device, shapes, dtype):
        from .device import CuArray
        self.layout = BucketLayout(shapes)
        self.array = CuArray(device, (self.layout.total,), dtype)
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
        self.event = torch.cuda.Event()
        self.flat = bucket.as_torch()
        self.ranges = ReadyRanges(min_elems=chunk_elems // 4)
        self.launched = 0
        lay = bucket.layout
        for pi, p in enumerate(params):
            n = int(np.prod(lay.shapes[pi])) if lay.shapes[pi] else 1
            chunks = max(1, min(max_chunks, n // chunk_elems))
            p.set_grad_hook(lambda b, e, off=lay.offsets[pi]: self._ready(off + b, off + e), row_chunks=chunks)

    def _exchange(self, pieces) -> None:
        import torch.distributed as dist
        if not pieces:
            return
        self.event.record(self.compute)          # everything launched so far (incl. the kernels that wrote the grads)
        self.comm.wait_event(self.event)
        with self.torch.cuda.stream(self.comm):
            for lo, hi in pieces:
                dist.all_reduce(self.flat[lo:hi])
                self.launched += 1

    def _ready(self, lo: int, hi: int) -> None:
        self._exchange(self.ranges.add(lo, hi))

    def wait(self) -> None:
        self._exchange(self.ranges.flush())
        self.compute.wait_stream(self.comm)
