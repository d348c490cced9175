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


class OverlappedAllReduce:
    """All-reduce groups of leaf gradients (contiguous slices of a GradientBucket) on a side stream as soon as
    backward has produced them, so the exchange overlaps the remaining backward kernels.

        sync = OverlappedAllReduce(bucket, compute_stream, groups=[[0, 1], [2, 3], [4, 5]], params=params)
        loss.backward(1.0)        # hooks fire inside; every group's all-reduce is in flight when this returns
        sync.wait()               # compute stream waits for the exchanges; then optimizer.step()
    """

    def __init__(self, bucket: GradientBucket, compute_stream, groups, params):
        import torch
        self.torch = torch
        self.bucket, self.compute, self.params = bucket, compute_stream, params
        self.comm = torch.cuda.Stream(device=bucket.array.device.index)
        self.event = torch.cuda.Event()
        flat = bucket.as_torch()
        self.groups = []
        lay = bucket.layout
        for g in groups:
            lo = lay.offsets[g[0]]
            last = g[-1]
            hi = lay.offsets[last] + (int(np.prod(lay.shapes[last])) if lay.shapes[last] else 1)
            self.groups.append({"members": list(g), "view": flat[lo:hi], "pending": len(g)})
        for gi, g in enumerate(self.groups):
            for pi in g["members"]:
                params[pi].set_grad_hook(lambda gi=gi: self._ready(gi))

    def _ready(self, gi: int) -> None:
        import torch.distributed as dist
        g = self.groups[gi]
        g["pending"] -= 1
        if g["pending"] > 0:
            return
        g["pending"] = len(g["members"])
        self.event.record(self.compute)          # everything launched so far (incl. the kernels that wrote the grads)
        self.comm.wait_event(self.event)
        with self.torch.cuda.stream(self.comm):
            dist.all_reduce(g["view"])

    def wait(self) -> None:
        self.compute.wait_stream(self.comm)
