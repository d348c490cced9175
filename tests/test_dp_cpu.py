"""World-size-2 data-parallel host logic on CPU (gloo): batch sharding + flat gradient bucket + all-reduce(sum)
+ SGD with grad_scale = 1/world reproduces the single-process full-batch step of the oracle.
No GPU, no device code: this covers the plumbing bench.py uses for N > 1."""
import os
import socket

import numpy as np
import pytest

F32 = np.float32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import oracle as O
    from neuronika_b200.parallel import BucketLayout, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                      # identical weights and data on every rank
    sizes = [12, 16, 5]
    params = []
    for i, o in zip(sizes[:-1], sizes[1:]):
        k = 1 / np.sqrt(i)
        params.append((O.uniform_init(rng, (o, i), k), O.uniform_init(rng, (o,), k)))
    x = rng.uniform(-1, 1, (8, sizes[0])).astype(F32)
    t = np.eye(sizes[-1], dtype=F32)[rng.integers(0, sizes[-1], 8)]
    lo, hi = shard_rows(8, rank, world)
    shapes = [a.shape for p in params for a in p]
    layout = BucketLayout(shapes)
    flat = np.zeros(layout.total, F32)
    views = layout.views(flat)
    # local backward on the shard (lr = 0 keeps the weights; grads land in the bucket views)
    local = [(w.copy(), b.copy()) for w, b in params]
    _, grads = O.mlp_step(x[lo:hi], t[lo:hi], local, lr=0.0)
    for li, (dw, db) in enumerate(grads):
        views[2 * li][...] = dw
        views[2 * li + 1][...] = db
    tt = torch.from_numpy(flat)
    dist.all_reduce(tt)                                  # sum over ranks, in place on the bucket
    for li, (w, b) in enumerate(params):
        O.sgd_step(w, views[2 * li] / world, 0.1)        # grad_scale = 1/world
        O.sgd_step(b, views[2 * li + 1] / world, 0.1)
    if rank == 0:
        q.put([p.copy() for wb in params for p in wb])
    dist.destroy_process_group()


def test_bucket_layout_and_sharding():
    from neuronika_b200.parallel import BucketLayout, shard_rows
    lay = BucketLayout([(4, 3), (4,), (2, 4), (2,), ()])
    assert lay.offsets == [0, 12, 16, 24, 26] and lay.total == 27
    flat = np.arange(27, dtype=F32)
    v = lay.views(flat)
    assert v[0].shape == (4, 3) and v[2][1, 0] == 20 and v[4].shape == ()
    v[1][...] = -1                                        # views alias the flat buffer
    assert np.all(flat[12:16] == -1)
    assert shard_rows(8192, 3, 8) == (3072, 4096)
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)


def test_two_rank_step_matches_single_process():
    import torch.multiprocessing as mp

    import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process full-batch reference
    rng = np.random.default_rng(0)
    sizes = [12, 16, 5]
    params = []
    for i, o in zip(sizes[:-1], sizes[1:]):
        k = 1 / np.sqrt(i)
        params.append((O.uniform_init(rng, (o, i), k), O.uniform_init(rng, (o,), k)))
    x = rng.uniform(-1, 1, (8, sizes[0])).astype(F32)
    t = np.eye(sizes[-1], dtype=F32)[rng.integers(0, sizes[-1], 8)]
    O.mlp_step(x, t, params, lr=0.1)
    want = [p for wb in params for p in wb]
    # mean-reduced loss: the average of the two shard gradients equals the full-batch gradient
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_ready_ranges_merges_small_pieces_with_neighbours():
    """Host logic of the overlapped exchange: big row blocks go out at once and absorb adjacent small pieces
    (the bias stored right after its weight); leftovers are flushed merged."""
    from neuronika_b200.parallel import ReadyRanges
    r = ReadyRanges(min_elems=100)
    W, b = 400, 10                      # weight at [0, 400), bias at [400, 410), next bias at [410, 415)
    assert r.add(W, W + b) == []        # bias first (its backward node runs before the matmul's)
    assert r.add(0, 100) == [(0, 100)]
    assert r.add(100, 200) == [(100, 200)]
    assert r.add(200, 300) == [(200, 300)]
    assert r.add(300, 400) == [(300, 410)]   # last block carries the bias
    assert r.flush() == []
    assert r.add(410, 415) == [] and r.add(415, 420) == [] and r.add(500, 505) == []
    assert r.flush() == [(410, 420), (500, 505)]
    assert r.flush() == []
    # every element is exchanged exactly once whatever the arrival order
    import itertools
    pieces = [(0, 150), (150, 160), (160, 400), (400, 405)]
    for order in itertools.permutations(pieces):
        rr = ReadyRanges(min_elems=100)
        out = []
        for lo, hi in order:
            out += rr.add(lo, hi)
        out += rr.flush()
        cover = sorted(out)
        assert cover[0][0] == 0 and cover[-1][1] == 405
        assert all(a[1] == b_[0] for a, b_ in zip(cover, cover[1:]))


def test_fused_exchange_eligibility():
    """which parameters take the fused reduce-scatter GEMM epilogue: matrices whose rows split into world shards of
    whole 128-row tiles; biases, the 10-wide output layer and conv kernels stay on the all-reduce"""
    from neuronika_b200.parallel import rs_eligible
    assert rs_eligible((4096, 4096), 2) and rs_eligible((4096, 4096), 8) and rs_eligible((4096, 1024), 8)
    assert not rs_eligible((4096,), 2) and not rs_eligible((10, 4096), 2) and not rs_eligible((64, 3, 3, 3), 2)
    assert not rs_eligible((4096, 4096), 1) and not rs_eligible((384, 4096), 2)      # 192 rows per rank: not whole tiles
    assert not rs_eligible((256, 256), 2)                                             # too small to matter


class _FakePeerLib:
    """Stand-in for the nk_ipc_* entry points (no GPU): 'allocates' fake addresses and hands out 64-byte handles that
    encode them, so the collective control flow of parallel.PeerMemory can run under gloo."""

    def __init__(self, rank, fail_alloc_on=None, fail_open_on=None):
        self.rank, self.fail_alloc_on, self.fail_open_on = rank, fail_alloc_on, fail_open_on
        self.freed, self.closed = [], []

    def nk_ipc_alloc(self, ctx, nbytes, out):
        if self.rank == self.fail_alloc_on:
            return -3
        out._obj.value = 0x10000000 * (self.rank + 1)
        return 0

    def nk_ipc_export(self, ctx, ptr, handle):
        handle.raw = int(ptr.value).to_bytes(8, "little") + bytes(56)
        return 0

    def nk_ipc_open(self, ctx, handle, out):
        if self.rank == self.fail_open_on:
            return -2
        out._obj.value = int.from_bytes(handle.raw[:8], "little") + 0x1000      # a different (mapped) address
        return 0

    def nk_ipc_close(self, ctx, ptr):
        self.closed.append(int(ptr.value))
        return 0

    def nk_ipc_free(self, ctx, ptr):
        self.freed.append(int(ptr.value))
        return 0


def _peer_worker(rank, world, port, q, fail_alloc_on, fail_open_on):
    import torch.distributed as dist

    from neuronika_b200 import _lib as L
    from neuronika_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fake = _FakePeerLib(rank, fail_alloc_on, fail_open_on)
    real_lib, real_check = L.lib, L.check
    L.lib = fake

    def check(rc, ctx=None):
        if rc != 0:
            raise L.NkError(rc, "fake failure")
    L.check = check

    class Dev:
        ctx = None
    try:
        m = parallel.PeerMemory(Dev(), 1024, world, rank)
        out = ("ok", m.ptrs, [int(v) for v in m.offset_table(16)])
    except RuntimeError as e:
        out = ("raised", str(e), (list(fake.freed), list(fake.closed)))
    finally:
        L.lib, L.check = real_lib, real_check
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_alloc_on,fail_open_on", [(None, None), (1, None), (None, 0)])
def test_peer_memory_setup_is_collective(fail_alloc_on, fail_open_on):
    """parallel.PeerMemory (the buffers of the fused NVLink exchange): every rank ends up with the same view of who
    owns which address, and a failure on ONE rank raises on ALL ranks (nobody is left waiting in a collective), after
    releasing what was mapped -- that is what lets bench.py fall back to the NCCL exchange consistently."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q, fail_alloc_on, fail_open_on)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if fail_alloc_on is None and fail_open_on is None:
        for r in (0, 1):
            kind, ptrs, off = res[r]
            assert kind == "ok"
            assert ptrs[r] == 0x10000000 * (r + 1)                          # own allocation
            assert ptrs[1 - r] == 0x10000000 * (2 - r) + 0x1000             # the peer's buffer, as mapped here
            assert off == [p + 16 for p in ptrs]
    else:
        assert res[0][0] == "raised" and res[1][0] == "raised"
        for r in (0, 1):
            freed, closed = res[r][2]
            if r != fail_alloc_on:
                assert freed == [0x10000000 * (r + 1)]                      # local buffer released on the way out
