"""Two-or-more-GPU worker for tests/test_gpu_dp.py (launched with torchrun, one process per GPU):
the fused data-parallel exchange (nk_gemm_rs -> nk_reduce_exchange, nk_peer_allreduce_small) must hand every replica the
same gradients as computing dW locally and summing it with an NCCL all-reduce; then nk_allreduce_sum (the context-owned
NCCL communicator behind the C ABI) on its own."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import neuronika_b200 as nk
    from neuronika_b200 import variable as V
    from neuronika_b200.parallel import FusedGradientExchange

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream(device=local)
    dev = nk.Device(local, stream=stream.cuda_stream)
    B, I, O = 512, 384, 256 * world                      # dW (O, I): world shards of 256 rows
    shapes = [(O, I), (O,)]
    ex = FusedGradientExchange(dev, stream, shapes, world, rank, min_elems=1)
    assert ex.fused == [0], ex.fused
    rng = np.random.default_rng(0)
    w0 = rng.uniform(-0.1, 0.1, (O, I)).astype(np.float32)
    b0 = rng.uniform(-0.1, 0.1, (O,)).astype(np.float32)
    drng = np.random.default_rng(100 + rank)
    W = V.from_ndarray(dev, w0, nk.BF16).requires_grad(nk.F32, ex.bucket.views[0])
    b = V.from_ndarray(dev, b0, nk.BF16).requires_grad(nk.F32, ex.bucket.views[1])
    ex.attach([W, b])
    # reference replica in the same process: plain graph gradients + NCCL all-reduce
    W2 = V.from_ndarray(dev, w0, nk.BF16).requires_grad(nk.F32)
    b2 = V.from_ndarray(dev, b0, nk.BF16).requires_grad(nk.F32)
    ok = True
    for step in range(3):
        x = drng.uniform(-1, 1, (B, I)).astype(np.float32)
        t = drng.uniform(-1, 1, (B, O)).astype(np.float32)
        outs = []
        for (w_, b_) in ((W, b), (W2, b2)):
            w_.zero_grad()
            b_.zero_grad()
            loss = (V.from_ndarray(dev, x, nk.BF16).mm_t(w_) + b_).relu().mse_loss(V.from_ndarray(dev, t, nk.BF16))
            loss.forward()
            loss.backward(1.0)
            if w_ is W:
                ex.wait()
            stream.synchronize()
            outs.append((w_.grad().copy(), b_.grad().copy()))
        assert ex.pushed == step + 1, ex.pushed
        gw = torch.from_numpy(outs[1][0]).cuda()
        gb = torch.from_numpy(outs[1][1]).cuda()
        dist.all_reduce(gw)
        dist.all_reduce(gb)
        want_w, want_b = gw.cpu().numpy(), gb.cpu().numpy()
        scale = float(np.abs(want_w).max()) + 1e-12
        err_w = float(np.abs(outs[0][0] - want_w).max())
        err_b = float(np.abs(outs[0][1] - want_b).max())
        exact = world == 2 and np.array_equal(outs[0][0], want_w)     # a + b is order independent for two ranks
        good = err_w <= 1e-5 * scale and err_b <= 1e-5 * (float(np.abs(want_b).max()) + 1e-12) and (world != 2 or exact)
        ok = ok and good
        # accumulate-into-existing-gradient falls back to the local GEMM + all-reduce path
    # second backward without zero_grad: gradient not zero -> no push, nccl fallback, result = 2x
    loss = (V.from_ndarray(dev, x, nk.BF16).mm_t(W) + b).relu().mse_loss(V.from_ndarray(dev, t, nk.BF16))
    W.zero_grad()
    b.zero_grad()
    loss.forward()
    loss.backward(1.0)
    ex.wait()
    stream.synchronize()
    once = W.grad().copy()
    pushed_before = ex.pushed
    loss.backward(1.0)      # accumulates locally on top of the already-summed gradient, then all-reduces the range
    ex.wait()
    stream.synchronize()
    assert ex.pushed == pushed_before
    # ---- the plain collective behind the C ABI (nk_comm_*, nk_allreduce_sum): what a host without peer-memory set-up (or
    # without Python) would call after backward().  Only the 128-byte id travels over the process group.
    import ctypes as C
    from neuronika_b200 import _lib as L
    box = [None]
    if rank == 0:
        idbuf = C.create_string_buffer(128)
        L.check(L.lib.nk_comm_unique_id(dev.ctx, idbuf), dev.ctx)
        box[0] = bytes(idbuf.raw)
    dist.broadcast_object_list(box, src=0)
    L.check(L.lib.nk_comm_init_rank(dev.ctx, world, rank, C.create_string_buffer(box[0], 128)), dev.ctx)
    assert L.lib.nk_comm_world(dev.ctx) == world and L.lib.nk_comm_rank(dev.ctx) == rank
    n_ar = 1 << 20
    base = np.arange(n_ar, dtype=np.float32) % 251
    mine = dev.from_ndarray(base * (rank + 1), nk.F32)
    L.check(L.lib.nk_allreduce_sum(dev.ctx, C.c_void_p(mine.ptr.value), n_ar, 0), dev.ctx)
    stream.synchronize()
    ok = ok and bool(np.array_equal(mine.as_ndarray(), base * (world * (world + 1) // 2)))   # small integers: exact in f32
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_WORKER", "OK" if int(flag.item()) == 1 else "FAIL", f"world={world} err_w={err_w:.3e} err_b={err_b:.3e}",
              flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
