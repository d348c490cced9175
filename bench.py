#!/usr/bin/env python
"""bench.py -- forward+backward throughput of the dense hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 300 --warmup 10                      # own arm, default workload
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --steps 5 --warmup 1                 # the reference's CPU path (oracle port)

Workloads (BASELINE.json configs):
  linear  (default, config 2)  nn::Linear 4096->4096, bf16, batch 4096 per GPU, fwd + bwd (dX, dW, db)
                               N>1: weak scaling, NCCL all-reduce of the gradient bucket after backward
  mlp     (config 4)           MLP 1024-4096-4096-10 + ReLU/Softmax, MSE, SGD step; global batch 8192 sharded
                               over the ranks (strong scaling), all-reduce of the flat gradient bucket
  conv    (config 3)           nn::Conv2d 3->64 k3 s1, 224x224, batch 256 per GPU, fwd + bwd (dX, dW, db)

A step of the own arm = zero_grad -> forward -> backward (-> all-reduce -> SGD where the workload has them)
through the package's public API (Var/VarDiff/nn/optim over the C++ graph and the C ABI).  `value` times the
steps with inputs resident in HBM; `e2e` repeats them with the step's inputs copied from pinned host memory
and the loss read back inside the timed region.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "forward+backward GFLOP/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops_burst": float(p["bf16_tflops"]),
                "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------- workloads
def workload_spec(name: str, world: int):
    if name == "linear":
        n = fin = fout = 4096
        flops = 3 * 2.0 * n * fin * fout                     # fwd + dX + dW (bias terms < 0.05 %)
        return {"name": "nn::Linear 4096->4096 bf16, batch 4096 per GPU, fwd+bwd (dX+dW+db)", "scaling": "weak",
                "flops_per_rank_step": flops, "samples_per_rank_step": n, "batch": n, "fin": fin, "fout": fout,
                "params": fin * fout + fout}
    if name == "mlp":
        gb = 8192
        assert gb % world == 0
        b = gb // world
        sizes = [1024, 4096, 4096, 10]
        fwd = sum(2.0 * b * i * o for i, o in zip(sizes[:-1], sizes[1:]))
        bwd = 2 * fwd - 2.0 * b * sizes[0] * sizes[1]        # dW for all layers, dX for layers 2,3
        return {"name": "MLP 1024-4096-4096-10 ReLU/Softmax MSE SGD step, global batch 8192", "scaling": "strong",
                "flops_per_rank_step": fwd + bwd, "samples_per_rank_step": b, "batch": b, "sizes": sizes,
                "params": sum(i * o + o for i, o in zip(sizes[:-1], sizes[1:]))}
    if name == "conv":
        n, cin, h, w, cout, k = 256, 3, 224, 224, 64, 3
        ho, wo = h - k + 1, w - k + 1
        fwd = 2.0 * n * cout * ho * wo * cin * k * k
        return {"name": "nn::Conv2d 3->64 k3 s1 p0, 224x224, batch 256 per GPU, fwd+bwd (dX+dW+db)",
                "scaling": "weak", "flops_per_rank_step": 3 * fwd, "samples_per_rank_step": n, "batch": n,
                "shape": (n, cin, h, w), "cout": cout, "k": k, "params": cout * cin * k * k + cout,
                "bytes_per_step_bf16": 2.0 * (3 * n * cout * ho * wo + 3 * n * cin * h * w)}
    raise SystemExit(f"unknown workload {name}")


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period_s, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if get_reasons:
                    r = int(get_reasons(self.h))
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------- own arm
class _CAI:
    def __init__(self, arr):
        self.__cuda_array_interface__ = arr.cuda_array_interface()


def as_torch(arr, torch, index):
    t = torch.as_tensor(_CAI(arr), device=f"cuda:{index}")
    return t.view(torch.bfloat16) if arr.dtype == 1 else t


def run_own(args):
    import torch
    import torch.distributed as dist

    import neuronika_b200 as nk
    from neuronika_b200 import variable as V
    from neuronika_b200.parallel import FusedGradientExchange, GradientBucket, OverlappedAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = torch.cuda.Stream(device=local)
    dev = nk.Device(local, stream=stream.cuda_stream)
    spec = workload_spec(args.workload, world)
    rng = np.random.default_rng(0)          # identical initial weights on every rank
    drng = np.random.default_rng(1000 + rank)  # per-rank data shard
    peaks = load_peaks()
    BF = nk.BF16
    gdt = nk.F32 if args.grad_dtype == "f32" else nk.BF16

    # ---- parameters with gradients in ONE contiguous bucket (single all-reduce)
    exchange_kind = os.environ.get("NK_DP_EXCHANGE", "fused") if world > 1 else "none"
    if gdt != nk.F32 and exchange_kind == "fused":
        exchange_kind = "nccl"       # the fused reduce-scatter epilogue sums f32 gradients
    fused_exchange = []
    exchange_state = {"kind": exchange_kind}

    def make_bucket(shapes):
        if world == 1:      # no exchange step: let the graph own (and lazily clear) the gradients
            return None, [None] * len(shapes)
        if exchange_kind == "fused":
            try:
                ex = FusedGradientExchange(dev, stream, shapes, world, rank,
                                           reduce_ctas=int(os.environ.get("NK_DP_REDUCE_CTAS", "20")))
                fused_exchange.append(ex)
                return ex.bucket, ex.bucket.views
            except RuntimeError as e:
                # peer memory could not be set up (the error is collective: every rank gets here together): the
                # exchange falls back from the fused NVLink path to NCCL all-reduce, and the JSON line says so
                if rank == 0:
                    print(f"bench: fused exchange unavailable ({e}); using the NCCL all-reduce exchange", file=sys.stderr)
                exchange_state["kind"] = "nccl"
        b = GradientBucket(dev, shapes, gdt)
        return b, b.views

    def param(values, grad_view):
        return V.from_ndarray(dev, values, BF).requires_grad(gdt, grad_view)


    def pinned_bf16(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(torch.bfloat16).pin_memory()
        return t

    if args.workload == "linear":
        n, fin, fout = spec["batch"], spec["fin"], spec["fout"]
        k = 1.0 / np.sqrt(fin)
        bucket, (gw, gb) = make_bucket([(fout, fin), (fout,)])
        W = param(rng.uniform(-k, k, (fout, fin)).astype(np.float32), gw)
        b = param(rng.uniform(-k, k, (fout,)).astype(np.float32), gb)
        params = [W, b]
        x_host = drng.uniform(-1, 1, (n, fin)).astype(np.float32)
        t_host = drng.uniform(-1, 1, (n, fout)).astype(np.float32)
        px, pt = pinned_bf16(x_host), pinned_bf16(t_host)

        def make_inputs():
            x = V.from_ndarray(dev, x_host, BF).requires_grad()   # input as VarDiff => dX is computed
            t = V.from_ndarray(dev, t_host, BF)
            return {"x": x, "t": t, "copies": [(px, x), (pt, t)]}
        opt = None
        live = {}

        # every step builds its graph anew (define-by-run, as a user of the reference does each iteration: the op
        # nodes and their intermediate tensors / gradients are fresh, the leaves persist)
        def step_resident(inp):         # root = y, backward(seed)
            for p in params:
                p.zero_grad()
            inp["x"].zero_grad()
            y = inp["x"].mm_t(W) + b
            y.forward()
            y.backward(1.0 / (n * fout))
            live["root"] = y

        def step_e2e_compute(inp):      # loss = mse(y, t)
            for p in params:
                p.zero_grad()
            inp["x"].zero_grad()
            loss = (inp["x"].mm_t(W) + b).mse_loss(inp["t"])
            loss.forward()
            loss.backward(1.0)
            live["root"] = loss
    elif args.workload == "mlp":
        sizes, bsz = spec["sizes"], spec["batch"]
        shapes = []
        for i, o in zip(sizes[:-1], sizes[1:]):
            shapes += [(o, i), (o,)]
        bucket, gviews = make_bucket(shapes)
        params = []
        for li, (i, o) in enumerate(zip(sizes[:-1], sizes[1:])):
            k = 1.0 / np.sqrt(i)
            params.append(param(rng.uniform(-k, k, (o, i)).astype(np.float32), gviews[2 * li]))
            params.append(param(rng.uniform(-k, k, (o,)).astype(np.float32), gviews[2 * li + 1]))
        x_host = drng.uniform(-1, 1, (bsz, sizes[0])).astype(np.float32)
        t_host = np.eye(10, dtype=np.float32)[np.argmax(x_host[:, :10], 1)]
        px, pt = pinned_bf16(x_host), pinned_bf16(t_host)

        def make_inputs():
            x = V.from_ndarray(dev, x_host, BF)
            t = V.from_ndarray(dev, t_host, BF)
            return {"x": x, "t": t, "copies": [(px, x), (pt, t)]}
        opt = nk.optim.StochasticGD.new(0.01, nk.optim.L2(0.0), grad_scale=1.0 / world,
                                        master_weights=args.master_weights)
        for p in params:
            opt.register(p)
        live = {}

        def step_resident(inp):         # a training iteration as written against the reference: new graph every step
            opt.zero_grad()
            h = inp["x"]
            for li in range(3):
                h = h.mm_t(params[2 * li]) + params[2 * li + 1]
                h = h.relu() if li < 2 else h.softmax(1)
            loss = h.mse_loss(inp["t"])
            loss.forward()
            loss.backward(1.0)
            live["root"] = loss
        step_e2e_compute = step_resident
    else:  # conv
        n, cin, hh, ww = spec["shape"]
        cout, ks = spec["cout"], spec["k"]
        k = 1.0 / np.sqrt(cin * ks * ks)
        bucket, (gw, gb) = make_bucket([(cout, cin, ks, ks), (cout, 1, 1)])
        Wc = param(rng.uniform(-k, k, (cout, cin, ks, ks)).astype(np.float32), gw)
        bc = param(rng.uniform(-k, k, (cout, 1, 1)).astype(np.float32), gb)
        params = [Wc, bc]
        x_host = drng.uniform(0, 1, (n, cin, hh, ww)).astype(np.float32)
        px = pinned_bf16(x_host)

        def make_inputs():
            x = V.from_ndarray(dev, x_host, BF).requires_grad()
            return {"x": x, "copies": [(px, x)]}
        opt = None
        live = {}

        def step_resident(inp):
            for p in params:
                p.zero_grad()
            inp["x"].zero_grad()
            y = Wc.convolution(inp["x"], (1, 1), (1, 1), 1) + bc
            y.forward()
            y.backward(1.0 / 1e6)
            live["root"] = y

        def step_e2e_compute(inp):
            for p in params:
                p.zero_grad()
            inp["x"].zero_grad()
            loss = (Wc.convolution(inp["x"], (1, 1), (1, 1), 1) + bc).mean()
            loss.forward()
            loss.backward(1.0)
            live["root"] = loss

    # N > 1: every layer's (W, b) slice of the bucket is all-reduced on a side stream as soon as backward has
    # produced it (gradient-ready hooks), overlapping the exchange with the remaining backward kernels
    sync = None
    if world > 1 and fused_exchange:
        # the dW GEMM epilogue pushes gradient shards to their owners over NVLink; reduce + broadcast on a side stream
        sync = fused_exchange[0]
        sync.attach(params)
    elif world > 1:
        sync = OverlappedAllReduce(bucket, stream, params, max_chunks=int(os.environ.get("NK_DP_CHUNKS", "1")))

    def exchange_and_update():
        if sync is not None:
            sync.wait()
        if opt is not None:
            opt.step()

    def full_step(compute):
        compute()
        exchange_and_update()

    # two sets of input leaves: while step k computes on one, the pinned-host -> HBM copy of step k+1 fills the other
    # on a copy stream (the input pipeline a training loop runs), and the loss of step k is read back one step late
    sets = [make_inputs(), make_inputs()]
    copy_stream = torch.cuda.Stream(device=local)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    loss_pinned = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_views = [[(src, as_torch(var.data_array(), torch, local)) for src, var in st["copies"]] for st in sets]

    def issue_copy(k):
        i = k & 1
        copy_stream.wait_event(consumed[i])          # the step that last read this set has finished with it
        with torch.cuda.stream(copy_stream):
            for src, dst in h2d_views[i]:
                dst.copy_(src.view(dst.shape), non_blocking=True)
        copied[i].record(copy_stream)

    def e2e_run(steps):
        """`steps` end-to-end iterations; returns the losses read back (one per step)."""
        losses = []
        for i in range(2):
            consumed[i].record(stream)
        issue_copy(0)
        for k in range(steps):
            i = k & 1
            if k + 1 < steps:
                issue_copy(k + 1)
            stream.wait_event(copied[i])
            full_step(lambda: step_e2e_compute(sets[i]))
            loss_t = as_torch(live["root"].data_array(), torch, local)      # this step's loss scalar
            with torch.cuda.stream(stream):
                loss_pinned[i].copy_(loss_t.view(()), non_blocking=True)
            loss_ready[i].record(stream)
            consumed[i].record(stream)
            if k >= 1:                                   # read the previous step's loss while this one runs
                loss_ready[i ^ 1].synchronize()
                losses.append(float(loss_pinned[i ^ 1]))
        loss_ready[(steps - 1) & 1].synchronize()
        losses.append(float(loss_pinned[(steps - 1) & 1]))
        return losses

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = dev.launches + (sync.launches if hasattr(sync, 'launches') else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        for _ in range(steps):
            fn()
        timed.host_ms = (time.perf_counter() - h0) * 1e3 / steps   # host enqueue time per step (diagnostic)
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        launches = dev.launches + (sync.launches if hasattr(sync, 'launches') else 0) - l0
        if world > 1:
            tt = torch.tensor([ms], device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, launches, clocks

    W_ = max(args.warmup, 3)
    ms, launches, clocks = timed(lambda: full_step(lambda: step_resident(sets[0])), args.steps, W_, sample_clocks=True)
    host_ms = timed.host_ms
    if args.profile:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "profile_only": True, "ms_per_step": round(ms / args.steps, 5),
                              "gpu_launches": int(launches), "workload": spec["name"]}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    e2e_steps = max(3, min(args.steps, args.e2e_steps))

    def timed_e2e(steps):
        e2e_run(3)                                      # warm-up (pinned buffers, allocator pools)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)                               # every copy of the run is ordered after this point ...
        losses = e2e_run(steps)
        e1.record(stream)                               # ... and the last loss read-back before this one
        e1.synchronize()
        torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms_], device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_ = float(tt.item())
        return ms_, losses

    ms_e2e, e2e_losses = timed_e2e(e2e_steps)

    flops_step_total = spec["flops_per_rank_step"] * world
    value = flops_step_total / (ms / args.steps * 1e-3) / 1e9
    e2e_value = flops_step_total / (ms_e2e / e2e_steps * 1e-3) / 1e9
    h2d = sum(int(src.numel()) * 2 for src, _ in sets[0]["copies"])

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": W_, "ms_per_step": round(ms / args.steps, 5), "higher_is_better": True,
        "scaling": spec["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "samples_per_s": round(spec["samples_per_rank_step"] * world / (ms / args.steps * 1e-3), 1),
        "config": {"workload": spec["name"], "grad_dtype": args.grad_dtype, "parallelism": f"dp{world}", "exchange": exchange_state["kind"],
                   "step": "zero_grad -> forward -> backward" + ({"none": "", "nccl": " (+ overlapped nccl all_reduce of each layer's grad slice)", "fused": " (dW GEMM epilogue reduce-scatters over NVLink peer memory, owner reduce + broadcast on a side stream; nccl for the small tensors)"}[exchange_state["kind"]])
                           + (" -> sgd" if opt is not None else ""),
                   "l2": "working set per step exceeds the 126 MB L2 (no flush needed)",
                   "kernels": {"gemm": dev.last_gemm_kernel, "conv": dev.last_conv_kernel}},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 5), "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / e2e_steps, 5), "steps": e2e_steps,
                "last_loss": round(float(e2e_losses[-1]), 6),
                "what": "per step: pinned host -> HBM copy of the step's inputs (double-buffered on a copy stream, so the "
                        "copy of step k+1 overlaps the compute of step k), graph build, forward, backward"
                        + (", exchange" if world > 1 else "") + (", sgd" if opt is not None else "")
                        + ", loss scalar read back to pinned host memory (read one step late)"},
    }

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream
    if rank == 0:
        out["roofline"] = roofline(args, dev, nk, spec, peaks, stream, torch)
        out["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measured_traffic(key):
    """DRAM bytes per launch measured with ncu for the roofline kernels (profiles/r01_traffic.json, committed with the
    launch lists it was extracted from); None when the file does not have the kernel."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")) as f:
            return float(json.load(f)[key]["bytes_per_launch"])
    except Exception:
        return None


def roofline(args, dev, nk, spec, peaks, stream, torch):
    from neuronika_b200 import ops
    iters = max(20, min(200, args.steps))
    if args.workload in ("linear", "mlp"):
        n = 4096
        a = nk.CuArray(dev, (n, n), nk.BF16)
        b = nk.CuArray(dev, (n, n), nk.BF16)
        c = nk.CuArray(dev, (n, n), nk.BF16)
        a.fill_(0.01)
        b.fill_(0.01)
        forms = {}
        for name, kw in (("nt_fwd", dict(trans_b=True)), ("nn_dx", dict()), ("tn_dw", dict(trans_a=True))):
            for _ in range(3):
                ops.gemm(a, b, c, **kw)
            stream.synchronize()
            dev.timer_start()
            for _ in range(iters):
                ops.gemm(a, b, c, **kw)
            forms[name] = dev.timer_stop() / iters
        dur = float(np.mean(list(forms.values())))
        achieved = 2.0 * n ** 3 / (dur * 1e-3) / 1e12
        peak = peaks["tflops_sustained"]
        return {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                "frac": round(achieved / peak, 4), "traffic": measured_traffic("gemm_tc_4096"),
                "traffic_note": "tensor-bound kernel: DRAM bytes per launch (ncu, profiles/r01_launches.md) vs 96-128 MB of operands + output",
                "kernel": "gemm_tc_kernel (tcgen05, 128x256x64 tiles); mean of NT/NN/TN 4096^3 launches",
                "per_form_ms": {k: round(v, 5) for k, v in forms.items()}, "launches_timed": 3 * iters,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); burst {peaks['tflops_burst']}",
                "algorithmic_flops_per_launch": 2.0 * n ** 3}
    # conv: HBM bound
    nb, cin, h, w = spec["shape"]
    cout, k = spec["cout"], spec["k"]
    x = nk.CuArray(dev, (nb, cin, h, w), nk.BF16)
    wt = nk.CuArray(dev, (cout, cin, k, k), nk.BF16)
    x.fill_(0.5)
    wt.fill_(0.01)
    y = nk.CuArray(dev, (nb, cout, h - k + 1, w - k + 1), nk.BF16)
    for _ in range(2):
        ops.conv2d(x, wt, out=y)
    stream.synchronize()
    iters = min(iters, 20)
    dev.timer_start()
    for _ in range(iters):
        ops.conv2d(x, wt, out=y)
    dur = dev.timer_stop() / iters
    bytes_alg = 2.0 * (x.size + y.size)
    achieved = bytes_alg / (dur * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": measured_traffic("conv_fwd_tc"),
            "kernel": f"conv2d forward ({dev.last_conv_kernel})", "ms": round(dur, 5),
            "algorithmic_bytes_per_launch": bytes_alg, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['source']})"}


# ------------------------------------------------------------------------------------------- CPU arms
def cpu_step_fn(workload: str, sample_batch: int):
    """One forward+backward of the workload on the host with the oracle (the reference's algorithm
    restated in numpy f32 -- the reference itself cannot be built here).  Returns (fn, flops, description)."""
    import oracle as O
    rng = np.random.default_rng(0)
    if workload in ("linear", "mlp"):
        n, fin, fout = sample_batch, 4096, 4096
        k = 1 / 64.0
        x = rng.uniform(-1, 1, (n, fin)).astype(np.float32)
        w = rng.uniform(-k, k, (fout, fin)).astype(np.float32)
        b = rng.uniform(-k, k, (fout,)).astype(np.float32)
        g = (rng.uniform(-1, 1, (n, fout)) / n).astype(np.float32)

        def fn():
            y = O.linear_forward(x, w, b)
            dx, dw, db = np.zeros_like(x), np.zeros_like(w), np.zeros_like(b)
            O.linear_backward(x, w, g, dx, dw, db)
            return y
        return fn, 3 * 2.0 * n * fin * fout, f"Linear 4096->4096 fwd+bwd on a batch of {n} rows (of 4096), numpy f32"
    n = sample_batch
    x = rng.uniform(0, 1, (n, 3, 224, 224)).astype(np.float32)
    w = rng.uniform(-0.19, 0.19, (64, 3, 3, 3)).astype(np.float32)
    b = rng.uniform(-0.19, 0.19, (64, 1, 1)).astype(np.float32)

    def fn():
        y = O.conv2d_layer_forward(x, w, b)
        g = np.full(y.shape, 1.0 / y.size, np.float32)
        dx, dw, db = np.zeros_like(x), np.zeros_like(w), np.zeros_like(b)
        O.conv2d_layer_backward(x, w, g, dx, dw, db)
        return y
    return fn, 3 * 2.0 * n * 64 * 222 * 222 * 27, f"Conv2d 3->64 k3 224x224 fwd+bwd on a batch of {n} (of 256), im2col + sgemm in numpy f32"


def cpu_baseline(workload: str, budget_s: float = 15.0):
    cores = os.cpu_count() or 1
    sample = 512 if workload != "conv" else 4
    fn, flops, what = cpu_step_fn(workload, sample)
    fn()
    t0 = time.perf_counter()
    reps = 0
    while True:
        fn()
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 50:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(flops / dt / 1e9, 2), "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": f"{what}; {reps} repetitions, {dt:.3f} s each, BLAS threads = all {cores} host cores"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be
    built in this environment (no rustc/cargo), so this times the oracle port on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    spec = workload_spec(args.workload, max(1, world))
    cores = os.cpu_count() or 1
    sample = 512 if args.workload != "conv" else 4
    fn, flops, what = cpu_step_fn(args.workload, sample)
    for _ in range(max(1, min(args.warmup, 2))):
        fn()
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = (time.perf_counter() - t0) / steps
    val = round(flops / dt / 1e9, 2)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
           "scaling": spec["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": spec["name"], "parallelism": "cpu"},
           "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                            "sample": f"{what}; one step = one fwd+bwd of the sample"},
           "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="linear", choices=["linear", "mlp", "conv"])
    ap.add_argument("--grad-dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--master-weights", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--profile", action="store_true",
                    help="only the warm-up + timed steps (no e2e / roofline / cpu legs): for ncu launch lists")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
