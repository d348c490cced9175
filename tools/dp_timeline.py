#!/usr/bin/env python
"""Kernel timeline of one data-parallel step (development tool; the nsys substitute on this image): runs a bench workload
eagerly under torch.profiler (CUPTI sees every kernel of the process, ours included) and prints, for rank 0, each kernel of
the last profiled step with its stream, start time relative to the step's first kernel, and duration.
    python -m torch.distributed.run --nproc-per-node N ... tools/dp_timeline.py <linear|mlp|conv|convnet> [graph]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile

    import bench
    import neuronika_b200 as nk
    from neuronika_b200 import variable as V

    name = sys.argv[1] if len(sys.argv) > 1 else "linear"
    use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
    args = argparse.Namespace(grad_dtype="f32", master_weights=False, graph=use_graph, fusion=2)
    V.set_fusion(2)
    env = bench.Env()
    env.args, env.torch, env.dist = args, torch, dist
    world = env.world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = env.rank = int(os.environ.get("RANK", "0"))
    local = env.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    stream = env.stream = torch.cuda.Stream(device=local)
    env.dev = nk.Device(local, stream=stream.cuda_stream)
    wl = bench.Workload(env, name)
    fn, graphed = wl.runner("res", wl.step_resident, wl.sets[0])
    for _ in range(5):
        fn()
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = 4
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            fn()
        stream.synchronize()
        torch.cuda.synchronize()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
        evs.sort(key=lambda e: e.time_range.start)
        per = len(evs) // steps
        last = evs[-per:]
        t0 = last[0].time_range.start
        print(f"# {name} world={world} graph={graphed}: {per} device activities per step; last step:")
        for e in last:
            nm = e.name.replace("void ", "").replace("(anonymous namespace)::", "")[:70]
            print(f"{(e.time_range.start - t0):9.1f} us  +{e.time_range.elapsed_us():8.1f} us  stream {getattr(e, 'stream', getattr(e, 'device_resource_id', '?'))}  {nm}")
        print(f"# step span {(last[-1].time_range.end - t0):.1f} us")
    wl.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
