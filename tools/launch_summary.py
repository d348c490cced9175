#!/usr/bin/env python
"""Turn the ncu launch lists of the bench command (gpurun_out/r01_launches_<workload>.csv, captured with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
        --log-file gpurun_out/r01_launches_<w>.csv python bench.py --workload <w> --profile --steps 2 --warmup 3
) into the per-kernel table of ONE step (the last one) for profiles/.  Run here, no GPU needed:
    python tools/launch_summary.py r01 linear mlp conv > profiles/r01_launches.md"""
import collections
import csv
import io
import re
import sys


def load(path):
    lines = [l for l in open(path).read().splitlines() if not l.startswith("==")]
    by = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("\n".join(lines))):
        d = by.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(by.values())


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")
    return name[:70]


def traffic_json(tag, workloads):
    """profiles/<tag>_traffic.json: measured DRAM bytes per launch of the kernels bench.py reports a roofline for."""
    import json
    out = {"source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/{tag}_launches.md"}
    for w in workloads:
        ks = load(f"gpurun_out/{tag}_launches_{w}.csv")
        last = ks[-(len(ks) // 5):]
        if w == "linear":
            g = [k for k in last if "gemm_tc_kernel" in k["name"]]
            out["gemm_tc_4096"] = {"bytes_per_launch": sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in g) / max(len(g), 1),
                                   "launches": len(g)}
        if w == "conv":
            for key, pat in (("conv_fwd_tc", "conv_fwd_t"), ("conv_bwd_fused_tc", "conv_bwd_fused_kernel")):
                g = [k for k in last if pat in k["name"]]
                if g:
                    out[key] = {"bytes_per_launch": sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in g) / len(g),
                                "launches": len(g)}
    with open(f"profiles/{tag}_traffic.json", "w") as f:
        json.dump(out, f, indent=1)


def main():
    tag, workloads = sys.argv[1], sys.argv[2:]
    traffic_json(tag, workloads)
    steps_captured = 5   # --warmup 3 --steps 2
    print(f"# {tag} -- ncu launch lists of `python bench.py --workload W --profile --steps 2 --warmup 3`\n")
    print("`--clock-control none`; per-launch times are cold-cache and serialised (ncu replays every kernel), so they are\n"
          "used for each kernel's SHARE of the step, not as absolute numbers -- the bench's own CUDA-event timing is the\n"
          "number of record.  One step = the last of the 5 captured.  DRAM bytes are per launch.\n")
    for w in workloads:
        ks = load(f"gpurun_out/{tag}_launches_{w}.csv")
        n = len(ks) // steps_captured
        last = ks[-n:]
        tot = sum(k["gpu__time_duration.sum"] for k in last)
        print(f"## {w}: {n} launches per step, {tot / 1e3:.1f} us summed under ncu\n")
        print("| # | kernel | grid | us | share | DRAM read MB | DRAM write MB |")
        print("|---|---|---|---|---|---|---|")
        for i, k in enumerate(last):
            t = k["gpu__time_duration.sum"]
            print(f"| {i} | `{short(k['name'])}` | {k['grid']} | {t / 1e3:.1f} | {100 * t / tot:.1f} % | "
                  f"{k.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {k.get('dram__bytes_write.sum', 0) / 1e6:.1f} |")
        agg = collections.OrderedDict()
        for k in last:
            key = re.sub(r"<.*", "", short(k["name"]))
            agg[key] = agg.get(key, 0.0) + k["gpu__time_duration.sum"]
        print("\nby kernel family: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])) + "\n")


if __name__ == "__main__":
    main()
