#!/usr/bin/env python
"""Summarise an .ncu-rep captured on the GPU box (run HERE, no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [object.o kernel_substring source.cu]
Prints the headline metrics and, when an object file is given, the hottest source lines by
executed warp instructions (SASS address -> line through nvdisasm --print-line-info)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg"]
STALLS = "smsp__average_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))


def main():
    rep = sys.argv[1]
    vals, units = raw(rep)
    print(f"# {os.path.basename(rep)}: {vals.get('Kernel Name', '')[:80]}")
    for k in KEYS:
        if k in vals:
            print(f"{k:70s} {vals[k]:>16s} {units.get(k, '')}")
    st = sorted(((float(v), k[len(STALLS):].replace('_per_issue_active.ratio', '')) for k, v in vals.items()
                 if k.startswith(STALLS) and k.endswith("per_issue_active.ratio") and v), reverse=True)
    print("stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
    if len(sys.argv) < 5:
        return
    obj, kname, src = sys.argv[2], sys.argv[3], sys.argv[4]
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    h = rows[1]
    ia, iex, ism = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout
    lines = dis.split("\n")
    start = end = None
    for i, l in enumerate(lines):
        if l.startswith(".text.") and kname in l and start is None:
            start = i
        elif l.startswith(".text.") and start is not None and i > start:
            end = i
            break
    cur, off2line = None, {}
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    addrs = [int(r[ia], 16) for r in rows[2:] if r and r[ia]]
    base = min(addrs)
    agg, smp, tot, ts = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[2:]:
        if not r or not r[ia]:
            continue
        ln = off2line.get(int(r[ia], 16) - base)
        e, s = int(r[iex] or 0), int(r[ism] or 0)
        agg[ln] += e
        smp[ln] += s
        tot += e
        ts += s
    text = {os.path.basename(src): open(src).read().split("\n")}
    hdr = os.path.join(os.path.dirname(src), "nk_ptx.cuh")
    if os.path.exists(hdr):
        text["nk_ptx.cuh"] = open(hdr).read().split("\n")
    print(f"warp instructions executed: {tot}")
    for ln, e in agg.most_common(18):
        t = ""
        if ln and ln[0] in text and ln[1] - 1 < len(text[ln[0]]):
            t = text[ln[0]][ln[1] - 1].strip()[:78]
        print(f"  {str(ln):28s} {e / tot * 100:5.1f}% inst {smp[ln] / max(1, ts) * 100:5.1f}% samples  {t}")


if __name__ == "__main__":
    main()
