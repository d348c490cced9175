#!/usr/bin/env python
"""Hottest SASS instructions of one kernel in an .ncu-rep (warp-state samples per instruction, with their top stall reasons).
    python tools/ncu_hot.py gpurun_out/x.ncu-rep <kernel regex> [launch index] [top N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out))]
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    num = lambda v: int(v) if v.isdigit() else 0
    tot = sum(num(r[isamp]) for r in data)
    print(f"# {rows[0][1][:90] if rows and len(rows[0]) > 1 else pat}: {tot} samples, {len(data)} instructions")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    order = sorted(range(len(data)), key=lambda i: -num(data[i][isamp]))[:topn]
    for i in order:
        r = data[i]
        st = sorted(((num(r[c]), hdr[c][6:]) for c in stall_cols if num(r[c]) > 0), reverse=True)[:2]
        print(f"{i:5d} {num(r[isamp]):7d} {100.0 * num(r[isamp]) / max(tot, 1):5.1f}% ex={r[iex]:>9s}  {r[isrc].strip()[:64]:64s} "
              + " ".join(f"{n}={v}" for v, n in st))


if __name__ == "__main__":
    main()
