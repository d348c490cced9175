#!/usr/bin/env python
"""Hardware probe for the tcgen05 GEMM engine: runs each (form, shape, out dtype) in its own
subprocess (a trap poisons the CUDA context), prints max error vs a numpy f32 product of the
same bf16-rounded operands, and timing.  Development tool, not part of the product path.

    python tools/gemm_probe.py            # all cases
    python tools/gemm_probe.py one NT 256 256 256 f32   # a single case (used by the driver loop)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bf16r(x):
    import numpy as np
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + (((b >> 16) & 1) + 0x7FFF)) & 0xFFFF0000).astype(np.uint32)
    return b.view(np.float32).reshape(x.shape)


def one(form, M, N, K, cdt, beta=0.0, iters=0):
    import numpy as np
    import neuronika_b200 as nk
    from neuronika_b200 import ops
    dev = nk.Device(0)
    dev.gemm_engine("tcgen05")
    rng = np.random.default_rng(1)
    ta, tb = form[0] == "T", form[1] == "T"
    a = bf16r(rng.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32))
    b = bf16r(rng.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32))
    c0 = bf16r(rng.uniform(-1, 1, (M, N)).astype(np.float32))
    da, db = dev.from_ndarray(a, nk.BF16), dev.from_ndarray(b, nk.BF16)
    dc = dev.from_ndarray(c0, nk.BF16 if cdt == "bf16" else nk.F32)
    ops.gemm(da, db, dc, trans_a=ta, trans_b=tb, beta=beta)
    got = dc.as_ndarray()
    ref = (a.T if ta else a).astype(np.float64) @ (b if not tb else b.T).astype(np.float64) + beta * c0
    err = np.abs(got - ref)
    res = {"form": form, "M": M, "N": N, "K": K, "c": cdt, "beta": beta, "kernel": dev.last_gemm_kernel,
           "max_err": float(err.max()), "ref_rms": float(np.sqrt((ref ** 2).mean())),
           "bad_frac": float((err > 0.05 * max(1.0, np.sqrt(K) / 8)).mean())}
    if res["bad_frac"] > 0:
        bad = np.argwhere(err > 0.05 * max(1.0, np.sqrt(K) / 8))
        res["first_bad"] = bad[:6].tolist()
        res["bad_rows"] = int(len(set(bad[:, 0].tolist())))
        res["bad_cols"] = int(len(set(bad[:, 1].tolist())))
    if iters:
        for _ in range(3):
            ops.gemm(da, db, dc, trans_a=ta, trans_b=tb, beta=beta)
        dev.synchronize()
        dev.timer_start()
        for _ in range(iters):
            ops.gemm(da, db, dc, trans_a=ta, trans_b=tb, beta=beta)
        ms = dev.timer_stop() / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
    print("PROBE " + json.dumps(res), flush=True)


CASES = [
    ("NT", 128, 256, 64, "f32", 0.0, 0), ("NT", 128, 256, 256, "f32", 0.0, 0), ("NT", 256, 512, 512, "f32", 0.0, 0),
    ("NN", 128, 256, 64, "f32", 0.0, 0), ("NN", 256, 512, 512, "f32", 0.0, 0),
    ("TN", 128, 256, 64, "f32", 0.0, 0), ("TN", 256, 512, 512, "f32", 0.0, 0),
    ("TT", 256, 512, 512, "f32", 0.0, 0),
    ("NT", 200, 72, 136, "f32", 0.0, 0), ("NN", 200, 72, 136, "bf16", 1.0, 0), ("TN", 200, 72, 136, "f32", 1.0, 0),
    ("NT", 1000, 10, 4096, "f32", 0.0, 0), ("NT", 300, 130, 1000, "bf16", 0.0, 0),
    ("NT", 4096, 4096, 4096, "bf16", 0.0, 20), ("NN", 4096, 4096, 4096, "bf16", 0.0, 20),
    ("TN", 4096, 4096, 4096, "f32", 1.0, 20), ("NT", 8192, 4096, 1024, "bf16", 0.0, 20),
]

# the CTA-pair (cta_group::2) variant: every form, tails in M / N / K, accumulate, then the timed shapes
PAIR_CASES = [
    ("NT", 256, 256, 64, "f32", 0.0, 0), ("NT", 512, 512, 512, "f32", 0.0, 0), ("NN", 512, 512, 512, "bf16", 0.0, 0),
    ("TN", 512, 512, 512, "f32", 0.0, 0), ("TT", 256, 512, 512, "f32", 0.0, 0), ("NT", 300, 300, 200, "bf16", 0.0, 0),
    ("TN", 200, 520, 136, "f32", 1.0, 0), ("NN", 1000, 264, 72, "bf16", 1.0, 0),
    ("NT", 4096, 4096, 4096, "bf16", 0.0, 30), ("NN", 4096, 4096, 4096, "bf16", 0.0, 30),
    ("TN", 4096, 4096, 4096, "f32", 0.0, 30), ("NT", 8192, 4096, 1024, "bf16", 0.0, 30),
    ("TN", 4096, 4096, 8192, "f32", 0.0, 30),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pairs":
        CASES = PAIR_CASES
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        form, M, N, K, cdt = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
        beta = float(sys.argv[7]) if len(sys.argv) > 7 else 0.0
        iters = int(sys.argv[8]) if len(sys.argv) > 8 else 0
        one(form, M, N, K, cdt, beta, iters)
        sys.exit(0)
    for c in CASES:
        t0 = time.time()
        cmd = [sys.executable, os.path.abspath(__file__), "one"] + [str(v) for v in c]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            out = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            if out:
                print(out[-1], flush=True)
            else:
                print(f"PROBE-FAIL {c} rc={r.returncode} stdout={r.stdout[-400:]!r} stderr={r.stderr[-600:]!r}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"PROBE-TIMEOUT {c}", flush=True)
