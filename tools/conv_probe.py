#!/usr/bin/env python
"""Hardware probe for the tcgen05 convolution engine (development tool): one case per process.
    python tools/conv_probe.py fwd N C H W COUT KH KW [iters]
    python tools/conv_probe.py dw  N C H W COUT KH KW [iters]
    python tools/conv_probe.py dx  N C H W COUT KH KW [iters]
    python tools/conv_probe.py bwd N C H W COUT KH KW [iters]     (dW + dbias + dX in one pass over G)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    mode = sys.argv[1]
    n, c, h, w, cout, kh, kw = (int(v) for v in sys.argv[2:9])
    iters = int(sys.argv[9]) if len(sys.argv) > 9 else 0
    import neuronika_b200 as nk
    import oracle as O
    from neuronika_b200 import ops
    dev = nk.Device(0)
    rng = np.random.default_rng(0)
    x = O.bf16_round(rng.uniform(0, 1, (n, c, h, w)).astype(np.float32))
    wt = O.bf16_round(rng.uniform(-0.3, 0.3, (cout, c, kh, kw)).astype(np.float32))
    ho, wo = h - kh + 1, w - kw + 1
    g = O.bf16_round(rng.uniform(-1, 1, (n, cout, ho, wo)).astype(np.float32))
    dx_, dw_, dg_ = dev.from_ndarray(x, nk.BF16), dev.from_ndarray(wt, nk.BF16), dev.from_ndarray(g, nk.BF16)
    check = n * cout * ho * wo * c * kh * kw < 4e9
    res = {"mode": mode, "shape": [n, c, h, w, cout, kh, kw]}
    if mode == "fwd":
        y = nk.CuArray(dev, (n, cout, ho, wo), nk.BF16)
        run = lambda: ops.conv2d(dx_, dw_, out=y)
        run()
        dev.synchronize()
        got = y.as_ndarray()
        bytes_alg = 2.0 * (x.size + got.size)
        if check:
            want = O.conv_forward(x, wt, (1, 1), (1, 1)).astype(np.float64)
    elif mode == "dw":
        out = nk.CuArray(dev, wt.shape, nk.F32)
        db = nk.CuArray(dev, (cout, 1, 1), nk.F32)
        run = lambda: ops.conv2d_bwd_kernel(out, dg_, dx_, beta=0.0, dbias=db)
        run()
        dev.synchronize()
        got = out.as_ndarray()
        bytes_alg = 2.0 * (x.size + g.size)
        if check:
            want = np.zeros_like(wt)
            O.conv_backward_kernel(want, g, x, (1, 1), (1, 1))
            want = want.astype(np.float64)
            res["dbias_err"] = float(np.abs(db.as_ndarray().ravel() - g.astype(np.float64).sum((0, 2, 3))).max())
    elif mode == "bwd":
        out = nk.CuArray(dev, x.shape, nk.BF16)
        odw = nk.CuArray(dev, wt.shape, nk.F32)
        db = nk.CuArray(dev, (cout, 1, 1), nk.F32)
        run = lambda: ops.conv2d_bwd(out, odw, dg_, dx_, dw_, beta_dx=0.0, beta_dw=0.0, dbias=db)
        run()
        dev.synchronize()
        got = out.as_ndarray()
        bytes_alg = 2.0 * (2 * x.size + g.size)
        if check:
            want = np.zeros_like(x)
            O.conv_backward_input(want, g, wt, (1, 1), (1, 1))
            want = want.astype(np.float64)
            wdw = np.zeros_like(wt)
            O.conv_backward_kernel(wdw, g, x, (1, 1), (1, 1))
            res["dw_err"] = float(np.abs(odw.as_ndarray() - wdw).max())
            res["dw_rms"] = float(np.sqrt((wdw.astype(np.float64) ** 2).mean()))
            res["dbias_err"] = float(np.abs(db.as_ndarray().ravel() - g.astype(np.float64).sum((0, 2, 3))).max())
    else:
        out = nk.CuArray(dev, x.shape, nk.BF16)
        run = lambda: ops.conv2d_bwd_input(out, dg_, dw_, beta=0.0)
        run()
        dev.synchronize()
        got = out.as_ndarray()
        bytes_alg = 2.0 * (x.size + g.size)
        if check:
            want = np.zeros_like(x)
            O.conv_backward_input(want, g, wt, (1, 1), (1, 1))
            want = want.astype(np.float64)
    res["kernel"] = dev.last_conv_kernel
    if check:
        err = np.abs(got - want)
        res["max_err"] = float(err.max())
        res["rms"] = float(np.sqrt((want ** 2).mean()))
        bad = np.argwhere(err > 0.02 * res["rms"] + 0.02 * np.abs(want))
        res["bad_frac"] = float(len(bad) / err.size)
        res["first_bad"] = bad[:5].tolist()
    if iters:
        for _ in range(2):
            run()
        dev.synchronize()
        dev.timer_start()
        for _ in range(iters):
            run()
        ms = dev.timer_stop() / iters
        res["ms"] = ms
        res["GBps_alg"] = bytes_alg / ms / 1e6
    print("CONV " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
