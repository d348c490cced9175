"""Functional wrappers: one Python function per C-ABI operator entry point (include/nk_b200.h).

Each takes/returns CuArray and launches asynchronously on the device's stream.  `beta`
selects the reference's accumulate protocol on backward ops (1 = `+=`, the reference
behaviour; 0 = overwrite a buffer known to be zero)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .device import BF16, F32, CuArray, Device

lib = L.lib


def _ck(rc, dev: Device):
    L.check(rc, dev.ctx)


# ---------------------------------------------------------------- gemm
def gemm(a: CuArray, b: CuArray, c: CuArray, trans_a=False, trans_b=False, alpha=1.0, beta=0.0,
         bias: CuArray | None = None, relu=False) -> CuArray:
    """c = alpha*op(a).op(b) + beta*c (+bias, relu); shapes are the stored (row-major) shapes."""
    dev = a.device
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    kb, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if k != kb or tuple(c.shape) != (m, n):
        raise ValueError(f"gemm: shape mismatch op(a)=({m},{k}) op(b)=({kb},{n}) c={c.shape}")
    if a.dtype != b.dtype:
        raise ValueError("gemm: operand dtypes differ")
    _ck(lib.nk_gemm_bias_act(dev.ctx, int(trans_a), int(trans_b), m, n, k, float(alpha), a.ptr, a.shape[1],
                             b.ptr, b.shape[1], float(beta), c.ptr, n, a.dtype, c.dtype,
                             bias.ptr if bias is not None else None, bias.dtype if bias is not None else F32,
                             int(relu)), dev)
    return c


def mm(a, b, out=None, out_dtype=None):
    out = out or CuArray(a.device, (a.shape[0], b.shape[1]), out_dtype if out_dtype is not None else a.dtype)
    return gemm(a, b, out)


def mm_t(x, w, out=None, out_dtype=None, bias=None, relu=False):
    out = out or CuArray(x.device, (x.shape[0], w.shape[0]), out_dtype if out_dtype is not None else x.dtype)
    return gemm(x, w, out, trans_b=True, bias=bias, relu=relu)


# ---------------------------------------------------------------- broadcast add
def cobroadcast(ls, rs):
    """utils.rs:97-125"""
    big, small = (ls, rs) if len(ls) >= len(rs) else (rs, ls)
    out = list(big)
    off = len(big) - len(small)
    for i, r in enumerate(small):
        l = out[off + i]
        if l != r:
            if l == 1:
                out[off + i] = r
            elif r != 1:
                raise ValueError("The two tensors have incompatible shape.")
    return tuple(out)


def add(l: CuArray, r: CuArray, out: CuArray | None = None) -> CuArray:
    dev = l.device
    shape = cobroadcast(l.shape, r.shape)
    out = out or CuArray(dev, shape, l.dtype)
    _ck(lib.nk_add_bcast_fwd(dev.ctx, out.ptr, l.ptr, r.ptr, l.dtype, len(shape), L.shape_arr(shape),
                             l.ndim, L.shape_arr(l.shape), r.ndim, L.shape_arr(r.shape)), dev)
    return out


def unbroadcast_acc(dst: CuArray, g: CuArray, beta=1.0) -> CuArray:
    dev = g.device
    _ck(lib.nk_unbroadcast_acc(dev.ctx, dst.ptr, dst.dtype, dst.ndim, L.shape_arr(dst.shape), g.ptr, g.dtype,
                               g.ndim, L.shape_arr(g.shape), float(beta)), dev)
    return dst


# ---------------------------------------------------------------- relu / softmax
def relu(x: CuArray, out=None) -> CuArray:
    out = out or CuArray(x.device, x.shape, x.dtype)
    _ck(lib.nk_relu_fwd(x.device.ctx, out.ptr, x.ptr, x.size, x.dtype), x.device)
    return out


def relu_bwd(dx: CuArray, x: CuArray, g: CuArray, beta=1.0) -> CuArray:
    _ck(lib.nk_relu_bwd(x.device.ctx, dx.ptr, x.ptr, g.ptr, x.size, x.dtype, float(beta)), x.device)
    return dx


def _lanes(shape, axis):
    outer = int(np.prod(shape[:axis])) if axis > 0 else 1
    inner = int(np.prod(shape[axis + 1:])) if axis + 1 < len(shape) else 1
    return outer, int(shape[axis]), inner


def softmax(x: CuArray, axis: int, out=None, log=False) -> CuArray:
    out = out or CuArray(x.device, x.shape, x.dtype)
    o, n, i = _lanes(x.shape, axis)
    fn = lib.nk_log_softmax_fwd if log else lib.nk_softmax_fwd
    _ck(fn(x.device.ctx, out.ptr, x.ptr, o, n, i, x.dtype), x.device)
    return out


def softmax_bwd(dx: CuArray, y: CuArray, g: CuArray, axis: int, beta=1.0, log=False) -> CuArray:
    o, n, i = _lanes(y.shape, axis)
    fn = lib.nk_log_softmax_bwd if log else lib.nk_softmax_bwd
    _ck(fn(y.device.ctx, dx.ptr, y.ptr, g.ptr, o, n, i, y.dtype, float(beta)), y.device)
    return dx


# ---------------------------------------------------------------- losses / reductions
def mse(x: CuArray, t: CuArray, mean=True, out=None) -> CuArray:
    out = out or CuArray(x.device, (), F32)
    _ck(lib.nk_mse_fwd(x.device.ctx, out.ptr, x.ptr, t.ptr, x.size, x.dtype, int(mean)), x.device)
    return out


def mse_bwd(dx, x, t, g: CuArray, mean=True, beta=1.0):
    _ck(lib.nk_mse_bwd(x.device.ctx, dx.ptr, x.ptr, t.ptr, g.ptr, x.size, x.dtype, int(mean), float(beta)), x.device)
    return dx


def nll(logp: CuArray, target: CuArray, mean=True, out=None) -> CuArray:
    out = out or CuArray(logp.device, (), F32)
    n, c = logp.shape
    _ck(lib.nk_nll_fwd(logp.device.ctx, out.ptr, logp.ptr, target.ptr, n, c, logp.dtype, int(mean)), logp.device)
    return out


def nll_bwd(dlogp, target, g, mean=True, beta=1.0):
    n, c = dlogp.shape
    _ck(lib.nk_nll_bwd(dlogp.device.ctx, dlogp.ptr, target.ptr, g.ptr, n, c, dlogp.dtype, int(mean), float(beta)),
        dlogp.device)
    return dlogp


def reduce_sum(x: CuArray, mean=False, out=None) -> CuArray:
    out = out or CuArray(x.device, (), F32)
    _ck(lib.nk_sum_fwd(x.device.ctx, out.ptr, x.ptr, x.size, x.dtype, int(mean)), x.device)
    return out


def reduce_sum_bwd(dx: CuArray, g: CuArray, mean=False, beta=1.0):
    _ck(lib.nk_sum_bwd(dx.device.ctx, dx.ptr, g.ptr, dx.size, dx.dtype, int(mean), float(beta)), dx.device)
    return dx


# ---------------------------------------------------------------- pad / conv
def pad2d(x: CuArray, padding, value=0.0, out=None) -> CuArray:
    ph, pw = padding
    *lead, h, w = x.shape
    out = out or CuArray(x.device, tuple(lead) + (h + 2 * ph, w + 2 * pw), x.dtype)
    planes = int(np.prod(lead)) if lead else 1
    _ck(lib.nk_pad2d_fwd(x.device.ctx, out.ptr, x.ptr, planes, h, w, ph, pw, float(value), x.dtype), x.device)
    return out


def pad2d_bwd(dx: CuArray, g: CuArray, padding, beta=1.0):
    ph, pw = padding
    *lead, h, w = dx.shape
    planes = int(np.prod(lead)) if lead else 1
    _ck(lib.nk_pad2d_bwd(dx.device.ctx, dx.ptr, g.ptr, planes, h, w, ph, pw, dx.dtype, float(beta)), dx.device)
    return dx


def conv_out_shape(xs, ws, stride, dilation):
    """utils.rs:207-237"""
    out = [xs[0], ws[0]]
    for i, k, s, d in zip(xs[2:], ws[2:], stride, dilation):
        out.append((i - d * (k - 1) - 1) // s + 1)
    return tuple(out)


def _conv_args(x_shape, w_shape, stride, dilation, groups):
    n, cin, h, w = x_shape
    cout, _, kh, kw = w_shape
    return [n, cin, h, w, cout, kh, kw, stride[0], stride[1], dilation[0], dilation[1], groups]


def conv2d(x: CuArray, w: CuArray, stride=(1, 1), dilation=(1, 1), groups=1, bias=None, relu=False, out=None):
    dev = x.device
    if x.ndim != 4 or w.ndim != 4:
        raise ValueError(f"Invalid kernel shape {list(w.shape)} for 2d conv")
    out = out or CuArray(dev, conv_out_shape(x.shape, w.shape, stride, dilation), x.dtype)
    _ck(lib.nk_conv2d_fwd(dev.ctx, out.ptr, x.ptr, w.ptr, bias.ptr if bias is not None else None, int(relu),
                          *_conv_args(x.shape, w.shape, stride, dilation, groups), x.dtype), dev)
    return out


def conv2d_bwd_input(dx: CuArray, g: CuArray, w: CuArray, stride=(1, 1), dilation=(1, 1), groups=1, beta=1.0):
    dev = g.device
    _ck(lib.nk_conv2d_bwd_input(dev.ctx, dx.ptr, g.ptr, w.ptr,
                                *_conv_args(dx.shape, w.shape, stride, dilation, groups), g.dtype, float(beta)), dev)
    return dx


def conv2d_bwd_kernel(dw: CuArray, g: CuArray, x: CuArray, stride=(1, 1), dilation=(1, 1), groups=1, beta=1.0,
                      dbias: CuArray | None = None):
    dev = g.device
    _ck(lib.nk_conv2d_bwd_kernel(dev.ctx, dw.ptr, dw.dtype, dbias.ptr if dbias is not None else None, g.ptr, x.ptr,
                                 *_conv_args(x.shape, dw.shape, stride, dilation, groups), g.dtype, float(beta)), dev)
    return dw


def conv2d_bwd(dx: CuArray, dw: CuArray, g: CuArray, x: CuArray, w: CuArray, stride=(1, 1), dilation=(1, 1), groups=1,
               beta_dx=1.0, beta_dw=1.0, dbias: CuArray | None = None):
    """Both halves of ConvolutionBackward in one call (one pass over `g` on the tensor-core path)."""
    dev = g.device
    _ck(lib.nk_conv2d_bwd(dev.ctx, dx.ptr, float(beta_dx), dw.ptr, dw.dtype, dbias.ptr if dbias is not None else None,
                          float(beta_dw), g.ptr, x.ptr, w.ptr,
                          *_conv_args(x.shape, w.shape, stride, dilation, groups), g.dtype), dev)
    return dx, dw


# ---------------------------------------------------------------- sgd
def sgd_step(w: CuArray, g: CuArray, lr, l2=0.0, momentum=0.0, dampening=0.0, nesterov=False,
             buf: CuArray | None = None, master: CuArray | None = None, grad_scale=1.0, write_back_grad=True):
    dev = w.device
    _ck(lib.nk_sgd_step(dev.ctx, w.ptr, w.dtype, g.ptr, g.dtype, buf.ptr if buf is not None else None,
                        master.ptr if master is not None else None, w.size, float(lr), float(l2), float(momentum),
                        float(dampening), int(nesterov), float(grad_scale), int(write_back_grad)), dev)
    return w
