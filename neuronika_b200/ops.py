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
    _ck(lib.nk_nll_fwd(logp.device.ctx, out.ptr, logp.ptr, target.ptr, target.dtype, n, c, logp.dtype, int(mean)),
        logp.device)
    return out


def nll_bwd(dlogp, target, g, mean=True, beta=1.0):
    n, c = dlogp.shape
    _ck(lib.nk_nll_bwd(dlogp.device.ctx, dlogp.ptr, target.ptr, target.dtype, g.ptr, n, c, dlogp.dtype, int(mean),
                       float(beta)), dlogp.device)
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


def conv2d_bwd_uniform(dx: CuArray, dw: CuArray, g_value: float, x: CuArray, w: CuArray, stride=(1, 1), dilation=(1, 1),
                       groups=1, beta_dx=1.0, beta_dw=1.0, dbias: CuArray | None = None) -> bool:
    """ConvolutionBackward for an output gradient that is `g_value` everywhere (a deferred `backward(seed)` fill):
    the kernel synthesises G.  Returns False (nothing done) when the shape is outside the tensor-core engine."""
    dev = x.device
    rc = lib.nk_conv2d_bwd_uniform(dev.ctx, dx.ptr, float(beta_dx), dw.ptr, dw.dtype, dbias.ptr if dbias is not None else None,
                                   float(beta_dw), float(g_value), x.ptr, w.ptr,
                                   *_conv_args(x.shape, w.shape, stride, dilation, groups), x.dtype)
    if rc == -5:
        return False
    _ck(rc, dev)
    return True


# ---------------------------------------------------------------- sgd
def sgd_step(w: CuArray, g: CuArray, lr, l2=0.0, momentum=0.0, dampening=0.0, nesterov=False,
             buf: CuArray | None = None, master: CuArray | None = None, grad_scale=1.0, write_back_grad=True):
    dev = w.device
    _ck(lib.nk_sgd_step(dev.ctx, w.ptr, w.dtype, g.ptr, g.dtype, buf.ptr if buf is not None else None,
                        master.ptr if master is not None else None, w.size, float(lr), float(l2), float(momentum),
                        float(dampening), int(nesterov), float(grad_scale), int(write_back_grad)), dev)
    return w


# ---------------------------------------------------------------- 8-f: elementwise family
BIN = {"add": L.NK_BIN_ADD, "sub": L.NK_BIN_SUB, "mul": L.NK_BIN_MUL, "div": L.NK_BIN_DIV}
UN = {"neg": L.NK_UN_NEG, "exp": L.NK_UN_EXP, "ln": L.NK_UN_LN, "sqrt": L.NK_UN_SQRT, "sigmoid": L.NK_UN_SIGMOID,
      "tanh": L.NK_UN_TANH, "softplus": L.NK_UN_SOFTPLUS, "leaky_relu": L.NK_UN_LEAKY_RELU, "powi": L.NK_UN_POWI}
PAD = {"constant": L.NK_PAD_CONSTANT, "reflective": L.NK_PAD_REFLECTIVE, "replicative": L.NK_PAD_REPLICATIVE}


def binary(op: str, l: CuArray, r: CuArray, out: CuArray | None = None) -> CuArray:
    dev = l.device
    shape = cobroadcast(l.shape, r.shape)
    out = out or CuArray(dev, shape, l.dtype)
    _ck(lib.nk_binary_bcast_fwd(dev.ctx, BIN[op], out.ptr, l.ptr, r.ptr, l.dtype, len(shape), L.shape_arr(shape),
                                l.ndim, L.shape_arr(l.shape), r.ndim, L.shape_arr(r.shape)), dev)
    return out


def binary_bwd(op: str, side: int, dst: CuArray, g: CuArray, l: CuArray, r: CuArray, beta=1.0) -> CuArray:
    """dst (the gradient of operand `side`) = beta*dst + unbroadcast(factor(g, l, r))."""
    dev = g.device
    _ck(lib.nk_binary_bcast_bwd(dev.ctx, BIN[op], int(side), dst.ptr, dst.dtype, g.ptr, l.ptr, r.ptr, g.dtype,
                                l.ndim, L.shape_arr(l.shape), r.ndim, L.shape_arr(r.shape), float(beta)), dev)
    return dst


def unary(op: str, x: CuArray, iparam: int = 0, out: CuArray | None = None) -> CuArray:
    out = out or CuArray(x.device, x.shape, x.dtype)
    _ck(lib.nk_unary_fwd(x.device.ctx, UN[op], out.ptr, x.ptr, x.size, x.dtype, int(iparam)), x.device)
    return out


def unary_bwd(op: str, dx: CuArray, saved: CuArray | None, g: CuArray, iparam: int = 0, beta=1.0) -> CuArray:
    _ck(lib.nk_unary_bwd(g.device.ctx, UN[op], dx.ptr, saved.ptr if saved is not None else None, g.ptr, g.size,
                         g.dtype, int(iparam), float(beta)), g.device)
    return dx


def transpose(src: CuArray, out: CuArray | None = None, beta=0.0) -> CuArray:
    out = out or CuArray(src.device, tuple(reversed(src.shape)), src.dtype)
    _ck(lib.nk_transpose(src.device.ctx, out.ptr, out.dtype, src.ptr, src.dtype, src.ndim, L.shape_arr(src.shape),
                         float(beta)), src.device)
    return out


def pad_nd(x: CuArray, padding, mode="constant", value=0.0, out=None) -> CuArray:
    nsp = len(padding)
    lead, sp = x.shape[:x.ndim - nsp], x.shape[x.ndim - nsp:]
    out = out or CuArray(x.device, tuple(lead) + tuple(s + 2 * p for s, p in zip(sp, padding)), x.dtype)
    planes = int(np.prod(lead)) if lead else 1
    _ck(lib.nk_padnd_fwd(x.device.ctx, out.ptr, x.ptr, planes, nsp, L.shape_arr(sp), L.shape_arr(padding), PAD[mode],
                         float(value), x.dtype), x.device)
    return out


def pad_nd_bwd(dx: CuArray, g: CuArray, padding, beta=1.0) -> CuArray:
    nsp = len(padding)
    lead, sp = dx.shape[:dx.ndim - nsp], dx.shape[dx.ndim - nsp:]
    planes = int(np.prod(lead)) if lead else 1
    _ck(lib.nk_padnd_bwd(dx.device.ctx, dx.ptr, g.ptr, planes, nsp, L.shape_arr(sp), L.shape_arr(padding), dx.dtype,
                         float(beta)), dx.device)
    return dx


# ---------------------------------------------------------------- 8-f: mv / vm / vv
def gemv(a: CuArray, x: CuArray, y: CuArray | None = None, trans=False, beta=0.0) -> CuArray:
    rows, cols = a.shape
    y = y or CuArray(a.device, (cols if trans else rows,), a.dtype)
    _ck(lib.nk_gemv(a.device.ctx, int(trans), rows, cols, a.ptr, x.ptr, float(beta), y.ptr, a.dtype, y.dtype), a.device)
    return y


def outer_acc(a: CuArray, u: CuArray, v: CuArray, beta=1.0) -> CuArray:
    rows, cols = a.shape
    _ck(lib.nk_outer_acc(a.device.ctx, a.ptr, a.dtype, u.ptr, v.ptr, rows, cols, u.dtype, float(beta)), a.device)
    return a


def dot(a: CuArray, b: CuArray, out: CuArray | None = None) -> CuArray:
    out = out or CuArray(a.device, (), F32)
    _ck(lib.nk_dot(a.device.ctx, out.ptr, a.ptr, b.ptr, a.size, a.dtype), a.device)
    return out


def scale_acc(dst: CuArray, x: CuArray, scalar: CuArray, beta=1.0) -> CuArray:
    _ck(lib.nk_scale_acc(dst.device.ctx, dst.ptr, dst.dtype, x.ptr, x.dtype, scalar.ptr, x.size, float(beta)), dst.device)
    return dst


# ---------------------------------------------------------------- 8-f: 1-d / 3-d convolution
def _convnd_args(x_shape, w_shape, stride, dilation, groups):
    nsp = len(x_shape) - 2
    return [nsp, x_shape[0], x_shape[1], L.shape_arr(x_shape[2:]), w_shape[0], L.shape_arr(w_shape[2:]),
            L.shape_arr(stride), L.shape_arr(dilation), groups]


def convnd(x: CuArray, w: CuArray, stride, dilation, groups=1, out=None) -> CuArray:
    out = out or CuArray(x.device, conv_out_shape(x.shape, w.shape, stride, dilation), x.dtype)
    _ck(lib.nk_convnd_fwd(x.device.ctx, out.ptr, x.ptr, w.ptr, *_convnd_args(x.shape, w.shape, stride, dilation, groups),
                          x.dtype), x.device)
    return out


def convnd_bwd_input(dx: CuArray, g: CuArray, w: CuArray, stride, dilation, groups=1, beta=1.0) -> CuArray:
    _ck(lib.nk_convnd_bwd_input(g.device.ctx, dx.ptr, g.ptr, w.ptr,
                                *_convnd_args(dx.shape, w.shape, stride, dilation, groups), g.dtype, float(beta)), g.device)
    return dx


def convnd_bwd_kernel(dw: CuArray, g: CuArray, x: CuArray, stride, dilation, groups=1, beta=1.0) -> CuArray:
    _ck(lib.nk_convnd_bwd_kernel(g.device.ctx, dw.ptr, dw.dtype, g.ptr, x.ptr,
                                 *_convnd_args(x.shape, dw.shape, stride, dilation, groups), g.dtype, float(beta)), g.device)
    return dw
