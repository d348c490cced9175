"""numpy restatement of the reference's Forward/Backward nodes (test oracle).

All arrays are C-order float32 unless stated; ``+=`` accumulates into the
caller-owned gradient exactly like the reference's ``Backward`` nodes
(SURVEY.md Appendix B).  Citations are relative to ``/root/reference``.
``nv`` below abbreviates ``neuronika-variable/src``.
"""
from __future__ import annotations

import itertools
from typing import Sequence, Tuple

import numpy as np
from numpy.lib.stride_tricks import as_strided

F32 = np.float32

__all__ = [
    "bf16_round", "mm_forward", "mm_backward", "mm_t_forward", "mm_t_backward",
    "cobroadcast", "add_forward", "unbroadcast", "add_backward",
    "relu_forward", "relu_backward", "softmax_forward", "softmax_backward",
    "log_softmax_forward", "log_softmax_backward", "mse_forward", "mse_backward",
    "nll_forward", "nll_backward", "sum_forward", "sum_backward", "mean_forward",
    "mean_backward", "conv_out_shape", "check_conv_args", "check_groups_args",
    "im2col", "flatten_kernel", "conv_forward", "conv_backward_kernel",
    "conv_backward_input", "pad_forward", "pad_backward", "sgd_step",
    "linear_forward", "linear_backward", "conv2d_layer_forward",
    "conv2d_layer_backward", "mlp_step", "uniform_init",
]


# --------------------------------------------------------------------------- helpers
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round float32 -> bfloat16 (round-to-nearest-even) and return as float32.

    Not reference behaviour (the reference is f32 only); used so that the
    oracle and the device consume the *same* bf16-rounded operands
    (tier-1 parity)."""
    x = np.ascontiguousarray(x, dtype=F32)
    bits = x.view(np.uint32).astype(np.uint64)
    bias = ((bits >> 16) & 1) + 0x7FFF
    out = ((bits + bias) & 0xFFFF0000).astype(np.uint32)
    return out.view(F32).reshape(x.shape)


def uniform_init(rng: np.random.Generator, shape, k: float) -> np.ndarray:
    """U(-k, k) like ``init::uniform`` (neuronika-nn/src/init.rs:177-183); the
    reference draws from an unseeded thread_rng so only the distribution is
    mirrored, with a seeded generator."""
    return rng.uniform(-k, k, size=shape).astype(F32)


# --------------------------------------------------------------------------- matmul
def mm_forward(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """C = A.B -- MatrixMatrixMul::forward, nv/node/matrix_matrix_mul/mod.rs:31-41
    (general_mat_mul(1, A, B, 0, C))."""
    return np.matmul(a.astype(F32), b.astype(F32)).astype(F32)


def mm_backward(a, b, g, da=None, db=None):
    """dA += G.B^T ; dB += A^T.G -- nv/node/matrix_matrix_mul/mod.rs:63-73, 95-105
    (general_mat_mul with beta = 1)."""
    if da is not None:
        da += np.matmul(g, b.T)
    if db is not None:
        db += np.matmul(a.T, g)
    return da, db


def mm_t_forward(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Y = X.W^T -- MatrixMatrixMulT::forward, nv/node/matrix_matrix_mul_t/mod.rs:31-41."""
    return np.matmul(x.astype(F32), w.astype(F32).T).astype(F32)


def mm_t_backward(x, w, g, dx=None, dw=None):
    """dX += G.W ; dW += G^T.X -- nv/node/matrix_matrix_mul_t/mod.rs:63-73, 95-105."""
    if dx is not None:
        dx += np.matmul(g, w)
    if dw is not None:
        dw += np.matmul(g.T, x)
    return dx, dw


# --------------------------------------------------------------------------- broadcast add
def cobroadcast(left: Sequence[int], right: Sequence[int]) -> Tuple[int, ...]:
    """Result shape of a broadcasting binary op -- nv/utils.rs:97-125 (right-aligned
    shapes, dims equal or 1, panics otherwise)."""
    bigger, smaller = (left, right) if len(left) >= len(right) else (right, left)
    out = list(bigger)
    off = len(bigger) - len(smaller)
    for i, r in enumerate(smaller):
        l = out[off + i]
        if l != r:
            if l == 1:
                out[off + i] = r
            elif r != 1:
                raise ValueError("The two tensors have incompatible shape.")
    return tuple(out)


def add_forward(l: np.ndarray, r: np.ndarray) -> np.ndarray:
    """Y = L + R with co-broadcast -- Addition::forward, nv/node/addition/mod.rs:39-50."""
    shape = cobroadcast(l.shape, r.shape)
    return (np.broadcast_to(l, shape) + np.broadcast_to(r, shape)).astype(F32)


def unbroadcast(g: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    """Reduce ``g`` back to ``shape`` (intent of utils::accumulate, nv/utils.rs:152-192:
    sum the leading extra axes, then keep-dim sum every axis whose target size is 1).
    The reference body is wrong for non-symmetric data (SURVEY.md 8-c defect 1); this is
    the intended maths and agrees with nv/node/addition/test.rs:91-202."""
    shape = tuple(shape)
    k = g.ndim - len(shape)
    out = g.astype(F32)
    if k > 0:
        out = out.sum(axis=tuple(range(k)), dtype=F32)
    for ax, s in enumerate(shape):
        if s == 1 and out.shape[ax] != 1:
            out = out.sum(axis=ax, keepdims=True, dtype=F32)
    return out.reshape(shape).astype(F32)


def add_backward(g, dl=None, dr=None):
    """dL += unbroadcast(G) ; dR += unbroadcast(G) -- nv/node/addition/mod.rs:81-92,124-135."""
    if dl is not None:
        dl += unbroadcast(g, dl.shape)
    if dr is not None:
        dr += unbroadcast(g, dr.shape)
    return dl, dr


# --------------------------------------------------------------------------- relu
def relu_forward(x: np.ndarray) -> np.ndarray:
    """y = max(x, 0) -- ReLU::forward, nv/node/relu/mod.rs:29-38 (f32::max: NaN -> 0)."""
    return np.where(x > 0, x, F32(0)).astype(F32)


def relu_backward(x, g, dx):
    """dx += (x > 0) * g -- ReLUBackward, nv/node/relu/mod.rs:67-79 (uses the *input* x)."""
    dx += np.where(x > 0, g, F32(0)).astype(F32)
    return dx


# --------------------------------------------------------------------------- softmax family
def softmax_forward(x: np.ndarray, axis: int) -> np.ndarray:
    """Per lane: m = max, e = exp(x - m), y = e / sum(e) -- nv/node/softmax/mod.rs:37-53."""
    m = x.max(axis=axis, keepdims=True)
    e = np.exp((x - m).astype(F32)).astype(F32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def softmax_backward(y, g, dx, axis: int):
    """dx += y * (g - sum(g*y)) -- nv/node/softmax/mod.rs:84-104."""
    s = (g * y).sum(axis=axis, keepdims=True, dtype=F32)
    dx += (y * (g - s)).astype(F32)
    return dx


def log_softmax_forward(x: np.ndarray, axis: int) -> np.ndarray:
    """y = x - ln(sum(exp(x - m))) - m -- nv/node/logsoftmax/mod.rs:37-53."""
    m = x.max(axis=axis, keepdims=True)
    lse = np.log(np.exp((x - m).astype(F32)).sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)
    return (x - lse - m).astype(F32)


def log_softmax_backward(y, g, dx, axis: int):
    """dx += g - exp(y) * sum(g) -- nv/node/logsoftmax/mod.rs:84-102."""
    s = g.sum(axis=axis, keepdims=True, dtype=F32)
    dx += (g - np.exp(y).astype(F32) * s).astype(F32)
    return dx


# --------------------------------------------------------------------------- losses / reductions
def mse_forward(x, t, reduction: str = "mean") -> np.ndarray:
    """sum((x - t)^2) [/ numel] -- SquaredError::forward, nv/node/squared_error/mod.rs:46-58."""
    tot = ((x.astype(np.float64) - t) ** 2).sum()
    if reduction == "mean":
        tot = tot / x.size
    return np.asarray(tot, dtype=F32)


def mse_backward(x, t, g, dx, reduction: str = "mean"):
    """dx += 2 (x - t) g [/ numel] -- nv/node/squared_error/mod.rs:98-122."""
    g = F32(np.asarray(g).reshape(()))
    if reduction == "mean":
        dx += (F32(2.0) * (x - t) * g / F32(x.size)).astype(F32)
    else:
        dx += (F32(2.0) * (x - t) * g).astype(F32)
    return dx


def nll_forward(logp, target, reduction: str = "mean") -> np.ndarray:
    """-sum_n logp[n, t_n] [/ N]; targets are f32 class ids cast ``as usize``.
    Intended semantics per nv/node/nll/test.rs:8-118; the body at
    nv/node/nll/mod.rs:42-68 only has valid shapes when N == C (SURVEY.md 8-c defect 4)."""
    idx = np.asarray(target).astype(np.int64)
    n = logp.shape[0]
    tot = -logp.astype(np.float64)[np.arange(n), idx].sum()
    if reduction == "mean":
        tot = tot / n
    return np.asarray(tot, dtype=F32)


def nll_backward(target, g, dlogp, reduction: str = "mean"):
    """dlogp[n, t_n] -= g [/ N] -- nv/node/nll/mod.rs:100-133 (by intent)."""
    idx = np.asarray(target).astype(np.int64)
    n = dlogp.shape[0]
    g = F32(np.asarray(g).reshape(()))
    val = g / F32(n) if reduction == "mean" else g
    dlogp[np.arange(n), idx] -= val
    return dlogp


def sum_forward(x):
    """nv/node/sum/mod.rs:32-34."""
    return np.asarray(x.astype(np.float64).sum(), dtype=F32)


def sum_backward(g, dx):
    """dx += g -- nv/node/sum/mod.rs:64-66."""
    dx += F32(np.asarray(g).reshape(()))
    return dx


def mean_forward(x):
    """nv/node/mean/mod.rs:32-34."""
    return np.asarray(x.astype(np.float64).mean(), dtype=F32)


def mean_backward(g, dx):
    """dx += g / numel -- nv/node/mean/mod.rs:64-71."""
    dx += F32(np.asarray(g).reshape(())) / F32(dx.size)
    return dx


# --------------------------------------------------------------------------- convolution
def conv_out_shape(input_shape, kernel_shape, stride, dilation) -> Tuple[int, ...]:
    """(in - dil*(k-1) - 1)/stride + 1 per spatial axis, input already padded --
    nv/utils.rs:207-237."""
    out = [input_shape[0], kernel_shape[0]]
    for i, k, s, d in zip(input_shape[2:], kernel_shape[2:], stride, dilation):
        out.append((i - d * (k - 1) - 1) // s + 1)
    return tuple(out)


def check_conv_args(input_shape, kernel_shape, stride, dilation) -> None:
    """Same predicates (and messages) as nv/utils.rs:427-474."""
    nd = len(input_shape) - 2
    if nd != len(stride):
        raise ValueError(f"Invalid stride {list(stride)} for {nd}d conv.")
    if nd != len(dilation):
        raise ValueError(f"Invalid dilation {list(dilation)} for {nd}d conv.")
    if len(kernel_shape) != len(input_shape):
        raise ValueError(f"Invalid kernel shape {list(kernel_shape)} for {nd}d conv")
    for i, k, d in zip(input_shape[2:], kernel_shape[2:], dilation):
        if i < (k - 1) * d + 1:
            raise ValueError("The kernel size can't be greater than actual input size.")


def check_groups_args(input_shape, kernel_shape, groups: int) -> None:
    """nv/utils.rs:481-496."""
    if input_shape[1] % groups != 0:
        raise ValueError(f"In channels {input_shape[1]} is not divisible by groups {groups}")
    if kernel_shape[0] % groups != 0:
        raise ValueError(f"Out channels {kernel_shape[0]} is not divisible by groups {groups}")


def _windows(x: np.ndarray, kernel_shape, stride, dilation) -> np.ndarray:
    """Rolling-window view (N, 1, out.., Cin, k..) -- as_windows, nv/utils.rs:249-353."""
    x = np.ascontiguousarray(x)
    out = conv_out_shape(x.shape, kernel_shape, stride, dilation)
    shape = (out[0], 1) + tuple(out[2:]) + tuple(kernel_shape[1:])
    es = x.strides
    idx_strides = (es[0], es[1]) + tuple(e * s for e, s in zip(es[2:], stride))
    win_strides = (es[1],) + tuple(e * d for e, d in zip(es[2:], dilation))
    return as_strided(x, shape=shape, strides=idx_strides + win_strides, writeable=False)


def im2col(x: np.ndarray, kernel_shape, stride, dilation) -> np.ndarray:
    """Materialised columns (N, L, K): K ordered (c, i, j), L ordered (p, q) --
    ``as_windows(..).to_shape(columns_shape(..))``, nv/node/convolution/mod.rs:104-108,
    nv/utils.rs:403-420; layout pinned by nv/node/convolution/test.rs:11-84."""
    w = _windows(x, kernel_shape, stride, dilation)
    n = w.shape[0]
    k = int(np.prod(kernel_shape[1:]))
    return np.ascontiguousarray(w).reshape(n, -1, k)


def flatten_kernel(kernel: np.ndarray) -> np.ndarray:
    """(Cout, Cin*k..) -- flat_shape, nv/node/convolution/mod.rs:50-58."""
    return kernel.reshape(kernel.shape[0], -1)


def _conv_forward_g1(x, kernel, stride, dilation) -> np.ndarray:
    """Per sample out[n] (Cout x L) = Wflat (Cout x K) . cols[n]^T (K x L) --
    nv/node/convolution/mod.rs:85-123."""
    out_shape = conv_out_shape(x.shape, kernel.shape, stride, dilation)
    cols = im2col(x, kernel.shape, stride, dilation)            # (N, L, K)
    wf = flatten_kernel(kernel)                                 # (Cout, K)
    y = np.matmul(wf[None, :, :], cols.transpose(0, 2, 1))       # (N, Cout, L)
    return y.reshape(out_shape).astype(F32)


def conv_forward(x, kernel, stride, dilation, groups: int = 1) -> np.ndarray:
    """convolution()/grouped_convolution() -- nv/node/convolution/mod.rs:85-144.
    Cross-correlation, no padding, beta = 0."""
    check_conv_args(x.shape, kernel.shape, stride, dilation)
    if groups < 2:
        return _conv_forward_g1(x, kernel, stride, dilation)
    check_groups_args(x.shape, kernel.shape, groups)
    cin_g = x.shape[1] // groups
    cout_g = kernel.shape[0] // groups
    outs = []
    for gidx in range(groups):
        outs.append(_conv_forward_g1(x[:, gidx * cin_g:(gidx + 1) * cin_g],
                                     kernel[gidx * cout_g:(gidx + 1) * cout_g], stride, dilation))
    return np.concatenate(outs, axis=1).astype(F32)


def _conv_backward_kernel_g1(dk, g, x, stride, dilation):
    """dW[o, :] += G[:, o].flat (1 x N.L) . cols (N.L x K), beta = 1 --
    nv/node/convolution/mod.rs:191-226."""
    cols = im2col(x, dk.shape, stride, dilation)               # (N, L, K)
    nl = cols.shape[0] * cols.shape[1]
    mat = cols.reshape(nl, -1)
    gf = np.moveaxis(g, 1, 0).reshape(g.shape[1], nl)           # (Cout, N.L)
    dk += np.matmul(gf, mat).reshape(dk.shape).astype(F32)
    return dk


def conv_backward_kernel(dk, g, x, stride, dilation, groups: int = 1):
    """convolution_backward_kernel / grouped -- nv/node/convolution/mod.rs:191-226, 276-294."""
    if groups < 2:
        return _conv_backward_kernel_g1(dk, g, x, stride, dilation)
    cin_g = x.shape[1] // groups
    cout_g = dk.shape[0] // groups
    for gi in range(groups):
        sub = dk[gi * cout_g:(gi + 1) * cout_g]
        _conv_backward_kernel_g1(sub, g[:, gi * cout_g:(gi + 1) * cout_g],
                                 x[:, gi * cin_g:(gi + 1) * cin_g], stride, dilation)
    return dk


def _conv_backward_input_g1(dx, g, kernel, stride, dilation):
    """dX += col2im(G[n]^T . Wflat) -- intent of nv/node/convolution/mod.rs:146-189 with
    assign_from_cols :63-83.  The reference computes the transposed (K x L) buffer and then
    reinterprets it as (L x K) (SURVEY.md 8-c defect 2); with the all-ones kernels and
    gradients of nv/node/convolution/test.rs both readings give the same numbers."""
    n = g.shape[0]
    wf = flatten_kernel(kernel)                                   # (Cout, K)
    gl = g.reshape(n, g.shape[1], -1)                             # (N, Cout, L)
    cols = np.matmul(gl.transpose(0, 2, 1), wf[None])             # (N, L, K)
    out_sp = g.shape[2:]
    ksp = kernel.shape[2:]
    cin = kernel.shape[1]
    cols = cols.reshape((n,) + tuple(out_sp) + (cin,) + tuple(ksp))
    nd = len(out_sp)
    for koff in itertools.product(*[range(k) for k in ksp]):
        sl = [slice(None), slice(None)]
        for ax in range(nd):
            start = koff[ax] * dilation[ax]
            stop = start + (out_sp[ax] - 1) * stride[ax] + 1
            sl.append(slice(start, stop, stride[ax]))
        contrib = cols[(slice(None),) + (slice(None),) * nd + (slice(None),) + koff]
        # contrib: (N, out.., Cin) -> (N, Cin, out..)
        dx[tuple(sl)] += np.moveaxis(contrib, -1, 1)
    return dx


def conv_backward_input(dx, g, kernel, stride, dilation, groups: int = 1):
    """convolution_backward_input / grouped -- nv/node/convolution/mod.rs:146-189, 256-274."""
    if groups < 2:
        return _conv_backward_input_g1(dx, g, kernel, stride, dilation)
    cin_g = dx.shape[1] // groups
    cout_g = kernel.shape[0] // groups
    for gi in range(groups):
        sub = dx[:, gi * cin_g:(gi + 1) * cin_g]
        _conv_backward_input_g1(sub, g[:, gi * cout_g:(gi + 1) * cout_g],
                                kernel[gi * cout_g:(gi + 1) * cout_g], stride, dilation)
    return dx


# --------------------------------------------------------------------------- pad
def pad_forward(x: np.ndarray, padding: Sequence[int], value: float = 0.0) -> np.ndarray:
    """Constant / Zero padding of the spatial axes: output filled with ``value``, interior = x
    (bit-exact copy) -- nv/node/pad/mod.rs:97-129, pad/constant/mod.rs:14-39, pad/zero/mod.rs:14-22."""
    pads = [(0, 0)] * (x.ndim - len(padding)) + [(p, p) for p in padding]
    return np.pad(x, pads, mode="constant", constant_values=F32(value)).astype(F32)


def pad_backward(g: np.ndarray, dx: np.ndarray, padding: Sequence[int]):
    """dx += g[interior] -- PadBackward, nv/node/pad/mod.rs:157-182."""
    sl = [slice(None)] * (g.ndim - len(padding))
    for p, n in zip(padding, dx.shape[-len(padding):]):
        sl.append(slice(p, p + n))
    dx += g[tuple(sl)]
    return dx


# --------------------------------------------------------------------------- SGD
def sgd_step(w, g, lr: float, l2: float = 0.0, momentum: float | None = None,
             dampening: float | None = None, nesterov: bool = False, buf=None):
    """SGDParam::optimize -- neuronika-optim/src/sgd/mod.rs:191-231 with L2::penalize
    (penalty.rs:63-67).  Mutates w, g (g += 2*l2*w first, as the reference does) and
    returns the momentum buffer (None when momentum is off / <= f32::EPSILON)."""
    lr = F32(lr)
    g += (F32(2.0) * F32(l2) * w).astype(F32)
    if momentum is None or momentum <= np.finfo(F32).eps:
        w -= (g * lr).astype(F32)
        return None
    mu = F32(momentum)
    damp = F32(dampening or 0.0)
    if buf is None:
        buf = np.zeros_like(g)
    buf[...] = buf * mu + g * (F32(1.0) - damp)
    if nesterov:
        w -= ((g + buf * mu) * lr).astype(F32)
    else:
        w -= (buf * lr).astype(F32)
    return buf


# --------------------------------------------------------------------------- layers
def linear_forward(x, w, b):
    """Linear::forward = input.mm_t(W) + b -- neuronika-nn/src/lib.rs:441-447."""
    return add_forward(mm_t_forward(x, w), b)


def linear_backward(x, w, g, dx=None, dw=None, db=None):
    """Backward of mm_t + broadcast add (nodes above)."""
    if db is not None:
        db += unbroadcast(g, db.shape)
    mm_t_backward(x, w, g, dx, dw)
    return dx, dw, db


def conv2d_layer_forward(x, w, b, padding=(0, 0), stride=(1, 1), dilation=(1, 1), pad_value=0.0):
    """Intended Conv2d::forward (reference body is todo!(), neuronika-nn/src/lib.rs:809-814;
    documented intent :789-808 and bias shape (Cout,1,1) :774): pad -> convolution -> + bias."""
    xp = pad_forward(x, padding, pad_value) if any(padding) else x
    y = conv_forward(xp, w, stride, dilation, 1)
    return add_forward(y, b) if b is not None else y


def conv2d_layer_backward(x, w, g, dx=None, dw=None, db=None, padding=(0, 0), stride=(1, 1),
                          dilation=(1, 1), pad_value=0.0):
    xp = pad_forward(x, padding, pad_value) if any(padding) else x
    if db is not None:
        db += unbroadcast(g, db.shape)
    if dw is not None:
        conv_backward_kernel(dw, g, xp, stride, dilation)
    if dx is not None:
        if any(padding):
            dxp = np.zeros_like(xp)
            conv_backward_input(dxp, g, w, stride, dilation)
            pad_backward(dxp, dx, padding)
        else:
            conv_backward_input(dx, g, w, stride, dilation)
    return dx, dw, db


def mlp_step(x, t, params, lr: float, l2: float = 0.0, final: str = "softmax"):
    """One training step of config 4: relu(L1) -> relu(L2) -> softmax(L3),
    MSE(mean) loss, backward(1.0), SGD (no momentum).  ``params`` = [(W1,b1),(W2,b2),(W3,b3)],
    updated in place.  Returns (loss, grads) with grads = [(dW, db), ...] before the update."""
    acts = [x]
    pre = []
    h = x
    for li, (w, b) in enumerate(params):
        z = linear_forward(h, w, b)
        pre.append(z)
        if li < len(params) - 1:
            h = relu_forward(z)
        else:
            h = softmax_forward(z, 1) if final == "softmax" else log_softmax_forward(z, 1)
        acts.append(h)
    p = acts[-1]
    loss = mse_forward(p, t, "mean")
    dp = np.zeros_like(p)
    mse_backward(p, t, F32(1.0), dp, "mean")
    dz = np.zeros_like(p)
    if final == "softmax":
        softmax_backward(p, dp, dz, 1)
    else:
        log_softmax_backward(p, dp, dz, 1)
    grads = [None] * len(params)
    for li in reversed(range(len(params))):
        w, b = params[li]
        dw = np.zeros_like(w)
        db = np.zeros_like(b)
        dh = np.zeros_like(acts[li]) if li > 0 else None
        linear_backward(acts[li], w, dz, dh, dw, db)
        grads[li] = (dw, db)
        if li > 0:
            dz = np.zeros_like(pre[li - 1])
            relu_backward(pre[li - 1], dh, dz)
    for (w, b), (dw, db) in zip(params, grads):
        sgd_step(w, dw.copy(), lr, l2)
        sgd_step(b, db.copy(), lr, l2)
    return loss, grads
