"""numpy restatement of the SURVEY.md 8-f nodes (test oracle, same rules as nodes.py): the rest of the broadcasting /
elementwise / shape family, mv / vm / vv, the Adam-family optimizers.  f32 arithmetic, ``+=`` accumulates into the
caller-owned gradient like the reference's ``Backward`` nodes.  ``nv`` = ``neuronika-variable/src``,
``no`` = ``neuronika-optim/src`` (relative to ``/root/reference``).

Pinning: tests/test_oracle_next.py replays the literal vectors of the reference's own node tests
(tests/golden/tensors_next.json, lifted by tests/golden/make_goldens.py), the closed forms its enabled tests state
(exp/test.rs:24-36, logn/test.rs, multiplication/test.rs ...) and torch-CPU autograd / torch.optim on random data.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .nodes import cobroadcast, unbroadcast

F32 = np.float32

__all__ = [
    "binary_forward", "binary_backward", "unary_forward", "unary_backward", "UNARY_SAVES_OUTPUT",
    "transpose_forward", "transpose_backward", "pad_mode_forward", "pad_mode_backward",
    "mv_forward", "mv_backward", "vm_forward", "vm_backward", "vv_forward", "vv_backward",
    "penalize", "adam_step", "rmsprop_step", "adagrad_step",
]


# --------------------------------------------------------------------------- binary, broadcasting
def binary_forward(op: str, l: np.ndarray, r: np.ndarray) -> np.ndarray:
    """Zip::and_broadcast over both operands: nv/node/subtraction/mod.rs:44-49, multiplication/mod.rs:44-49,
    division/mod.rs:44-49 (addition/mod.rs:39-50)."""
    cobroadcast(l.shape, r.shape)
    l, r = l.astype(F32), r.astype(F32)
    return {"add": l + r, "sub": l - r, "mul": l * r, "div": l / r}[op].astype(F32)


def binary_backward(op: str, g, l, r, dl=None, dr=None):
    """buffer = factor(g, l, r) over the broadcast shape, then un-broadcast into the operand gradient
    (intended `accumulate`, see nodes.unbroadcast):
      sub  dL += g, dR += -g                       nv/node/subtraction/mod.rs:87-92, 130-135
      mul  dL += g*r, dR += g*l                    nv/node/multiplication/mod.rs:90-101, 139-148
      div  dL += g/r, dR += -g*l / r.powi(2)       nv/node/division/mod.rs:90-100, 142-151"""
    g, l, r = g.astype(F32), l.astype(F32), r.astype(F32)
    if op == "add":
        fl, fr = g, g
    elif op == "sub":
        fl, fr = g, -g
    elif op == "mul":
        fl, fr = g * r, g * l
    else:
        fl, fr = g / r, (-g * l) / (r * r)
    if dl is not None:
        dl += unbroadcast(np.broadcast_to(fl, g.shape).astype(F32), dl.shape)
    if dr is not None:
        dr += unbroadcast(np.broadcast_to(fr, g.shape).astype(F32), dr.shape)
    return dl, dr


# --------------------------------------------------------------------------- unary
UNARY_SAVES_OUTPUT = {"exp", "sqrt", "sigmoid", "tanh"}   # the Backward node keeps `data`; the others keep `operand_data`


def _powi(x, e: int):
    return np.power(x.astype(F32), F32(e)).astype(F32) if e >= 0 else (F32(1) / np.power(x.astype(F32), F32(-e))).astype(F32)


def unary_forward(op: str, x: np.ndarray, iparam: int = 0) -> np.ndarray:
    """nv/node/negation/mod.rs:35, exp/mod.rs:35, logn/mod.rs:35, sqrt/mod.rs:35, sigmoid/mod.rs:35 (1/(1+e^-x)),
    tanh/mod.rs:35, softplus/mod.rs:35 ((1+e^x).ln()), leaky_relu/mod.rs:36-38 (slope 0.01), power/mod.rs:44 (powi)."""
    x = x.astype(F32)
    with np.errstate(all="ignore"):
        if op == "neg":
            y = -x
        elif op == "exp":
            y = np.exp(x)
        elif op == "ln":
            y = np.log(x)
        elif op == "sqrt":
            y = np.sqrt(x)
        elif op == "sigmoid":
            y = F32(1) / (F32(1) + np.exp(-x))
        elif op == "tanh":
            y = np.tanh(x)
        elif op == "softplus":
            y = np.log(F32(1) + np.exp(x))
        elif op == "leaky_relu":
            y = np.where(x > 0, x, F32(0.01) * x)
        elif op == "powi":
            y = _powi(x, iparam)
        else:
            raise ValueError(op)
    return y.astype(F32)


def unary_backward(op: str, g, saved, dx, iparam: int = 0):
    """dx += g * f'(saved); `saved` is the node's output for UNARY_SAVES_OUTPUT, else its input:
    negation/mod.rs:67 (-= g), exp/mod.rs:73 (g*y), logn/mod.rs:73 (g/x), sqrt/mod.rs:73 (g/(2y)),
    sigmoid/mod.rs:74 (g*y*(1-y)), tanh/mod.rs:74 (g*(1-y^2)), softplus/mod.rs:74 (g/(1+e^-x)),
    power/mod.rs:86 (g * x.powi(e-1) * e).  leaky_relu/mod.rs:78-79 adds the constant 0.01 instead of 0.01*g on the
    negative side -- a defect its own test (g = 1) cannot see; the intended slope*g is restated."""
    g = g.astype(F32)
    s = saved.astype(F32) if saved is not None else None
    with np.errstate(all="ignore"):
        if op == "neg":
            d = -g
        elif op == "exp":
            d = g * s
        elif op == "ln":
            d = g / s
        elif op == "sqrt":
            d = g / (s * F32(2))
        elif op == "sigmoid":
            d = g * s * (F32(1) - s)
        elif op == "tanh":
            d = g * (F32(1) - s * s)
        elif op == "softplus":
            d = g / (F32(1) + np.exp(-s))
        elif op == "leaky_relu":
            d = np.where(s > 0, g, F32(0.01) * g)
        elif op == "powi":
            d = g * _powi(s, iparam - 1) * F32(iparam)
        else:
            raise ValueError(op)
    dx += d.astype(F32)
    return dx


# --------------------------------------------------------------------------- transpose
def transpose_forward(x: np.ndarray) -> np.ndarray:
    """`.t()` reverses every axis -- nv/node/transpose/mod.rs:32-36.  Bit exact."""
    return np.ascontiguousarray(x.T)


def transpose_backward(g: np.ndarray, dx: np.ndarray):
    """dX += G^T -- nv/node/transpose/mod.rs:66-68."""
    dx += g.T
    return dx


# --------------------------------------------------------------------------- padding with a mode
def _src_index(o: np.ndarray, length: int, p: int, mode: str) -> np.ndarray:
    """Source coordinate of padded coordinate o (pad/reflective/mod.rs:22-31, pad/replicative/mod.rs:22-31);
    -1 = fill value (constant mode)."""
    inside = (o >= p) & (o < length + p)
    if mode == "reflective":
        src = np.where(o < p, 2 * p - o, 2 * (length + p - 1) - o) - p
    elif mode == "replicative":
        src = np.where(o < p, 0, length - 1)
    else:
        src = np.full_like(o, -1)
    return np.where(inside, o - p, src)


def pad_mode_forward(x: np.ndarray, padding: Sequence[int], mode: str = "constant", value: float = 0.0) -> np.ndarray:
    """Pad::forward over the trailing len(padding) sample dims of (N, C, ...) -- nv/node/pad/mod.rs:97-129 with the
    modes of pad/{constant,zero,reflective,replicative}/mod.rs.  A bit-exact copy."""
    nsp = len(padding)
    out = x
    for k, p in enumerate(padding):
        ax = x.ndim - nsp + k
        length = x.shape[ax]
        o = np.arange(length + 2 * p)
        src = _src_index(o, length, p, mode)
        taken = np.take(out, np.clip(src, 0, max(length - 1, 0)), axis=ax)
        if mode == "constant":
            shape = [1] * out.ndim
            shape[ax] = -1
            taken = np.where((src < 0).reshape(shape), np.asarray(value, dtype=x.dtype), taken)
        out = taken
    return np.ascontiguousarray(out)


def pad_mode_backward(g: np.ndarray, dx: np.ndarray, padding: Sequence[int]):
    """dx += g[interior] whatever the mode -- nv/node/pad/mod.rs:157-182."""
    nsp = len(padding)
    sl = [slice(None)] * (g.ndim - nsp) + [slice(p, g.shape[g.ndim - nsp + k] - p) for k, p in enumerate(padding)]
    dx += g[tuple(sl)]
    return dx


# --------------------------------------------------------------------------- mv / vm / vv
def mv_forward(a, v):
    """y = A.v -- nv/node/matrix_vector_mul/mod.rs:32-40 (general_mat_vec_mul(1, A, v, 0, y))."""
    return (a.astype(F32) @ v.astype(F32)).astype(F32)


def mv_backward(a, v, g, da=None, dv=None):
    """dA += g (x) v (:64-69); dv += A^T.g (:93-101)."""
    if da is not None:
        da += np.outer(g, v).astype(F32)
    if dv is not None:
        dv += (a.T @ g).astype(F32)
    return da, dv


def vm_forward(v, a):
    """y = v.A -- nv/node/vector_matrix_mul/mod.rs:32-40 (general_mat_vec_mul(1, A^T, v, 0, y))."""
    return (v.astype(F32) @ a.astype(F32)).astype(F32)


def vm_backward(v, a, g, dv=None, da=None):
    """dv += A.g (:64-72); dA += v (x) g (:96-101)."""
    if dv is not None:
        dv += (a @ g).astype(F32)
    if da is not None:
        da += np.outer(v, g).astype(F32)
    return dv, da


def vv_forward(l, r):
    """s = <l, r> as a 0-d tensor -- nv/node/vector_vector_mul/mod.rs:32-34."""
    return np.asarray(np.dot(l.astype(np.float64), r.astype(np.float64)), dtype=F32)


def vv_backward(l, r, g, dl=None, dr=None):
    """dl += r*g ; dr += l*g (g a 0-d tensor) -- nv/node/vector_vector_mul/mod.rs:58-63."""
    gv = F32(np.asarray(g).reshape(()))
    if dl is not None:
        dl += r * gv
    if dr is not None:
        dr += l * gv
    return dl, dr


# --------------------------------------------------------------------------- optimizers
def penalize(w: np.ndarray, l1: float = 0.0, l2: float = 0.0) -> np.ndarray:
    """L1: lambda*signum(w); L2: 2*lambda*w; ElasticNet: both -- no/penalty.rs:63-79 (f32::signum(+-0) = +-1)."""
    sign = np.where(np.signbit(w), F32(-1), F32(1)).astype(F32)
    return (F32(l1) * sign + F32(2) * F32(l2) * w).astype(F32)


def adam_step(w, g, exp_avg, exp_avg_sq, step: int, lr, beta1, beta2, eps, l1=0.0, l2=0.0, max_exp_avg_sq=None):
    """One AdamParam::optimize (no/adam/mod.rs:131-169); with max_exp_avg_sq: AMSGradParam::optimize
    (no/amsgrad/mod.rs:159-204).  All arrays are updated in place; `step` is the value AFTER `self.step += 1`."""
    lr, beta1, beta2, eps = F32(lr), F32(beta1), F32(beta2), F32(eps)
    bc1 = F32(1) - F32(beta1 ** F32(step))
    bc2 = F32(1) - F32(beta2 ** F32(step))
    g += penalize(w, l1, l2)
    exp_avg[...] = exp_avg * beta1 + g * (F32(1) - beta1)
    exp_avg_sq[...] = exp_avg_sq * beta2 + g * g * (F32(1) - beta2)
    v = exp_avg_sq
    if max_exp_avg_sq is not None:
        np.maximum(max_exp_avg_sq, exp_avg_sq, out=max_exp_avg_sq)
        v = max_exp_avg_sq
    w -= exp_avg / ((np.sqrt(v) / np.sqrt(bc2)) + eps) * (lr / bc1)
    return w


def rmsprop_step(w, g, square_avg, lr, alpha, eps, momentum=None, centered=False, grad_avg=None, buffer=None,
                 l1=0.0, l2=0.0):
    """RMSPropParam::optimize, the four (centered, momentum) branches -- no/rmsprop/mod.rs:193-300."""
    lr, alpha, eps = F32(lr), F32(alpha if alpha is not None else 0.0), F32(eps)
    g += penalize(w, l1, l2)
    square_avg[...] = square_avg * alpha + g * g * (F32(1) - alpha)
    use_mom = momentum is not None and momentum > np.finfo(F32).eps
    if centered:
        grad_avg[...] = grad_avg * alpha + g * (F32(1) - alpha)
        denom = np.sqrt(square_avg + (-grad_avg * grad_avg)) + eps
    else:
        denom = np.sqrt(square_avg) + eps
    if use_mom:
        buffer[...] = buffer * F32(momentum) + g / denom
        w -= buffer * lr
    else:
        w -= g / denom * lr
    return w


def adagrad_step(w, g, grad_sq, step: int, lr, lr_decay, eps, l1=0.0, l2=0.0):
    """AdagradParam::optimize -- no/adagrad/mod.rs:113-140; `step` after the increment."""
    clr = F32(lr) / (F32(1) + F32(step - 1) * F32(lr_decay))
    g += penalize(w, l1, l2)
    grad_sq += g * g
    w -= g / (np.sqrt(grad_sq) + F32(eps)) * clr
    return w
