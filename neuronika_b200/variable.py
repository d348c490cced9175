"""Var / VarDiff: thin Python handles over the C++ graph (csrc/nk_graph.cpp, include/nk_graph.h).

Method names and semantics follow the reference's op surface (neuronika-variable/src/var.rs,
vardiff.rs): ops are lazy, `.forward()` recomputes the tape, `.backward(seed)` runs the backward
tape in reverse and every node accumulates into its operands' gradients."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .device import BF16, F32, CuArray, Device, as_shape, dtype_of

lib = L.lib
vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
pvp = C.POINTER(vp)
pi64 = C.POINTER(C.c_int64)
pf32 = C.POINTER(C.c_float)

_G = {
    "nkg_last_error": (C.c_char_p, []),
    "nkg_leaf": (i32, [vp, i32, pi64, i32, pvp]),
    "nkg_leaf_external": (i32, [vp, i32, pi64, i32, vp, pvp]),
    "nkg_requires_grad": (i32, [vp, i32, vp, pvp]),
    "nkg_clone": (i32, [vp, pvp]),
    "nkg_release": (i32, [vp]),
    "nkg_is_diff": (i32, [vp]),
    "nkg_ndim": (i32, [vp]),
    "nkg_shape": (i32, [vp, pi64]),
    "nkg_dtype": (i32, [vp]),
    "nkg_grad_dtype": (i32, [vp]),
    "nkg_data_ptr": (vp, [vp]),
    "nkg_grad_ptr": (vp, [vp]),
    "nkg_history_len": (i32, [vp]),
    "nkg_backward_history_len": (i32, [vp]),
    "nkg_forward": (i32, [vp]),
    "nkg_backward": (i32, [vp, f32]),
    "nkg_zero_grad": (i32, [vp]),
    "nkg_no_grad": (i32, [vp]),
    "nkg_with_grad": (i32, [vp]),
    "nkg_set_fusion": (i32, [i32]),
    "nkg_mm": (i32, [vp, vp, pvp]),
    "nkg_mm_t": (i32, [vp, vp, pvp]),
    "nkg_add": (i32, [vp, vp, pvp]),
    "nkg_relu": (i32, [vp, pvp]),
    "nkg_softmax": (i32, [vp, i32, pvp]),
    "nkg_log_softmax": (i32, [vp, i32, pvp]),
    "nkg_sum": (i32, [vp, pvp]),
    "nkg_mean": (i32, [vp, pvp]),
    "nkg_mse_loss": (i32, [vp, vp, i32, pvp]),
    "nkg_nll_loss": (i32, [vp, vp, i32, pvp]),
    "nkg_pad": (i32, [vp, i64, i64, f32, pvp]),
    "nkg_convolution": (i32, [vp, vp, i64, i64, i64, i64, i64, pvp]),
    "nkg_flatten": (i32, [vp, pvp]),
    "nkg_sgd_step": (i32, [vp, vp, vp, f32, f32, f32, f32, i32, f32]),
    "nkg_sub": (i32, [vp, vp, pvp]),
    "nkg_mul": (i32, [vp, vp, pvp]),
    "nkg_div": (i32, [vp, vp, pvp]),
    "nkg_unary": (i32, [vp, i32, i32, pvp]),
    "nkg_neg": (i32, [vp, pvp]),
    "nkg_exp": (i32, [vp, pvp]),
    "nkg_ln": (i32, [vp, pvp]),
    "nkg_sqrt": (i32, [vp, pvp]),
    "nkg_sigmoid": (i32, [vp, pvp]),
    "nkg_tanh": (i32, [vp, pvp]),
    "nkg_softplus": (i32, [vp, pvp]),
    "nkg_leaky_relu": (i32, [vp, pvp]),
    "nkg_pow": (i32, [vp, i32, pvp]),
    "nkg_transpose": (i32, [vp, pvp]),
    "nkg_pad_mode": (i32, [vp, i32, pi64, i32, f32, pvp]),
    "nkg_mv": (i32, [vp, vp, pvp]),
    "nkg_vm": (i32, [vp, vp, pvp]),
    "nkg_vv": (i32, [vp, vp, pvp]),
    "nkg_convolution_nd": (i32, [vp, vp, i32, pi64, pi64, i64, pvp]),
    "nkg_adam_step": (i32, [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, f32]),
    "nkg_rmsprop_step": (i32, [vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, f32, f32]),
    "nkg_adagrad_step": (i32, [vp, vp, vp, i64, f32, f32, f32, f32, f32, f32]),
    "nkg_set_grad_hook": (i32, [vp, vp, vp, i32]),
    "nkg_set_grad_rs": (i32, [vp, i32, i32, pvp, vp, vp]),
}
for _n, (_r, _a) in _G.items():
    _f = getattr(lib, _n)
    _f.restype = _r
    _f.argtypes = _a


GRAD_HOOK = C.CFUNCTYPE(None, vp, i64, i64)
GRAD_RS_HOOK = C.CFUNCTYPE(None, vp, i32)


def graph_symbols():
    return sorted(_G)


def _ck(rc: int) -> None:
    if rc != 0:
        raise L.NkError(rc, lib.nkg_last_error().decode())


def set_fusion(level) -> None:
    """Host-side peephole fusion: False / 0 off, True / 1 (default) fusions that are invisible for any use of a tape,
    2 additionally fuses a layer's ReLU backward into the dX GEMM above it (exact for one backward() per tape, which
    is what a training loop does; a second backward() on such a tape raises); 3 additionally takes that layer's bias
    gradient in the same epilogue (nk_gemm_relu_bwd_colsum; correct, four launches fewer per MLP step, no measured gain,
    so nothing uses it by default)."""
    _ck(lib.nkg_set_fusion(int(level)))


class Reduction:
    """neuronika-variable/src/lib.rs:29-36"""
    Mean = 0
    Sum = 1


class Var:
    """A non-differentiable variable (var.rs:25-61) with data on the device."""

    def __init__(self, device: Device, handle):
        self.device = device
        self._h = vp(handle) if not isinstance(handle, vp) else handle

    def __del__(self):
        try:
            if self._h:
                lib.nkg_release(self._h)
                self._h = None
        except Exception:
            pass

    # ---- wrapping results
    def _wrap(self, out: vp):
        cls = VarDiff if lib.nkg_is_diff(out) else Var
        return cls(self.device, out)

    def _unary(self, fn, *args):
        out = vp()
        _ck(fn(self._h, *args, C.byref(out)))
        return self._wrap(out)

    def _binary(self, fn, other: "Var", *args):
        out = vp()
        _ck(fn(self._h, other._h, *args, C.byref(out)))
        return self._wrap(out)

    # ---- introspection
    @property
    def shape(self):
        n = lib.nkg_ndim(self._h)
        buf = (C.c_int64 * max(1, n))()
        _ck(lib.nkg_shape(self._h, buf))
        return tuple(int(buf[i]) for i in range(n))

    @property
    def dtype(self) -> int:
        return int(lib.nkg_dtype(self._h))

    def data(self) -> np.ndarray:
        """Copy of the data on the host (`Var::data`, var.rs:67-69).  Zeros before forward()."""
        return self.data_array().as_ndarray()

    def data_array(self) -> CuArray:
        ptr = lib.nkg_data_ptr(self._h)
        if not ptr:
            raise L.NkError(-1, lib.nkg_last_error().decode())
        return CuArray(self.device, self.shape, self.dtype, ptr=ptr, owner=self)

    def set_data(self, array: np.ndarray) -> None:
        """`*var.data_mut() = array` (var.rs:75-77)."""
        self.data_array().copy_from(array)

    def history_len(self) -> int:
        return int(lib.nkg_history_len(self._h))

    def clone(self) -> "Var":
        out = vp()
        _ck(lib.nkg_clone(self._h, C.byref(out)))
        return type(self)(self.device, out)

    # ---- differentiability
    def requires_grad(self, grad_dtype=None, grad_array: CuArray | None = None) -> "VarDiff":
        """`Var::requires_grad` (var.rs:104-107)."""
        out = vp()
        gd = -1 if grad_dtype is None else dtype_of(grad_dtype)
        _ck(lib.nkg_requires_grad(self._h, gd, grad_array.ptr if grad_array is not None else None, C.byref(out)))
        v = VarDiff(self.device, out)
        v._grad_owner = grad_array
        return v

    # ---- execution
    def forward(self) -> None:
        """var.rs:110-128"""
        _ck(lib.nkg_forward(self._h))

    # ---- operators (same names as the reference's methods / traits)
    def mm(self, other): return self._binary(lib.nkg_mm, other)                   # MatMatMul, core lib.rs:4-13
    def mm_t(self, other): return self._binary(lib.nkg_mm_t, other)               # MatMatMulT, core lib.rs:19-28
    def __add__(self, other): return self._binary(lib.nkg_add, other)
    def __sub__(self, other): return self._binary(lib.nkg_sub, other)             # subtraction/mod.rs
    def __mul__(self, other): return self._binary(lib.nkg_mul, other)             # multiplication/mod.rs
    def __truediv__(self, other): return self._binary(lib.nkg_div, other)         # division/mod.rs
    def __neg__(self): return self._unary(lib.nkg_neg)                            # negation/mod.rs
    def exp(self): return self._unary(lib.nkg_exp)
    def ln(self): return self._unary(lib.nkg_ln)
    def sqrt(self): return self._unary(lib.nkg_sqrt)
    def sigmoid(self): return self._unary(lib.nkg_sigmoid)
    def tanh(self): return self._unary(lib.nkg_tanh)
    def softplus(self): return self._unary(lib.nkg_softplus)
    def leaky_relu(self): return self._unary(lib.nkg_leaky_relu)
    def pow(self, exp: int): return self._unary(lib.nkg_pow, int(exp))            # power/mod.rs (`powi`)
    def t(self): return self._unary(lib.nkg_transpose)                            # transpose/mod.rs (reverses all axes)
    def mv(self, vector): return self._binary(lib.nkg_mv, vector)                 # MatVecMul, core lib.rs
    def vm(self, matrix): return self._binary(lib.nkg_vm, matrix)                 # VecMatMul
    def vv(self, other): return self._binary(lib.nkg_vv, other)                   # VecVecMul
    def relu(self): return self._unary(lib.nkg_relu)
    def softmax(self, axis: int): return self._unary(lib.nkg_softmax, int(axis))
    def log_softmax(self, axis: int): return self._unary(lib.nkg_log_softmax, int(axis))
    def sum(self): return self._unary(lib.nkg_sum)
    def mean(self): return self._unary(lib.nkg_mean)
    def mse_loss(self, target, reduction=Reduction.Mean): return self._binary(lib.nkg_mse_loss, target, int(reduction))
    def nll_loss(self, target, reduction=Reduction.Mean): return self._binary(lib.nkg_nll_loss, target, int(reduction))
    def flatten(self): return self._unary(lib.nkg_flatten)

    def pad(self, padding, value: float = 0.0, mode: str = "constant"):
        """`pad(padding, mode)` (var.rs:726-737): Zero / Constant(value) / Reflective / Replicative over the 1..3
        sample dimensions of a (N, C, ...) operand."""
        modes = {"constant": L.NK_PAD_CONSTANT, "zero": L.NK_PAD_CONSTANT, "reflective": L.NK_PAD_REFLECTIVE,
                 "replicative": L.NK_PAD_REPLICATIVE}
        if mode not in modes:
            raise L.NkError(-1, f"unknown padding mode {mode!r}")
        padding = tuple(int(p) for p in padding)
        if len(padding) == 2 and modes[mode] == L.NK_PAD_CONSTANT:
            return self._unary(lib.nkg_pad, padding[0], padding[1], float(value))
        return self._unary(lib.nkg_pad_mode, len(padding), L.shape_arr(padding), modes[mode], float(value))

    def convolution(self, input, stride=(1, 1), dilation=(1, 1), groups: int = 1):
        """`kernel.convolution(input, stride, dilation, groups)` -- the receiver is the kernel
        (Convolution trait, core lib.rs:91-106; var.rs:704-716)."""
        nsp = len(self.shape) - 2
        if len(stride) != nsp:
            raise L.NkError(-1, f"Invalid stride {list(stride)} for {nsp}d conv.")
        if len(dilation) != nsp:
            raise L.NkError(-1, f"Invalid dilation {list(dilation)} for {nsp}d conv.")
        if nsp != 2:       # 1-d / 3-d operands (convolution/mod.rs is generic over the sample dimensions)
            return self._binary(lib.nkg_convolution_nd, input, nsp, L.shape_arr(stride), L.shape_arr(dilation),
                                int(groups))
        return self._binary(lib.nkg_convolution, input, int(stride[0]), int(stride[1]), int(dilation[0]),
                            int(dilation[1]), int(groups))

    def item(self) -> float:
        return float(self.data().reshape(()))


class VarDiff(Var):
    """A differentiable variable (vardiff.rs:25-65)."""

    def __init__(self, device, handle):
        super().__init__(device, handle)
        self._grad_owner = None

    @property
    def grad_dtype(self) -> int:
        return int(lib.nkg_grad_dtype(self._h))

    def grad(self) -> np.ndarray:
        """`VarDiff::grad` (vardiff.rs:84-86)."""
        return self.grad_array().as_ndarray()

    def grad_array(self) -> CuArray:
        ptr = lib.nkg_grad_ptr(self._h)
        if not ptr:
            raise L.NkError(-1, "Trying to get a de-allocated gradient. Switch on the gradients first by "
                                "using `.with_grad()`")
        return CuArray(self.device, self.shape, self.grad_dtype, ptr=ptr, owner=self)

    def backward(self, seed: float) -> None:
        """vardiff.rs:125-141"""
        _ck(lib.nkg_backward(self._h, float(seed)))

    def zero_grad(self) -> None:
        _ck(lib.nkg_zero_grad(self._h))

    def no_grad(self) -> None:
        _ck(lib.nkg_no_grad(self._h))

    def with_grad(self) -> None:
        _ck(lib.nkg_with_grad(self._h))

    def backward_history_len(self) -> int:
        return int(lib.nkg_backward_history_len(self._h))

    def set_grad_hook(self, fn, row_chunks: int = 1) -> None:
        """Call `fn(begin, end)` from inside backward() as soon as elements [begin, end) of this leaf's gradient are
        final for the running pass (used to overlap the data-parallel all-reduce with the rest of backward).  With
        row_chunks > 1 a matmul backward that writes the gradient last delivers it in that many row blocks."""
        if fn is None:
            self._hook_ref = None
            _ck(lib.nkg_set_grad_hook(self._h, None, None, 1))
            return
        cb = GRAD_HOOK(lambda _user, b, e: fn(int(b), int(e)))
        self._hook_ref = cb  # keep the trampoline alive
        _ck(lib.nkg_set_grad_hook(self._h, C.cast(cb, vp), None, int(row_chunks)))


    def set_grad_rs(self, world: int, rank: int, slot_ptrs, fn) -> None:
        """Fused data-parallel exchange (nk_b200.h nk_gemm_rs): the matmul backward node that produces this leaf's
        gradient pushes row shard o of its product into `slot_ptrs[o]` (rank o's slot buffer, peer-mapped) and then
        calls `fn(pushed)`; pushed == 0 means the gradient was computed locally (caller falls back to all-reduce)."""
        if world <= 1:
            self._rs_ref = None
            _ck(lib.nkg_set_grad_rs(self._h, 0, 0, None, None, None))
            return
        arr = (vp * world)(*[int(p) for p in slot_ptrs])
        cb = GRAD_RS_HOOK(lambda _user, pushed: fn(int(pushed)))
        self._rs_ref = (cb, arr)
        _ck(lib.nkg_set_grad_rs(self._h, int(world), int(rank), arr, C.cast(cb, vp), None))


# ---- constructors (neuronika-variable/src/lib.rs:51-240), on a device
def zeros(device: Device, shape, dtype=F32) -> Var:
    shape = as_shape(shape)
    out = vp()
    _ck(lib.nkg_leaf(device.ctx, len(shape), L.shape_arr(shape), dtype_of(dtype), C.byref(out)))
    return Var(device, out)


def full(device: Device, shape, value: float, dtype=F32) -> Var:
    v = zeros(device, shape, dtype)
    v.data_array().fill_(value)
    return v


def ones(device: Device, shape, dtype=F32) -> Var:
    return full(device, shape, 1.0, dtype)


def from_ndarray(device: Device, array: np.ndarray, dtype=F32) -> Var:
    v = zeros(device, np.shape(array), dtype)
    v.set_data(array)
    return v


def rand(device: Device, shape, dtype=F32, rng: np.random.Generator | None = None) -> Var:
    """U(0,1) like `neuronika::rand`; host-generated (the reference uses an unseeded thread_rng)."""
    rng = rng or np.random.default_rng()
    return from_ndarray(device, rng.random(as_shape(shape), dtype=np.float32), dtype)


def from_device_memory(device: Device, array: CuArray) -> Var:
    """Leaf over caller-owned device memory (parameter / gradient buckets)."""
    out = vp()
    _ck(lib.nkg_leaf_external(device.ctx, array.ndim, L.shape_arr(array.shape), array.dtype, array.ptr, C.byref(out)))
    v = Var(device, out)
    v._keep = array  # keep the storage alive as long as the handle
    return v
