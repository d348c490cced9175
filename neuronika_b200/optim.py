"""Optimizer registry and SGD: the mirror of neuronika-optim (optimizer.rs:4-104, sgd/mod.rs:11-236,
penalty.rs:2-79 -- L2 only, the penalty on the hot path)."""
from __future__ import annotations

import numpy as np

from . import variable as V
from .device import F32, CuArray


class L2:
    """`L2::penalize(w) = 2*lambda*w` (penalty.rs:63-67)."""

    def __init__(self, lambda_: float = 0.0):
        self.lambda_ = float(lambda_)


class NoPenalty(L2):
    def __init__(self):
        super().__init__(0.0)


class Optimizer:
    """`Optimizer<T>`: register / step / zero_grad / get_lr / set_lr (optimizer.rs:33-95)."""

    def __init__(self, status):
        self.status = status
        self.params = []

    def get_lr(self) -> float:
        return self.status.lr

    def set_lr(self, lr: float) -> None:
        self.status.lr = float(lr)

    def register(self, variable: V.VarDiff) -> None:
        self.params.append(self.status.into_param(variable))

    def step(self) -> None:
        for p in self.params:
            p.optimize()

    def zero_grad(self) -> None:
        for p in self.params:
            p.zero_grad()


class _SGDParam:
    """SGDParam (sgd/mod.rs:150-236): momentum buffer created on first use."""

    def __init__(self, variable: V.VarDiff, status: "StochasticGD"):
        self.variable, self.status, self.buffer, self.master = variable, status, None, None
        if status.master_weights and variable.dtype != F32:
            self.master = variable.data_array().astype(F32)

    def optimize(self) -> None:
        s = self.status
        use_mom = s.momentum is not None and s.momentum > np.finfo(np.float32).eps
        if use_mom and self.buffer is None:
            self.buffer = CuArray(self.variable.device, self.variable.shape, F32)
        if not use_mom:
            self.buffer = None
        V._ck(V.lib.nkg_sgd_step(self.variable._h, self.buffer.ptr if self.buffer is not None else None,
                                 self.master.ptr if self.master is not None else None, float(s.lr),
                                 float(s.penalty.lambda_), float(s.momentum or 0.0), float(s.dampening or 0.0),
                                 int(bool(s.nesterov)), float(s.grad_scale)))

    def zero_grad(self) -> None:
        self.variable.zero_grad()


class StochasticGD:
    """`StochasticGD::new(lr, penalty, momentum, dampening, nesterov) -> Optimizer<Self>` (sgd/mod.rs:43-85),
    same argument validation.  `grad_scale` (1/world_size under data parallel) and `master_weights`
    (f32 master copy of bf16 parameters, kept as optimizer state) are additions."""

    def __init__(self, lr, penalty, momentum, dampening, nesterov, grad_scale=1.0, master_weights=False):
        if momentum is None:
            assert dampening is None and not nesterov, \
                "Dampening and Nesterov momentum flag should be enabled together with momentum."
        if dampening is not None:
            assert 0.0 <= dampening <= 1.0, f"Dampening value should be between 0.0 and 1.0, got: {dampening}"
        self.lr, self.penalty, self.momentum, self.dampening, self.nesterov = float(lr), penalty, momentum, dampening, nesterov
        self.grad_scale, self.master_weights = float(grad_scale), bool(master_weights)

    @staticmethod
    def new(lr, penalty=None, momentum=None, dampening=None, nesterov=False, **kw) -> Optimizer:
        return Optimizer(StochasticGD(lr, penalty or NoPenalty(), momentum, dampening, nesterov, **kw))

    def into_param(self, variable: V.VarDiff) -> _SGDParam:
        return _SGDParam(variable, self)
