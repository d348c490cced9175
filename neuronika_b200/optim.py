"""Optimizer registry and optimizers: the mirror of neuronika-optim (optimizer.rs:4-104, sgd/mod.rs:11-236,
adam/mod.rs, amsgrad/mod.rs, rmsprop/mod.rs, adagrad/mod.rs, penalty.rs:2-79).  Every `optimize()` is ONE fused kernel
over the parameter (nk_sgd_step / nk_adam_step / nk_rmsprop_step / nk_adagrad_step); optimizer state lives on the
device in f32."""
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


class L1:
    """`L1::penalize(w) = lambda*signum(w)` (penalty.rs:69-73)."""

    def __init__(self, lambda_: float):
        self.lambda_ = float(lambda_)


class ElasticNet:
    """`lambda_l1*signum(w) + 2*lambda_l2*w` (penalty.rs:75-79)."""

    def __init__(self, lambda_l1: float, lambda_l2: float):
        self.lambda_l1, self.lambda_l2 = float(lambda_l1), float(lambda_l2)


def _l1_l2(penalty):
    """(l1, l2) coefficients of a penalty object for the fused kernels."""
    if isinstance(penalty, ElasticNet):
        return penalty.lambda_l1, penalty.lambda_l2
    if isinstance(penalty, L1):
        return penalty.lambda_, 0.0
    return 0.0, penalty.lambda_


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


# ------------------------------------------------------------------------------------------- Adam family (8-f rank 2)
def _state(variable: V.VarDiff) -> CuArray:
    return CuArray(variable.device, variable.shape, F32)          # zero-filled, like Array::zeros(raw_dim)


def _ptr(a):
    return a.ptr if a is not None else None


class _MasterMixin:
    def _init_master(self, variable, status):
        self.master = None
        if getattr(status, "master_weights", False) and variable.dtype != F32:
            self.master = variable.data_array().astype(F32)


class _AdamParam(_MasterMixin):
    """AdamParam / AMSGradParam (adam/mod.rs:113-175, amsgrad/mod.rs:136-210)."""

    def __init__(self, variable, status):
        self.variable, self.status, self.step = variable, status, 0
        self.exp_avg, self.exp_avg_sq = _state(variable), _state(variable)
        self.max_exp_avg_sq = _state(variable) if status.amsgrad else None
        self._init_master(variable, status)

    def optimize(self) -> None:
        s = self.status
        self.step += 1
        l1, l2 = _l1_l2(s.penalty)
        V._ck(V.lib.nkg_adam_step(self.variable._h, self.exp_avg.ptr, self.exp_avg_sq.ptr, _ptr(self.max_exp_avg_sq),
                                  _ptr(self.master), self.step, float(s.lr), float(s.beta1), float(s.beta2),
                                  float(s.eps), l1, l2, float(s.grad_scale)))

    def zero_grad(self) -> None:
        self.variable.zero_grad()


class Adam:
    """`Adam::new(lr, beta1, beta2, penalty, eps)` (adam/mod.rs:43-60)."""
    amsgrad = False

    def __init__(self, lr, beta1, beta2, penalty, eps, grad_scale=1.0, master_weights=False):
        self.lr, self.beta1, self.beta2, self.penalty, self.eps = float(lr), float(beta1), float(beta2), penalty, float(eps)
        self.grad_scale, self.master_weights = float(grad_scale), bool(master_weights)

    @classmethod
    def new(cls, lr, beta1=0.9, beta2=0.999, penalty=None, eps=1e-8, **kw) -> Optimizer:
        return Optimizer(cls(lr, beta1, beta2, penalty or NoPenalty(), eps, **kw))

    def into_param(self, variable):
        return _AdamParam(variable, self)


class AMSGrad(Adam):
    """`AMSGrad::new(lr, beta1, beta2, penalty, eps)` (amsgrad/mod.rs:45-62)."""
    amsgrad = True


class _RMSPropParam(_MasterMixin):
    """RMSPropParam (rmsprop/mod.rs:150-305): `buffer` / `grad_avg` exist only while momentum / centered are on."""

    def __init__(self, variable, status):
        self.variable, self.status = variable, status
        self.square_avg = _state(variable)
        self.buffer = _state(variable) if status.momentum is not None else None
        self.grad_avg = _state(variable) if status.centered else None
        self._init_master(variable, status)

    def optimize(self) -> None:
        s = self.status
        use_mom = s.momentum is not None and s.momentum > np.finfo(np.float32).eps
        if use_mom and self.buffer is None:
            self.buffer = _state(self.variable)
        if not use_mom:
            self.buffer = None
        if s.centered and self.grad_avg is None:
            self.grad_avg = _state(self.variable)
        if not s.centered:
            self.grad_avg = None
        l1, l2 = _l1_l2(s.penalty)
        V._ck(V.lib.nkg_rmsprop_step(self.variable._h, self.square_avg.ptr, _ptr(self.grad_avg), _ptr(self.buffer),
                                     _ptr(self.master), float(s.lr), float(s.alpha if s.alpha is not None else 0.0),
                                     float(s.eps), float(s.momentum or 0.0), l1, l2, float(s.grad_scale)))

    def zero_grad(self) -> None:
        self.variable.zero_grad()


class RMSProp:
    """`RMSProp::new(lr, penalty, alpha, momentum, centered, eps)` (rmsprop/mod.rs:67-100), same validation."""

    def __init__(self, lr, penalty, alpha, momentum, centered, eps, grad_scale=1.0, master_weights=False):
        if alpha is not None:
            assert 0.0 <= alpha <= 1.0, f"Dampening value should be between 0.0 and 1.0, got: {alpha}"
        self.lr, self.penalty, self.alpha, self.momentum = float(lr), penalty, alpha, momentum
        self.centered, self.eps = bool(centered), float(eps)
        self.grad_scale, self.master_weights = float(grad_scale), bool(master_weights)

    @staticmethod
    def new(lr, penalty=None, alpha=0.99, momentum=None, centered=False, eps=1e-8, **kw) -> Optimizer:
        return Optimizer(RMSProp(lr, penalty or NoPenalty(), alpha, momentum, centered, eps, **kw))

    def into_param(self, variable):
        return _RMSPropParam(variable, self)


class _AdagradParam(_MasterMixin):
    """AdagradParam (adagrad/mod.rs:96-145)."""

    def __init__(self, variable, status):
        self.variable, self.status, self.step = variable, status, 0
        self.grad_sq = _state(variable)
        self._init_master(variable, status)

    def optimize(self) -> None:
        s = self.status
        self.step += 1
        l1, l2 = _l1_l2(s.penalty)
        V._ck(V.lib.nkg_adagrad_step(self.variable._h, self.grad_sq.ptr, _ptr(self.master), self.step, float(s.lr),
                                     float(s.lr_decay), float(s.eps), l1, l2, float(s.grad_scale)))

    def zero_grad(self) -> None:
        self.variable.zero_grad()


class Adagrad:
    """`Adagrad::new(lr, lr_decay, penalty, eps)` (adagrad/mod.rs:50-63)."""

    def __init__(self, lr, lr_decay, penalty, eps, grad_scale=1.0, master_weights=False):
        self.lr, self.lr_decay, self.penalty, self.eps = float(lr), float(lr_decay), penalty, float(eps)
        self.grad_scale, self.master_weights = float(grad_scale), bool(master_weights)

    @staticmethod
    def new(lr, lr_decay=0.0, penalty=None, eps=1e-10, **kw) -> Optimizer:
        return Optimizer(Adagrad(lr, lr_decay, penalty or NoPenalty(), eps, **kw))

    def into_param(self, variable):
        return _AdagradParam(variable, self)
