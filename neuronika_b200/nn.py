"""Layers: the mirror of neuronika-nn's Linear and Conv2d (neuronika-nn/src/lib.rs:406-448, 724-815)."""
from __future__ import annotations

import math

import numpy as np

from . import variable as V
from .device import F32, Device


def uniform(rng: np.random.Generator, shape, low: float, high: float) -> np.ndarray:
    """`init::uniform` (neuronika-nn/src/init.rs:177-183); seeded here, thread_rng in the reference."""
    return rng.uniform(low, high, size=shape).astype(np.float32)


class ZeroPad:
    """`Zero` padding mode (pad/zero/mod.rs:14-22)."""
    value = 0.0


class ConstantPad:
    """`Constant(value)` padding mode (pad/constant/mod.rs:14-39)."""

    def __init__(self, value: float):
        self.value = float(value)


class Linear:
    """y = x.A^T + b;  weight (out, in), bias (out,), both ~ U(-k, k), k = sqrt(1/in)
    (neuronika-nn/src/lib.rs:406-448)."""

    def __init__(self, device: Device, in_features: int, out_features: int, dtype=F32, grad_dtype=None,
                 rng: np.random.Generator | None = None):
        rng = rng or np.random.default_rng()
        k = math.sqrt(1.0 / in_features)
        self.weight = V.from_ndarray(device, uniform(rng, (out_features, in_features), -k, k), dtype).requires_grad(grad_dtype)
        self.bias = V.from_ndarray(device, uniform(rng, (out_features,), -k, k), dtype).requires_grad(grad_dtype)

    def forward(self, input: V.Var) -> V.VarDiff:
        """`input.mm_t(self.weight.clone()).into() + self.bias.clone()` (:441-447)"""
        return input.mm_t(self.weight) + self.bias

    def parameters(self):
        return [self.weight, self.bias]


class Conv2d:
    """2-D convolution layer: weight (Cout, Cin, kh, kw), bias (Cout, 1, 1) ~ U(-k, k),
    k = sqrt(1/(Cin*kh*kw)) (neuronika-nn/src/lib.rs:724-788).  The reference's `forward` is
    `todo!()` (:809-814); the documented intent (:789-808) is implemented:
    pad(input) -> weight.convolution(padded, stride, dilation, 1) + bias."""

    def __init__(self, device: Device, in_channels: int, out_channels: int, kernel_size, padding=(0, 0),
                 padding_mode=None, stride=(1, 1), dilation=(1, 1), dtype=F32, grad_dtype=None,
                 rng: np.random.Generator | None = None):
        rng = rng or np.random.default_rng()
        kh, kw = kernel_size
        k = math.sqrt(1.0 / (in_channels * kh * kw))
        self.padding, self.stride, self.dilation = tuple(padding), tuple(stride), tuple(dilation)
        self.padding_mode = padding_mode or ZeroPad()
        self.weight = V.from_ndarray(device, uniform(rng, (out_channels, in_channels, kh, kw), -k, k), dtype).requires_grad(grad_dtype)
        self.bias = V.from_ndarray(device, uniform(rng, (out_channels, 1, 1), -k, k), dtype).requires_grad(grad_dtype)

    def forward(self, input: V.Var) -> V.VarDiff:
        x = input.pad(self.padding, self.padding_mode.value) if any(self.padding) else input
        return self.weight.convolution(x, self.stride, self.dilation, 1) + self.bias

    def parameters(self):
        return [self.weight, self.bias]
