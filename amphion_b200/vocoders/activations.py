"""Parameter holders + native forward for the anti-aliased periodic activations.

Mirrors (names, state-dict keys, constructor arguments) the reference's
``modules/activation_functions/snake.py`` (Snake :11, SnakeBeta :64) and
``modules/anti_aliasing/{act,resample,filter}.py`` (Activation1d act.py:14,
UpSample1d resample.py:17, DownSample1d :48, LowPassFilter1d filter.py:64).
The arithmetic is the CUDA kernel behind ``ab_activation1d_forward``.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
from torch import nn

from .. import _capi


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """Kaiser-windowed sinc low-pass, normalised to unit sum, shape [1,1,k]
    (the filter design of filter.py:30-61)."""
    half = kernel_size // 2
    att = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    if att > 50.0:
        beta = 0.1102 * (att - 8.7)
    elif att >= 21.0:
        beta = 0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        time = torch.arange(-half, half) + 0.5
    else:
        time = torch.arange(kernel_size) - half
    if cutoff == 0:
        return torch.zeros(1, 1, kernel_size)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return (filt / filt.sum()).view(1, 1, kernel_size)


class _PeriodicBase(nn.Module):
    has_beta = False

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(init * alpha, requires_grad=alpha_trainable)
        if self.has_beta:
            self.beta = nn.Parameter(init.clone() * alpha, requires_grad=alpha_trainable)
        self.no_div_by_zero = 1e-9


class Snake(_PeriodicBase):
    """x + 1/a * sin^2(a x)   (snake.py:51-61)."""


class SnakeBeta(_PeriodicBase):
    """x + 1/b * sin^2(a x)   (snake.py:110-122)."""
    has_beta = True


class _FilterHolder(nn.Module):
    def __init__(self, kernel_size=12, ratio=2):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = kernel_size
        self.register_buffer("filter", kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, kernel_size))


class UpSample1d(_FilterHolder):
    pass


class LowPassFilter1d(_FilterHolder):
    pass


class DownSample1d(nn.Module):
    def __init__(self, ratio=2, kernel_size=12):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = kernel_size
        self.lowpass = LowPassFilter1d(kernel_size, ratio)


class Activation1d(nn.Module):
    """down2x(act(up2x(x))) with replicate edges, one fused CUDA kernel."""

    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2,
                 up_kernel_size: int = 12, down_kernel_size: int = 12):
        super().__init__()
        if (up_ratio, down_ratio, up_kernel_size, down_kernel_size) != (2, 2, 12, 12):
            raise NotImplementedError("amphion_b200: Activation1d is built for ratio 2 / kernel 12 "
                                      "(the only instance on the reference's path)")
        self.up_ratio, self.down_ratio = up_ratio, down_ratio
        self.act = activation
        self.upsample = UpSample1d(up_kernel_size, up_ratio)
        self.downsample = DownSample1d(down_ratio, down_kernel_size)

    def forward(self, x):
        _capi.require_cuda(x, "Activation1d.forward")
        if x.dim() != 3:
            raise ValueError("Activation1d expects [B, C, T]")
        x = x.contiguous().float()
        y = torch.empty_like(x)
        alpha = self.act.alpha.detach().float().contiguous()
        beta = self.act.beta.detach().float().contiguous() if self.act.has_beta else alpha
        fu = self.upsample.filter.reshape(-1).float().contiguous()
        fd = self.downsample.lowpass.filter.reshape(-1).float().contiguous()
        B, Cn, T = x.shape
        with torch.cuda.device(x.device):   # the launch and the stream belong to x's device
            _capi.check(_capi.lib.ab_activation1d_forward(
                _capi.ptr(x), _capi.ptr(y), B, Cn, T, _capi.ptr(alpha), _capi.ptr(beta),
                int(bool(self.act.alpha_logscale)), _capi.ptr(fu), _capi.ptr(fd), _capi.stream_ptr()),
                "ab_activation1d_forward")
        return y
