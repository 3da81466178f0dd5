"""NSF-HiFiGAN generator — drop-in for the reference class
``models/vocoders/gan/generator/nsfhifigan.py:181`` (``_vocoders["nsfhifigan"]``,
models/vocoders/vocoder_inference.py:39-49): same constructor (``cfg``), same
parameter names (``m_source.l_linear.*``, ``noise_convs.*``, ``conv_pre``, ``ups``,
``resblocks``, ``conv_post``), ``forward(mel[B,n_mel,T], f0[B,T_f0]) -> wav[B,1,T*hop]``.

What the reference's forward computes (:262-283): the harmonic source is generated
and pushed through ``noise_convs[i]``, but ``x_source = x[:, :, :length]`` (:269)
replaces it by the stage tensor itself, so each stage evaluates ``x = ups(x) + ups(x)``
and the source (with its ``torch.rand`` / ``torch.randn`` draws) only contributes
its LENGTH.  The native pipeline reproduces those samples: HiFi-GAN kernels with the
``ups`` weights doubled at load time (C ABI kind ``AB_GEN_NSFHIFIGAN``).  The source
module is kept as a parameter holder so reference checkpoints load unchanged; it is
not executed (the reference's result does not depend on it, but note that the
reference advances the global torch RNG there and this path does not).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.nn.utils import remove_weight_norm, weight_norm

from .. import _capi
from .generator import ConvBlock, NativeGenerator, init_weights


class SourceModuleHnNSF(nn.Module):
    """Parameter holder for ``m_source`` (nsfhifigan.py:162-178): ``l_linear`` merges
    harmonic_num + 1 sine waves.  Never evaluated on this path (see module docstring)."""

    def __init__(self, fs, harmonic_num=0, amp=0.1, noise_std=0.003, voiced_threshold=0):
        super().__init__()
        self.fs, self.harmonic_num, self.amp, self.noise_std = fs, harmonic_num, amp, noise_std
        self.voiced_threshold = voiced_threshold
        self.l_linear = nn.Linear(harmonic_num + 1, 1)


class NSFHiFiGAN(NativeGenerator):
    kind = "nsfhifigan"
    hp_key = "nsfhifigan"

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        hp = cfg.model.nsfhifigan
        if str(hp.resblock) != "1":
            # the reference's own ResBlock2 calls super(ResBlock1, self).__init__() (nsfhifigan.py:111) and
            # raises this very error when constructed
            raise TypeError("super(type, obj): obj must be an instance or subtype of type")
        self.num_kernels = len(hp.resblock_kernel_sizes)
        self.num_upsamples = len(hp.upsample_rates)
        # registration order as in the reference (:189-200) so that state_dict() lists the same keys in the same order
        self.m_source = SourceModuleHnNSF(fs=cfg.preprocess.sample_rate, harmonic_num=hp.harmonic_num)
        self.noise_convs = nn.ModuleList()
        c0 = hp.upsample_initial_channel
        self.conv_pre = weight_norm(nn.Conv1d(cfg.preprocess.n_mel, c0, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        rates = [int(u) for u in hp.upsample_rates]
        for i, (u, k) in enumerate(zip(rates, hp.upsample_kernel_sizes)):
            c_cur = c0 // (2 ** (i + 1))
            self.ups.append(weight_norm(nn.ConvTranspose1d(c0 // (2 ** i), c_cur, k, u, padding=(k - u) // 2)))
            if i + 1 < len(rates):
                stride_f0 = int(np.prod(rates[i + 1:]))
                self.noise_convs.append(nn.Conv1d(1, c_cur, kernel_size=stride_f0 * 2, stride=stride_f0,
                                                  padding=stride_f0 // 2))
            else:
                self.noise_convs.append(nn.Conv1d(1, c_cur, kernel_size=1))
        self.resblocks = nn.ModuleList()
        for i in range(self.num_upsamples):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(hp.resblock_kernel_sizes, hp.resblock_dilation_sizes):
                self.resblocks.append(ConvBlock(cfg, ch, k, d, "1"))
        self.conv_post = weight_norm(nn.Conv1d(ch, 1, 7, 1, padding=3))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)
        self.upp = int(np.prod(rates))

    def forward(self, x, f0):
        """mel [B, n_mel, T], f0 [B, T_f0] -> wav [B, 1, n].  The harmonic source only contributes its LENGTH
        (``x_source = x[:, :, :length]``, :269): when it covers every stage, n = T*hop; when the f0 track is shorter
        than the mel or a source stride is odd, each stage is truncated to the source length as in the reference
        (:264-268) and n = ``ab_generator_output_samples(T, T_f0)``."""
        if f0 is None:
            raise TypeError("NSFHiFiGAN.forward() missing 1 required positional argument: 'f0'")
        if f0.dim() != 2 or f0.shape[0] != x.shape[0]:
            raise ValueError(f"expected f0 of shape [B, T_f0] with B={x.shape[0]}, got {tuple(f0.shape)}")
        T, Tf = int(x.shape[2]), int(f0.shape[1])
        n = int(_capi.lib.ab_generator_output_samples(self._ensure_handle(), T, Tf))
        if n == T * self.upp:
            return super().forward(x)
        if n <= 0:
            raise ValueError(f"the f0 track ({Tf} frames) leaves no samples")
        # the C ABI writes T*hop-sample rows when no truncation happens and n-sample rows otherwise
        out = torch.empty(x.shape[0], 1, n, dtype=torch.float32, device=x.device)
        self.set_option("nsf_source_frames", Tf)
        return self._forward_native(x, out=out, out_samples=n)

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)
