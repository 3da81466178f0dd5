"""HiFi-GAN generator — drop-in for the reference class
``models/vocoders/gan/generator/hifigan.py:151`` (``_vocoders["hifigan"]``,
models/vocoders/vocoder_inference.py:39-49): same constructor (``cfg``), same
parameter names, ``forward(mel[B,n_mel,T]) -> wav[B,1,T*hop]``, ``remove_weight_norm``.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
from torch import nn
from torch.nn.utils import remove_weight_norm, weight_norm

from .generator import ConvBlock, NativeGenerator, init_weights


class HiFiGAN(NativeGenerator):
    kind = "hifigan"
    hp_key = "hifigan"

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        hp = cfg.model.hifigan
        self.num_kernels = len(hp.resblock_kernel_sizes)
        self.num_upsamples = len(hp.upsample_rates)
        c0 = hp.upsample_initial_channel
        self.conv_pre = weight_norm(nn.Conv1d(cfg.preprocess.n_mel, c0, 7, 1, padding=3))
        self.ups = nn.ModuleList(
            weight_norm(nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2))
            for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes)))
        self.resblocks = nn.ModuleList()
        for i in range(self.num_upsamples):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(hp.resblock_kernel_sizes, hp.resblock_dilation_sizes):
                self.resblocks.append(ConvBlock(cfg, ch, k, d, hp.resblock))
        self.conv_post = weight_norm(nn.Conv1d(ch, 1, 7, 1, padding=3))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)


class HiFiGAN_vits(NativeGenerator):
    """Drop-in for ``HiFiGAN_vits`` (hifigan.py:376-449), the waveform decoder inside VITS
    (models/tts/vits/vits.py:215-378 builds it as ``self.dec``): positional constructor, plain
    ``conv_pre`` / bias-free ``conv_post``, optional global conditioning ``cond`` (1x1 conv on the
    speaker embedding), ``forward(x[B, initial_channel, T], g[B, gin, 1] = None)``.  The reference's
    forward calls the residual blocks without a mask (:437-441), so the ``x_mask`` branches of
    ``ResBlock{1,2}_vits`` never run and the blocks are the plain HiFi-GAN ones."""
    kind = "hifigan"
    hp_key = "hifigan"

    def __init__(self, initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, gin_channels=0):
        super().__init__()
        hop = 1
        for u in upsample_rates:
            hop *= int(u)
        hp = SimpleNamespace(resblock=str(resblock), upsample_rates=list(upsample_rates),
                             upsample_kernel_sizes=list(upsample_kernel_sizes),
                             upsample_initial_channel=upsample_initial_channel,
                             resblock_kernel_sizes=list(resblock_kernel_sizes),
                             resblock_dilation_sizes=[list(d) for d in resblock_dilation_sizes])
        self.cfg = SimpleNamespace(preprocess=SimpleNamespace(n_mel=int(initial_channel), hop_size=hop),
                                   model=SimpleNamespace(hifigan=hp))
        self.gin_channels = int(gin_channels)
        self.conv_post_no_bias = True
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        c0 = upsample_initial_channel
        self.conv_pre = nn.Conv1d(initial_channel, c0, 7, 1, padding=3)
        self.ups = nn.ModuleList(
            weight_norm(nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2))
            for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)))
        self.resblocks = nn.ModuleList()
        for i in range(self.num_upsamples):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(ConvBlock(None, ch, k, d, hp.resblock))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=False)
        self.ups.apply(init_weights)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, c0, 1)

    def forward(self, x, g=None):
        if g is None:
            return self._forward_native(x)
        if self.gin_channels == 0:
            raise AttributeError("'HiFiGAN_vits' object has no attribute 'cond'")   # what the reference raises (:430)
        if g.dim() == 3:
            if g.shape[2] != 1:
                raise NotImplementedError("amphion_b200: time-varying conditioning g[B, gin, T] is not on this path "
                                          "(VITS passes a per-utterance embedding g[B, gin, 1])")
            g = g[:, :, 0]
        if g.dim() != 2 or g.shape[0] != x.shape[0] or g.shape[1] != self.gin_channels:
            raise ValueError(f"expected g of shape [{x.shape[0]}, {self.gin_channels}, 1], got {tuple(g.shape)}")
        _capi_g = g.to(device=x.device, dtype=torch.float32)
        if _capi_g.stride(1) != 1:
            _capi_g = _capi_g.contiguous()
        return self._forward_native(x, _capi_g)

    def remove_weight_norm(self):
        for l in self.ups:
            remove_weight_norm(l)
        for l in self.resblocks:
            l.remove_weight_norm()
