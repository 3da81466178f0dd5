"""HiFi-GAN generator — drop-in for the reference class
``models/vocoders/gan/generator/hifigan.py:151`` (``_vocoders["hifigan"]``,
models/vocoders/vocoder_inference.py:39-49): same constructor (``cfg``), same
parameter names, ``forward(mel[B,n_mel,T]) -> wav[B,1,T*hop]``, ``remove_weight_norm``.
"""
from __future__ import annotations

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
