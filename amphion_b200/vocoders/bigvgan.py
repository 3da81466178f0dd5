"""BigVGAN generator — drop-in for ``models/vocoders/gan/generator/bigvgan.py:232``
(``_vocoders["bigvgan"]``): same constructor, parameter names (``ups.{i}.0.*``,
``resblocks.{n}.activations.{a}.act.{alpha,beta}``, filter buffers), forward
contract and ``remove_weight_norm``.
"""
from __future__ import annotations

from torch import nn
from torch.nn.utils import remove_weight_norm, weight_norm

from .activations import Activation1d, Snake, SnakeBeta
from .generator import ConvBlock, NativeGenerator, init_weights


class BigVGAN(NativeGenerator):
    kind = "bigvgan"
    hp_key = "bigvgan"

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        hp = cfg.model.bigvgan
        self.num_kernels = len(hp.resblock_kernel_sizes)
        self.num_upsamples = len(hp.upsample_rates)
        c0 = hp.upsample_initial_channel
        if hp.activation == "snake":
            act_cls = Snake
        elif hp.activation == "snakebeta":
            act_cls = SnakeBeta
        else:
            raise NotImplementedError(
                "activation incorrectly specified. check the config file and look for 'activation'.")

        def make_activation(ch):
            return Activation1d(activation=act_cls(ch, alpha_logscale=hp.snake_logscale))

        self.conv_pre = weight_norm(nn.Conv1d(cfg.preprocess.n_mel, c0, 7, 1, padding=3))
        self.ups = nn.ModuleList(
            nn.ModuleList([weight_norm(nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u,
                                                          padding=(k - u) // 2))])
            for i, (u, k) in enumerate(zip(hp.upsample_rates, hp.upsample_kernel_sizes)))
        self.resblocks = nn.ModuleList()
        for i in range(self.num_upsamples):
            ch = c0 // (2 ** (i + 1))
            for k, d in zip(hp.resblock_kernel_sizes, hp.resblock_dilation_sizes):
                self.resblocks.append(ConvBlock(cfg, ch, k, d, hp.resblock, make_activation))
        self.activation_post = make_activation(ch)
        self.conv_post = weight_norm(nn.Conv1d(ch, 1, 7, 1, padding=3))
        for up in self.ups:
            up.apply(init_weights)
        self.conv_post.apply(init_weights)

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for l in self.ups:
            for l_i in l:
                remove_weight_norm(l_i)
        for l in self.resblocks:
            l.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)
