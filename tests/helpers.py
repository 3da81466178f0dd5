"""Shared helpers for the parity tests."""
from types import SimpleNamespace as NS

import numpy as np
import torch

from conftest import GOLDEN_MODELS, GOLDEN_NSF, load_golden

HP_V1 = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
             upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
             resblock_dilation_sizes=[[1, 3, 5]] * 3)
HP_BIGVGAN_BASE = dict(HP_V1, activation="snakebeta", snake_logscale=True)
# egs/vocoder/gan/nsfhifigan/exp_config.json:16-48
HP_NSF_EXP = dict(resblock="1", harmonic_num=8, upsample_rates=[8, 4, 2, 2, 2], upsample_kernel_sizes=[16, 8, 4, 4, 4],
                  upsample_initial_channel=768, resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3)


def make_cfg(kind, hp, n_mel):
    pre = NS(n_mel=n_mel, hop_size=int(np.prod(hp["upsample_rates"])), extract_amplitude_phase=False, sample_rate=24000)
    return NS(preprocess=pre, model=NS(generator=kind, **{kind: NS(**hp)}))


def build_model(kind, hp, n_mel, state_dict=None, seed=None):
    from amphion_b200.vocoders import _vocoders
    if seed is not None:
        torch.manual_seed(seed)
    m = _vocoders[kind](make_cfg(kind, hp, n_mel))
    if state_dict is not None:
        m.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()}, strict=True)
    return m.eval()


def golden_model(name):
    kind, hp, n_mel = GOLDEN_NSF if name == "nsfhifigan" else GOLDEN_MODELS[name]
    g, sd = load_golden(name)
    return kind, hp, g, sd, build_model(kind, hp, n_mel, sd)


def sd_numpy(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def randomize_snake(model, seed, logscale):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith(".alpha") or n.endswith(".beta"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3 + (0.0 if logscale else 1.0))
