"""``synthesis()`` / ``load_nnvocoder()`` with the reference's signatures
(models/vocoders/vocoder_inference.py:397-515): the call every TTS/SVC recipe
ends with."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _vocoder_infer_funcs, _vocoders

_model_cache = {}


def _strip_module_prefix(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def load_nnvocoder(cfg, vocoder_name, weights_file, from_multi_gpu=False):
    """Build ``_vocoders[vocoder_name](cfg)`` and load a checkpoint: a legacy
    ``.pt`` holding ``generator_state_dict`` (optionally with a ``module.``
    prefix, reference :415-440) or an accelerate directory / ``pytorch_model.bin``
    (:291-294).  The model is placed on the current CUDA device and set to eval."""
    print("Loading Vocoder from Weights file: {}".format(weights_file))
    model = _vocoders[vocoder_name](cfg)
    path = weights_file
    if os.path.isdir(path):
        path = os.path.join(path, "pytorch_model.bin")
    ckpt = torch.load(path, map_location="cpu")
    sd = ckpt.get("generator_state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if from_multi_gpu or any(k.startswith("module.") for k in sd):
        sd = _strip_module_prefix(sd)
    model.load_state_dict(sd)
    if not torch.cuda.is_available():
        raise RuntimeError("amphion_b200: no CUDA device; there is no CPU fallback")
    return model.cuda().eval()


def tensorize(data, device, n_samples):
    """data: a list of numpy arrays (reference :460-468)."""
    assert type(data) == list
    if n_samples:
        data = data[:n_samples]
    return [torch.as_tensor(x).to(device) for x in data]


def synthesis(cfg, vocoder_weight_file, n_samples, pred, f0s=None, batch_size=64, fast_inference=False):
    """pred: list of ``[T, n_mel]`` numpy mels -> list of 1-D CPU audios.
    The reference rebuilds and reloads the vocoder on every call (:498-500);
    here the loaded model is cached per (file, mtime)."""
    vocoder_name = cfg.model.generator
    st = os.stat(os.path.join(vocoder_weight_file, "pytorch_model.bin")
                 if os.path.isdir(vocoder_weight_file) else vocoder_weight_file)
    key = (os.path.abspath(vocoder_weight_file), st.st_mtime_ns, vocoder_name)
    if key not in _model_cache:
        _model_cache.clear()
        _model_cache[key] = load_nnvocoder(cfg, vocoder_name, vocoder_weight_file)
    vocoder = _model_cache[key]
    device = next(vocoder.parameters()).device
    mels_pred = tensorize([np.asarray(p).T for p in pred], device, n_samples)
    print("For predicted mels, #sample = {}...".format(len(mels_pred)))
    return _vocoder_infer_funcs[vocoder_name](cfg, vocoder, mels_pred, f0s=f0s, batch_size=batch_size,
                                              fast_inference=fast_inference)
