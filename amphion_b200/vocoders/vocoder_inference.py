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


def _epoch_key(path):
    """Sort key of the reference (:447): the number after ``epoch-`` in ``epoch-0012_step-0004000_loss-0.51``."""
    name = os.path.basename(path.rstrip("/"))
    try:
        return int(name.split("_")[-3].split("-")[-1])
    except (IndexError, ValueError):
        digits = "".join(ch if ch.isdigit() else " " for ch in name).split()
        return int(digits[0]) if digits else -1


def resolve_checkpoint(weights_file):
    """The file a checkpoint argument stands for, resolved the way the reference does:
    a file is taken as is (legacy ``.pt``, :415-440); an experiment directory is searched under
    ``<dir>/checkpoint/`` for the entry with the highest epoch, ignoring names containing ``audio`` (:443-448);
    a directory that is itself one ``epoch-*_step-*`` checkpoint (vocoder_inference.py:277-278) or that directly
    holds the weights is used as is.  Inside, accelerate's ``pytorch_model.bin`` or ``model.safetensors``."""
    if not os.path.isdir(weights_file):
        return weights_file
    d = weights_file
    sub = os.path.join(d, "checkpoint")
    if os.path.isdir(sub):
        ls = [os.path.join(sub, n) for n in os.listdir(sub) if "audio" not in n]
        if not ls:
            raise FileNotFoundError(f"no checkpoint under {sub}")
        ls.sort(key=_epoch_key, reverse=True)
        d = ls[0]
    for name in ("pytorch_model.bin", "model.safetensors"):
        if os.path.exists(os.path.join(d, name)):
            return os.path.join(d, name)
    raise FileNotFoundError(f"neither pytorch_model.bin nor model.safetensors in {d}")


def _read_state_dict(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    ckpt = torch.load(path, map_location="cpu")
    if isinstance(ckpt, dict):
        for key in ("generator_state_dict", "state_dict"):
            if key in ckpt:
                return ckpt[key]
    return ckpt


def load_nnvocoder(cfg, vocoder_name, weights_file, from_multi_gpu=False):
    """Build ``_vocoders[vocoder_name](cfg)`` and load a checkpoint: a legacy ``.pt`` holding
    ``generator_state_dict`` (optionally with a ``module.`` prefix, reference :415-440) or an accelerate experiment
    directory (:443-451, see ``resolve_checkpoint``).  ``from_multi_gpu`` strips ``module.`` and, like the reference
    (:424-437), keeps only entries whose name and shape match the model.  The model is placed on the current CUDA
    device and set to eval."""
    print("Loading Vocoder from Weights file: {}".format(weights_file))
    model = _vocoders[vocoder_name](cfg)
    sd = _read_state_dict(resolve_checkpoint(weights_file))
    if from_multi_gpu:
        own = model.state_dict()
        picked = {k.split("module.")[-1]: v for k, v in sd.items()
                  if k.split("module.")[-1] in own and v.shape == own[k.split("module.")[-1]].shape}
        own.update(picked)
        sd = own
    elif any(k.startswith("module.") for k in sd):
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
    st = os.stat(resolve_checkpoint(vocoder_weight_file))
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
