"""Batching helpers with the reference's signatures (utils/util.py:61-182)."""
from __future__ import annotations

import torch


def pad_mels_to_tensors(mels, batched=None):
    """Zero-pad a list of ``[n_mel, T_i]`` mels into batches of ``batched``
    utterances (``None`` = one batch).  Returns ``(tensors, mel_frames)`` like
    the reference; unlike the reference (which always builds CPU tensors and
    copies device inputs back to the host, utils/util.py:151-178) the padded
    batch stays on the device the mels live on.
    """
    tensors, mel_frames = [], []
    if len(mels) == 0:
        return tensors, mel_frames
    step = len(mels) if batched is None else int(batched)
    for start in range(0, len(mels), step):
        group = mels[start:start + step]
        size = max(int(m.shape[-1]) for m in group)
        t = torch.zeros(len(group), group[0].shape[0], size, dtype=torch.float32, device=group[0].device)
        frames = torch.zeros(len(group), dtype=torch.int32)
        for i, m in enumerate(group):
            t[i, :, : m.shape[-1]] = m
            frames[i] = m.shape[-1]
        tensors.append(t)
        mel_frames.append(frames)
    return tensors, mel_frames


def pad_f0_to_tensors(f0s, batched=None):
    """Zero-pad a list of 1-D f0 tracks ``[T_i]`` into ``[b, T_max]`` batches of ``batched``
    utterances (``None`` = one batch) — utils/util.py:61-111; the batches stay on the device of
    the inputs (the reference builds CPU tensors)."""
    tensors = []
    if len(f0s) == 0:
        return tensors
    step = len(f0s) if batched is None else int(batched)
    for start in range(0, len(f0s), step):
        group = f0s[start:start + step]
        size = max(int(f.shape[-1]) for f in group)
        t = torch.zeros(len(group), size, dtype=torch.float32, device=group[0].device)
        for i, f in enumerate(group):
            t[i, : f.shape[-1]] = f
        tensors.append(t)
    return tensors
