"""Batching helpers with the reference's signatures (utils/util.py:114-182)."""
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
