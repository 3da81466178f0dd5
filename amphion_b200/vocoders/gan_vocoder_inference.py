"""Function boundary of the GAN vocoders, signatures as in the reference's
``models/vocoders/gan/gan_vocoder_inference.py`` (vocoder_inference :11,
synthesis_audios :41)."""
from __future__ import annotations

import torch

from ..util import pad_f0_to_tensors, pad_mels_to_tensors


def vocoder_inference(cfg, model, mels, f0s=None, device=None, fast_inference=False):
    """mels [B, n_mel, T] (any device / strides) -> audios [B, T*hop] on the CPU,
    detached fp32 — the reference's output contract (:38).  ``fast_inference``
    is accepted and ignored, as in the reference."""
    if getattr(cfg.preprocess, "extract_amplitude_phase", False):
        raise NotImplementedError("amphion_b200: amplitude/phase generators (APNet) are not on this path yet")
    model.eval()
    with torch.no_grad():
        if device is None:
            device = next(model.parameters()).device
        mels = mels.to(device, non_blocking=True)
        if f0s is None:
            output = model.forward(mels)
        else:  # f0-conditioned generators (NSF-HiFiGAN), :36
            output = model.forward(mels, f0s.to(device, non_blocking=True))
        output = output.squeeze(1).detach()
        # the reference's `.cpu()` (:38) lands in pageable memory; a pinned buffer is
        # still a CPU tensor and lets the D2H copy run at full PCIe rate
        host = torch.empty(output.shape, dtype=output.dtype, pin_memory=True)
        host.copy_(output, non_blocking=True)
        torch.cuda.current_stream(output.device).synchronize()
        return host


def synthesis_audios(cfg, model, mels, f0s=None, batch_size=None, fast_inference=False):
    """List of ``[n_mel, T_i]`` mels -> list of 1-D CPU audios trimmed to
    ``T_i * hop``.  The reference pads to the batch maximum and then runs every
    utterance alone (B=1) on its zero-padded mel (:59-75); the generator has no
    cross-batch op, so running the padded batch in ONE forward gives the same
    samples with one launch sequence and one D2H copy per batch."""
    device = next(model.parameters()).device
    hop = model.cfg.preprocess.hop_size
    audios = []
    mel_batches, mel_frames = pad_mels_to_tensors([m.to(device) for m in mels], batch_size)
    # f0 tracks are padded per batch exactly like the mels (:55-56, :76-95)
    f0_batches = pad_f0_to_tensors([f.to(device) for f in f0s], batch_size) if f0s is not None else [None] * len(mel_batches)
    for mel_batch, mel_frame, f0_batch in zip(mel_batches, mel_frames, f0_batches):
        out = vocoder_inference(cfg, model, mel_batch, f0s=f0_batch, device=device, fast_inference=fast_inference)
        for i in range(mel_batch.shape[0]):
            audios.append(out[i, : int(mel_frame[i]) * hop].clone())
    return audios
