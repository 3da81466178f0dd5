"""Audio output with the reference's signature (``utils/io.py:49-76``): ``save_audio(path, waveform, fs,
add_silence=False, turn_up=False, volume_peak=0.9)``.

The arithmetic (peak normalisation, 50 ms silence, float -> 16-bit PCM) runs on the GPU behind
``ab_pcm16_forward`` so a batch of generator outputs crosses PCIe as int16 (half the bytes of the reference's
fp32 ``.cpu()``); the file itself is a canonical RIFF/WAVE PCM_S 16 written with the standard library.
``save_audios`` is the batched form used after ``synthesis_audios``-style generation: one kernel sequence and
one D2H copy for the whole batch.
"""
from __future__ import annotations

import ctypes as C
import wave

import numpy as np
import torch

from . import _capi


def waveform_to_pcm16(wav, lengths=None, silence=0, turn_up=False, volume_peak=0.9):
    """wav [B, T] fp32 CUDA (row-contiguous) -> int16 CUDA [B, T + 2*silence (+1 to make it even)].
    ``lengths`` [B] valid samples per row (default: all T); row b holds ``silence`` zeros, its
    ``lengths[b]`` quantised samples, ``silence`` zeros, then zero padding."""
    _capi.require_cuda(wav, "waveform_to_pcm16")
    if wav.dim() != 2 or wav.dtype != torch.float32:
        raise ValueError(f"expected a float32 [B, T] waveform, got {tuple(wav.shape)} {wav.dtype}")
    if wav.stride(1) != 1:
        wav = wav.contiguous()
    B, T = wav.shape
    if B == 0 or T == 0:
        raise ValueError("amphion_b200: empty waveform batch")
    out_stride = T + 2 * int(silence)
    out_stride += out_stride & 1
    with torch.cuda.device(wav.device):
        out = torch.empty(B, out_stride, dtype=torch.int16, device=wav.device)
        dev_len = None
        if lengths is not None:
            dev_len = torch.as_tensor(lengths, dtype=torch.int64).to(wav.device)
            if dev_len.shape != (B,):
                raise ValueError(f"lengths must have shape [{B}], got {tuple(dev_len.shape)}")
        need = _capi.lib.ab_pcm16_workspace_bytes(B)
        ws = torch.empty(need, dtype=torch.uint8, device=wav.device)
        _capi.check(_capi.lib.ab_pcm16_forward(_capi.ptr(wav), B, T, wav.stride(0),
                                               _capi.ptr(dev_len) if dev_len is not None else None,
                                               int(bool(turn_up)), C.c_float(float(volume_peak)), int(silence),
                                               _capi.ptr(out), out_stride, _capi.ptr(ws), need, _capi.stream_ptr()),
                    "ab_pcm16_forward")
    return out


def write_wav_pcm16(path, samples, fs):
    """1-D int16 numpy array -> mono RIFF/WAVE PCM_S 16 (what torchaudio.save(..., encoding="PCM_S",
    bits_per_sample=16) stores, utils/io.py:76)."""
    samples = np.ascontiguousarray(samples, dtype="<i2")
    with wave.open(str(path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(fs))
        f.writeframes(samples.tobytes())


def save_audios(paths, wavs, fs, lengths=None, add_silence=False, turn_up=False, volume_peak=0.9):
    """Batched save_audio: ``wavs`` [B, T] fp32 (CUDA, or host -> copied to the current device), one file
    per row trimmed to ``lengths[b]`` samples (+ silence)."""
    if not torch.is_tensor(wavs):
        wavs = torch.as_tensor(np.asarray(wavs), dtype=torch.float32)
    if not wavs.is_cuda:
        wavs = wavs.to("cuda", dtype=torch.float32)
    B, T = wavs.shape
    if len(paths) != B:
        raise ValueError(f"{len(paths)} paths for {B} waveforms")
    silence = int(fs) // 20 if add_silence else 0
    lens = [T] * B if lengths is None else [int(n) for n in lengths]
    pcm = waveform_to_pcm16(wavs, lengths, silence, turn_up, volume_peak)
    host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    host.copy_(pcm, non_blocking=True)
    torch.cuda.current_stream(pcm.device).synchronize()
    arr = host.numpy()
    for b, p in enumerate(paths):
        write_wav_pcm16(p, arr[b, : lens[b] + 2 * silence], fs)


def save_audio(path, waveform, fs, add_silence=False, turn_up=False, volume_peak=0.9):
    """Drop-in for utils/io.py:49 — ``waveform``: 1-D (or [1, T]) numpy array / tensor on any device."""
    if not torch.is_tensor(waveform):
        waveform = torch.as_tensor(np.asarray(waveform))
    waveform = waveform.to(dtype=torch.float32)
    if waveform.dim() == 2 and waveform.shape[0] != 1:
        # stereo: the reference scales all channels by the global peak (:59-62), then averages them (:73-75);
        # its np.concatenate of 1-D silence with a 2-D array (:67) raises, so add_silence is refused as well
        if add_silence:
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
        waveform = waveform.to("cuda")
        if turn_up:
            waveform = waveform * (volume_peak / torch.maximum(waveform.max(), waveform.min().abs()))
        waveform, turn_up = waveform.mean(dim=0, keepdim=True), False
    waveform = waveform.reshape(1, -1)
    save_audios([path], waveform, fs, None, add_silence, turn_up, volume_peak)
