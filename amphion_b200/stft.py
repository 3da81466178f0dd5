"""``STFT`` / ``TacotronSTFT`` with the reference's constructor, buffers and
``mel_spectrogram`` contract (utils/stft.py:115-278).  The reference evaluates
the STFT as a dense convolution with a windowed DFT matrix on the GPU
(:167-172); here the same transform is the cuFFT-driven native pipeline."""
from __future__ import annotations

import os

import torch

from . import mel as _mel


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    return torch.exp(x) / C


class STFT(torch.nn.Module):
    def __init__(self, filter_length, hop_length, win_length, window="hann"):
        super().__init__()
        if window != "hann":
            raise NotImplementedError("amphion_b200: only the hann window of the reference's call sites")
        assert filter_length >= win_length
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window
        cutoff = filter_length // 2 + 1
        fb = torch.fft.fft(torch.eye(filter_length, dtype=torch.float64))[:cutoff]
        basis = torch.cat([fb.real, fb.imag], 0)
        win = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
        lpad = (filter_length - win_length) // 2
        full = torch.zeros(filter_length, dtype=torch.float64)
        full[lpad:lpad + win_length] = win
        # buffer kept for state-dict / attribute compatibility (utils/stft.py:149)
        self.register_buffer("forward_basis", (basis.float() * full.float())[:, None, :])
        self.register_buffer("fft_window", win.float(), persistent=False)

    def transform(self, input_data):
        """[B, T] -> (magnitude [B, bins, F], phase unavailable -> None), CPU tensors like the reference (:172)."""
        y = input_data if input_data.is_cuda else input_data.cuda()
        mag, _, _ = _mel.native_stft_mel(y, self.filter_length, self.hop_length, self.win_length, self.fft_window,
                                         None, self.filter_length // 2, 0.0, want_mag=True, want_mel=False)
        return mag.cpu(), None


class TacotronSTFT(torch.nn.Module):
    def __init__(self, filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax):
        super().__init__()
        self.n_mel_channels = n_mel_channels
        self.sampling_rate = sampling_rate
        self.stft_fn = STFT(filter_length, hop_length, win_length)
        self.register_buffer("mel_basis", _mel.librosa_mel_fn(sampling_rate, filter_length, n_mel_channels,
                                                              mel_fmin, mel_fmax))

    def spectral_normalize(self, magnitudes):
        return dynamic_range_compression(magnitudes)

    def spectral_de_normalize(self, magnitudes):
        return dynamic_range_decompression(magnitudes)

    def mel_spectrogram(self, y):
        """y [B, T] in [-1, 1] -> (mel [B, n_mel, F], energy [B, F]), CPU tensors
        (the reference moves the conv result to the host, utils/stft.py:172)."""
        assert torch.min(y.data) >= -1
        assert torch.max(y.data) <= 1
        yc = y if y.is_cuda else y.cuda()   # the reference hard-codes .cuda() (:168)
        s = self.stft_fn
        # one fused kernel (own FFT) when n_fft = 1024; AMPHION_B200_MEL=cufft keeps the cuFFT pipeline.  The
        # reference computes this spectrum with a conv-DFT (:152-181), so there is no FFT backend to be identical to.
        fused = os.environ.get("AMPHION_B200_MEL", "fused") != "cufft"
        _, mel, energy = _mel.native_stft_mel(yc, s.filter_length, s.hop_length, s.win_length, s.fft_window,
                                              self.mel_basis, s.filter_length // 2, 0.0, want_energy=True, fused=fused)
        return mel.cpu(), energy.cpu()
