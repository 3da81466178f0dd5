"""Mel-spectrogram front end with the reference's signatures (utils/mel.py):
``extract_linear_features`` :20, ``mel_spectrogram_torch`` :55,
``extract_mel_features`` :111, ``extract_mel_features_tts`` :173.
``cfg`` is ``cfg.preprocess`` (sample_rate, n_fft, n_mel, fmin, fmax, win_size, hop_size).
All arithmetic runs in the CUDA pipeline behind ``ab_mel_forward``."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _capi

# module-level caches are part of the reference's visible surface (utils/mel.py:107-108)
mel_basis = {}
hann_window = {}
_handles = {}
_workspaces = {}


def librosa_mel_fn(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """Slaney-scale, Slaney-normalised triangular filterbank [n_mels, n_fft//2+1]
    float32 — the published algorithm of librosa.filters.mel (htk=False,
    norm='slaney'), which the reference imports at utils/mel.py:7."""
    if fmax is None:
        fmax = sr / 2.0
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    mels = torch.linspace(to_mel(float(fmin)), to_mel(float(fmax)), n_mels + 2, dtype=torch.float64)
    hz = torch.where(mels >= min_log_mel, min_log_hz * torch.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    fft_f = torch.linspace(0, sr / 2.0, 1 + n_fft // 2, dtype=torch.float64)
    ramps = hz[:, None] - fft_f[None, :]
    fdiff = hz[1:] - hz[:-1]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp(torch.minimum(lower, upper), min=0).to(torch.float32)
    enorm = (2.0 / (hz[2:] - hz[:-2])).to(torch.float64)
    return (w * enorm[:, None].to(torch.float32)).to(torch.float32)


def _basis_for(cfg, device):
    key = (cfg.sample_rate, cfg.n_fft, cfg.n_mel, cfg.fmin, cfg.fmax, str(device))
    if key not in mel_basis:  # the reference's cache never hits; this one does
        mel_basis[key] = librosa_mel_fn(cfg.sample_rate, cfg.n_fft, cfg.n_mel, cfg.fmin, cfg.fmax).to(device)
    wkey = (cfg.win_size, str(device))
    if wkey not in hann_window:
        hann_window[wkey] = torch.hann_window(cfg.win_size).to(device)
    return mel_basis[key], hann_window[wkey]


def _warn_range(y):
    if torch.min(y) < -1.0:
        print("min value is ", torch.min(y))
    if torch.max(y) > 1.0:
        print("max value is ", torch.max(y))


def _handle_for(n_fft, hop, win, n_mel, pad, eps, clamp, device):
    # cuFFT plans belong to a device: one handle per (configuration, device)
    key = (n_fft, hop, win, n_mel, pad, float(eps), float(clamp), str(device))
    if key not in _handles:
        h = C.c_void_p()
        cfg = _capi.MelConfig(n_fft, hop, win, n_mel, pad, eps, clamp)
        _capi.check(_capi.lib.ab_mel_create(C.byref(cfg), C.byref(h)), "ab_mel_create")
        _handles[key] = h
    return _handles[key]


def _workspace(device, need):
    ws = _workspaces.get(str(device))
    if ws is None or ws.numel() < need + 256:
        ws = torch.empty(need + 256, dtype=torch.uint8, device=device)
        _workspaces[str(device)] = ws
    return (ws.data_ptr() + 255) // 256 * 256


class _LogMel(torch.autograd.Function):
    """log-mel with a native backward (``ab_mel_backward``): what the trainers' mel loss differentiates
    (gan_vocoder_trainer.py:368-396, ``extract_mel_features(y_pred.squeeze(1), cfg.preprocess)``)."""

    @staticmethod
    def forward(ctx, y, window, basis, n_fft, hop, win, pad, eps, clamp):
        y = y.contiguous().float()
        mel = _stft_mel_forward(y, n_fft, hop, win, window, basis, pad, eps, False, True, False, clamp, False)[1]
        ctx.save_for_backward(y, window, basis)
        ctx.geom = (n_fft, hop, win, pad, eps, clamp)
        return mel

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        y, window, basis = ctx.saved_tensors
        n_fft, hop, win, pad, eps, clamp = ctx.geom
        B, T = y.shape
        lib = _capi.lib
        h = _handle_for(n_fft, hop, win, int(basis.shape[0]), pad, eps, clamp, y.device)
        g = g.contiguous().float()
        gy = torch.empty_like(y)
        with torch.cuda.device(y.device):
            need = lib.ab_mel_backward_workspace_bytes(h, B, T)
            wbase = _workspace(y.device, need)
            window = window.to(device=y.device, dtype=torch.float32).contiguous()
            bt = basis.to(device=y.device, dtype=torch.float32).contiguous()
            _capi.check(lib.ab_mel_backward(h, _capi.ptr(y), B, T, _capi.ptr(window), _capi.ptr(bt), _capi.ptr(g),
                                            _capi.ptr(gy), C.c_void_p(wbase), need, _capi.stream_ptr()), "ab_mel_backward")
        return gy, None, None, None, None, None, None, None, None


def native_stft_mel(y, n_fft, hop, win, window, basis, pad, eps, want_mag=False, want_mel=True,
                    want_energy=False, clamp=1e-5, fused=False):
    """y [B, T] CUDA fp32 -> (magnitude [B,bins,F] | None, mel [B,n_mel,F] | None, energy [B,F] | None).

    ``fused=True`` asks for the one-kernel front end (own FFT, n_fft = 1024, no magnitude output): the wav is read
    once and only mel / energy are written.  The default is the cuFFT pipeline, whose spectrum is bit-identical to
    ``torch.stft`` on the same device.  When ``y`` requires grad the log-mel output is differentiable (native
    backward); magnitude and energy are returned detached."""
    _capi.require_cuda(y, "mel front end")
    if y.dim() != 2:
        raise ValueError("expected audio of shape [B, T]")
    if torch.is_grad_enabled() and y.requires_grad and want_mel and basis is not None:
        mel = _LogMel.apply(y, window, basis, n_fft, hop, win, pad, eps, clamp)
        mag, en = None, None
        if want_mag or want_energy:
            mag, _, en = _stft_mel_forward(y.detach(), n_fft, hop, win, window, None, pad, eps, want_mag, False,
                                           want_energy, clamp, False)
        return mag, mel, en
    return _stft_mel_forward(y.detach(), n_fft, hop, win, window, basis, pad, eps, want_mag, want_mel, want_energy,
                             clamp, fused)


def _stft_mel_forward(y, n_fft, hop, win, window, basis, pad, eps, want_mag, want_mel, want_energy, clamp, fused):
    y = y.contiguous().float()
    B, T = y.shape
    n_mel = int(basis.shape[0]) if (want_mel and basis is not None) else 0
    h = _handle_for(n_fft, hop, win, n_mel, pad, eps, clamp, y.device)
    lib = _capi.lib
    F = lib.ab_mel_num_frames(h, T)
    if F <= 0:
        raise ValueError(f"audio of {T} samples is too short for n_fft={n_fft}, pad={pad}")
    with torch.cuda.device(y.device):
        fused = bool(fused) and n_fft == 1024 and not want_mag and 0 < n_mel <= 128
        need = 4096 if fused else lib.ab_mel_workspace_bytes(h, B, T)
        wbase = _workspace(y.device, need)
        bins = n_fft // 2 + 1
        mag = torch.empty(B, bins, F, device=y.device) if want_mag else None
        mel = torch.empty(B, n_mel, F, device=y.device) if n_mel else None
        en = torch.empty(B, F, device=y.device) if want_energy else None
        window = window.to(device=y.device, dtype=torch.float32).contiguous()
        bt = basis.to(device=y.device, dtype=torch.float32).contiguous() if n_mel else None
        if fused:
            _capi.check(lib.ab_mel_forward_fused(h, _capi.ptr(y), B, T, _capi.ptr(window), _capi.ptr(bt), _capi.ptr(mel),
                                                 _capi.ptr(en), C.c_void_p(wbase), need, _capi.stream_ptr()),
                        "ab_mel_forward_fused")
        else:
            _capi.check(lib.ab_mel_forward(h, _capi.ptr(y), B, T, _capi.ptr(window), _capi.ptr(bt), _capi.ptr(mag),
                                           _capi.ptr(mel), _capi.ptr(en), C.c_void_p(wbase), need, _capi.stream_ptr()),
                        "ab_mel_forward")
    return mag, mel, en


def _prepad(y, cfg, center):
    """(n_fft-hop)/2 reflect pad is done inside the kernel; torch.stft(center=True)
    adds a second reflect pad of n_fft//2 on top of it (utils/mel.py:153-164)."""
    p1 = int((cfg.n_fft - cfg.hop_size) / 2)
    if not center:
        return y, p1
    y = torch.nn.functional.pad(y.unsqueeze(1), (p1, p1), mode="reflect").squeeze(1)
    return y, cfg.n_fft // 2


def extract_linear_features(y, cfg, center=False):
    _warn_range(y)
    _, win = _basis_for(cfg, y.device)
    y, pad = _prepad(y, cfg, center)
    mag, _, _ = native_stft_mel(y, cfg.n_fft, cfg.hop_size, cfg.win_size, win, None, pad, 1e-9,
                                want_mag=True, want_mel=False)
    return torch.squeeze(mag, 0)


def mel_spectrogram_torch(y, cfg, center=False):
    _warn_range(y)
    basis, win = _basis_for(cfg, y.device)
    y, pad = _prepad(y, cfg, center)
    return native_stft_mel(y, cfg.n_fft, cfg.hop_size, cfg.win_size, win, basis, pad, 1e-6)[1]


def extract_mel_features(y, cfg, center=False, **kwargs):
    """utils/mel.py:111-170.  ``taco=True, _stft=...`` kwargs (passed by the SVC
    feature path, processors/acoustic_extractor.py:148-150) route to
    the TacotronSTFT variant instead of raising TypeError."""
    if kwargs.get("taco", False):
        return extract_mel_features_tts(y, cfg, center=center, taco=True, _stft=kwargs.get("_stft"))
    _warn_range(y)
    basis, win = _basis_for(cfg, y.device)
    y, pad = _prepad(y, cfg, center)
    return native_stft_mel(y, cfg.n_fft, cfg.hop_size, cfg.win_size, win, basis, pad, 1e-9)[1].squeeze(0)


def amplitude_phase_spectrum(y, cfg):
    """utils/mel.py:244-280 (APNet's training features): y [B, T] -> (log_amplitude, phase, rea, imag), each
    [B, n_fft//2+1, frames] ([n_fft//2+1, frames] for B == 1, the reference's squeeze)."""
    _capi.require_cuda(y, "amplitude_phase_spectrum")
    if y.dim() != 2:
        raise ValueError("expected audio of shape [B, T]")
    y = y.detach().contiguous().float()
    B, T = y.shape
    pad = int((cfg.n_fft - cfg.hop_size) / 2)
    h = _handle_for(cfg.n_fft, cfg.hop_size, cfg.win_size, 0, pad, 0.0, 1e-5, y.device)
    lib = _capi.lib
    F = lib.ab_mel_num_frames(h, T)
    if F <= 0:
        raise ValueError(f"audio of {T} samples is too short for n_fft={cfg.n_fft}")
    wkey = (cfg.win_size, str(y.device))
    if wkey not in hann_window:
        hann_window[wkey] = torch.hann_window(cfg.win_size).to(y.device)
    window = hann_window[wkey]
    bins = cfg.n_fft // 2 + 1
    outs = [torch.empty(B, bins, F, device=y.device) for _ in range(4)]
    with torch.cuda.device(y.device):
        need = lib.ab_mel_workspace_bytes(h, B, T)
        wbase = _workspace(y.device, need)
        _capi.check(lib.ab_amplitude_phase_forward(h, _capi.ptr(y), B, T, _capi.ptr(window), *[_capi.ptr(o) for o in outs],
                                                   C.c_void_p(wbase), need, _capi.stream_ptr()), "ab_amplitude_phase_forward")
    if B == 1:
        outs = [o.squeeze(0) for o in outs]
    return tuple(outs)


def extract_mel_features_tts(y, cfg, center=False, taco=False, _stft=None):
    """utils/mel.py:173-241."""
    if not taco:
        return extract_mel_features(y, cfg, center=center)
    audio = torch.clip(y, -1, 1)
    spec, _energy = _stft.mel_spectrogram(audio)
    return spec.squeeze(0)
