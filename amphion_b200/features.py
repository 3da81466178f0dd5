"""On-disk feature format either side of the vocoder path.

The reference stores one ``.npy`` per utterance and feature under
``<processed_dir>/<dataset>/<feature_dir>/<uid>.npy`` (``utils/io.py:12-30``; mel ``[n_mel, T]`` float32,
``processors/acoustic_extractor.py:397-401``), lists the utterances in ``<processed_dir>/<dataset>/{train,valid}.json``
(``config/base.json:62-90``) and reads them back in ``models/vocoders/vocoder_dataset.py:20-165``.  This module
keeps those layouts and names: mels are extracted on the GPU by the native front end, written in the reference's
format, read back by a dataset with the reference's ``__getitem__`` contract, and turned into audio files by the
batched synthesis path."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import mel as _mel


def save_feature(process_dir, feature_dir, item, feature, overrides=True):
    """utils/io.py:12-30: ``<process_dir>/<feature_dir>/<item>.npy``; an existing file is kept unless ``overrides``."""
    process_dir = os.path.join(process_dir, feature_dir)
    os.makedirs(process_dir, exist_ok=True)
    out_path = os.path.join(process_dir, item + ".npy")
    if os.path.exists(out_path) and not overrides:
        return out_path
    np.save(out_path, feature.detach().cpu().numpy() if torch.is_tensor(feature) else feature)
    return out_path


def extract_utt_mel_features(dataset_output, cfg, utt, wav):
    """The mel, from-mel energy and amplitude/phase part of ``extract_utt_acoustic_features_vocoder``
    (processors/acoustic_extractor.py:376-439) for one utterance whose audio ``wav`` [T] is already a CUDA tensor at
    ``cfg.preprocess.sample_rate`` (loading / resampling stays with the caller).  Returns the mel [n_mel, T]."""
    uid = utt["Uid"]
    pre = cfg.preprocess
    with torch.no_grad():
        mel = _mel.extract_mel_features(wav.unsqueeze(0), pre)        # [n_mel, T] after the reference's squeeze(0)
    if getattr(pre, "extract_mel", True):
        save_feature(dataset_output, getattr(pre, "mel_dir", "mels"), uid, mel)
    if getattr(pre, "extract_energy", False) and getattr(pre, "energy_extract_mode", "from_mel") == "from_mel":
        energy = (mel.exp() ** 2).sum(0).sqrt()                        # acoustic_extractor.py:408
        save_feature(dataset_output, getattr(pre, "energy_dir", "energys"), uid, energy)
    if getattr(pre, "extract_amplitude_phase", False):                 # acoustic_extractor.py:428-439 (APNet features)
        log_amplitude, phase, real, imaginary = _mel.amplitude_phase_spectrum(wav.unsqueeze(0), pre)
        save_feature(dataset_output, getattr(pre, "log_amplitude_dir", "log_amplitudes"), uid, log_amplitude)
        save_feature(dataset_output, getattr(pre, "phase_dir", "phases"), uid, phase)
        save_feature(dataset_output, getattr(pre, "real_dir", "reals"), uid, real)
        save_feature(dataset_output, getattr(pre, "imaginary_dir", "imaginarys"), uid, imaginary)
    return mel


def write_metadata(processed_dir, dataset, utts, file_name):
    """``<processed_dir>/<dataset>/<file_name>``: the JSON list of ``{"Dataset", "Uid", ...}`` records the datasets read."""
    d = os.path.join(processed_dir, dataset)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, file_name)
    with open(path, "w", encoding="utf-8") as f:
        json.dump(list(utts), f, indent=4, ensure_ascii=False)
    return path


class VocoderDataset(torch.utils.data.Dataset):
    """models/vocoders/vocoder_dataset.py:20-197, mel part: ``__getitem__`` returns ``{"mel": [n_mel, T] ndarray,
    "target_len": T}`` (+ ``"frame_pitch"`` aligned to ``target_len`` when ``use_frame_pitch``)."""

    def __init__(self, cfg, dataset, is_valid=False):
        assert isinstance(dataset, str)
        pre = cfg.preprocess
        processed_data_dir = os.path.join(pre.processed_dir, dataset)
        meta_file = pre.valid_file if is_valid else pre.train_file
        self.metafile_path = os.path.join(processed_data_dir, meta_file)
        self.metadata = self.get_metadata()
        self.data_root = processed_data_dir
        self.cfg = cfg
        self.utt2mel_path, self.utt2frame_pitch_path = {}, {}
        for utt_info in self.metadata:
            ds, uid = utt_info["Dataset"], utt_info["Uid"]
            utt = "{}_{}".format(ds, uid)
            if getattr(pre, "use_mel", True):
                self.utt2mel_path[utt] = os.path.join(pre.processed_dir, ds, getattr(pre, "mel_dir", "mels"), uid + ".npy")
            if getattr(pre, "use_frame_pitch", False):
                self.utt2frame_pitch_path[utt] = os.path.join(pre.processed_dir, ds, getattr(pre, "pitch_dir", "pitches"),
                                                              uid + ".npy")

    def get_metadata(self):
        with open(self.metafile_path, "r", encoding="utf-8") as f:
            return json.load(f)

    def get_dataset_name(self):
        return self.metadata[0]["Dataset"]

    def __len__(self):
        return len(self.metadata)

    def __getitem__(self, index):
        utt_info = self.metadata[index]
        utt = "{}_{}".format(utt_info["Dataset"], utt_info["Uid"])
        single_feature = {}
        if utt in self.utt2mel_path:
            mel = np.load(self.utt2mel_path[utt])
            assert mel.shape[0] == self.cfg.preprocess.n_mel       # [n_mels, T]
            single_feature.setdefault("target_len", mel.shape[1])
            single_feature["mel"] = mel
        if utt in self.utt2frame_pitch_path:
            pitch = np.load(self.utt2frame_pitch_path[utt])
            single_feature.setdefault("target_len", len(pitch))
            single_feature["frame_pitch"] = align_length(pitch, single_feature["target_len"])
        return single_feature


def align_length(feature, target_len, pad_value=0.0):
    """utils/data_utils.py:473-496: crop, or pad with ``pad_value``, the last axis to ``target_len``."""
    feature_len = feature.shape[-1]
    if feature.ndim == 2:
        if target_len > feature_len:
            return np.pad(feature, ((0, 0), (0, target_len - feature_len)), constant_values=pad_value)
        return feature[:, :target_len]
    if feature.ndim == 1:
        if target_len > feature_len:
            return np.pad(feature, (0, target_len - feature_len), constant_values=pad_value)
        return feature[:target_len]
    raise NotImplementedError


def synthesize_dataset(cfg, model, dataset, out_dir, batch_size=None, bucket=True, **save_kwargs):
    """``VocoderInference.inference`` over a feature directory (vocoder_inference.py:336-371): every utterance of
    ``dataset`` (a :class:`VocoderDataset`) -> ``<out_dir>/<uid>.wav``; batched forward, trim to ``target_len * hop``,
    PCM16 on the device, one D2H per batch.  Returns the paths in dataset order."""
    from .vocoders.gan_vocoder_inference import synthesize_to_files
    os.makedirs(out_dir, exist_ok=True)
    items = [dataset[i] for i in range(len(dataset))]
    mels = [torch.from_numpy(np.ascontiguousarray(it["mel"])) for it in items]
    f0s = [torch.from_numpy(np.ascontiguousarray(it["frame_pitch"])) for it in items] if items and "frame_pitch" in items[0] else None
    paths = [os.path.join(out_dir, info["Uid"] + ".wav") for info in dataset.metadata]
    synthesize_to_files(cfg, model, mels, paths, f0s=f0s, batch_size=batch_size, bucket=bucket, **save_kwargs)
    return paths
