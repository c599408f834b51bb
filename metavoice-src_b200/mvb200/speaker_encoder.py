"""Host-side mirror of ``fam/quantiser/audio/speaker_encoder/model.py::SpeakerEncoder`` over libmvb200.

Same constructor keywords, ``embed_utterance`` / ``embed_utterance_from_file`` / ``compute_partial_slices`` contract and
``speaker_encoder.pt`` layout (``{"model_state": {lstm.*, linear.*}}``, model.py:45-46).  The mel front-end, the 3-layer
LSTM, the projection and the partial-window averaging run on the device (csrc/speaker.cu); reading / resampling /
trimming the reference file stays on the host (the reference uses librosa for those: librosa.load at 16 kHz +
librosa.effects.trim(top_db=20), model.py:112-113 -- restated here with scipy's polyphase resampler, not pinned).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import struct
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import _lib

mel_window_length = 25
mel_window_step = 10
mel_n_channels = 40
sampling_rate = 16000
partials_n_frames = 160
model_hidden_size = 256
model_embedding_size = 256
model_num_layers = 3


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin, log0, step = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    return np.where(f >= log0, log0 / lin + np.log(np.maximum(f, 1e-10) / log0) / step, f / lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin, log0, step = 200.0 / 3, 1000.0, np.log(6.4) / 27.0
    return np.where(m >= log0 / lin, log0 * np.exp(step * (m - log0 / lin)), lin * m)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """Triangular filters on the Slaney mel scale with area normalisation ([n_mels, n_fft/2 + 1], fp32): what
    librosa.filters.mel(sr, n_fft, n_mels) returns with its defaults (fmin 0, fmax sr/2, htk False, norm "slaney")."""
    freqs = np.linspace(0, sr / 2, 1 + n_fft // 2)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2), n_mels + 2))
    width = np.diff(edges)
    fb = np.zeros((n_mels, freqs.size))
    for i in range(n_mels):
        up = (freqs - edges[i]) / width[i]
        down = (edges[i + 2] - freqs) / width[i + 1]
        fb[i] = np.clip(np.minimum(up, down), 0, None) * (2.0 / (edges[i + 2] - edges[i]))
    return fb.astype(np.float32)


def read_wav(path: str):
    """PCM16 / PCM32 / float32 RIFF wav -> (float32 mono [-1, 1], sample_rate).  Other containers (mp3, flac) need a
    decoder that is not part of this engine."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise NotImplementedError(f"{path}: only RIFF/WAVE speaker references are decoded here (mp3/flac need ffmpeg or librosa)")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: malformed wav")
    tag, ch, sr, _, _, bits = fmt
    if tag == 1 and bits == 16:
        x = np.frombuffer(pcm, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 32:
        x = np.frombuffer(pcm, dtype="<i4").astype(np.float32) / 2147483648.0
    elif tag == 3 and bits == 32:
        x = np.frombuffer(pcm, dtype="<f4").astype(np.float32)
    else:
        raise NotImplementedError(f"{path}: wav format tag {tag} / {bits} bit")
    if ch > 1:
        x = x.reshape(-1, ch).mean(axis=1)          # librosa.load(mono=True)
    return x, sr


def resample(x: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    if sr_in == sr_out:
        return x.astype(np.float32)
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(sr_in, sr_out)
    return resample_poly(x.astype(np.float64), sr_out // g, sr_in // g).astype(np.float32)


def trim_silence(y: np.ndarray, top_db: float = 20.0, frame_length: int = 2048, hop_length: int = 512) -> np.ndarray:
    """librosa.effects.trim(y, top_db): keep the span between the first and last frame whose RMS is within top_db of the
    loudest frame (centered frames, zero padded)."""
    if y.size == 0:
        return y
    pad = np.pad(y.astype(np.float64), frame_length // 2)
    n = 1 + (pad.size - frame_length) // hop_length
    idx = np.arange(frame_length)[None, :] + hop_length * np.arange(n)[:, None]
    rms = np.sqrt(np.mean(pad[idx] ** 2, axis=1))
    ref = rms.max()
    if ref <= 0:
        return y[:0]
    db = 20.0 * np.log10(np.maximum(rms, 1e-10) / ref)
    keep = np.nonzero(db > -top_db)[0]
    if keep.size == 0:
        return y[:0]
    return y[keep[0] * hop_length:min(y.size, (keep[-1] + 1) * hop_length)]


class SpeakerEncoder:
    """model.py:21-117 with the arithmetic in libmvb200."""

    def __init__(self, weights_fpath: Optional[str] = None, device: Optional[Union[str, torch.device]] = None, verbose: bool = True,
                 eval: bool = False, model_state: Optional[Dict[str, torch.Tensor]] = None, max_seconds: float = 180.0):
        if device is None:
            device = "cuda"
        self.device = torch.device(device)
        if model_state is None:
            model_state = torch.load(weights_fpath, map_location="cpu", weights_only=False)["model_state"]
        g = lambda k: model_state[k].detach().float().contiguous()
        tensors = []
        for l in range(model_num_layers):
            tensors += [g(f"lstm.weight_ih_l{l}"), g(f"lstm.weight_hh_l{l}"), g(f"lstm.bias_ih_l{l}") + g(f"lstm.bias_hh_l{l}")]
        n_fft = int(sampling_rate * mel_window_length / 1000)
        hop = int(sampling_rate * mel_window_step / 1000)
        i = np.arange(n_fft)
        tensors += [g("linear.weight"), g("linear.bias"),
                    torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, mel_n_channels)),
                    torch.from_numpy((0.5 - 0.5 * np.cos(2 * np.pi * i / n_fft)).astype(np.float32)),     # periodic Hann
                    torch.from_numpy(np.cos(2 * np.pi * i / n_fft).astype(np.float32)),
                    torch.from_numpy((-np.sin(2 * np.pi * i / n_fft)).astype(np.float32))]
        offs, total = [], 0
        for t in tensors:
            offs.append(total)
            total = (total + t.numel() * 4 + 255) // 256 * 256
        arena = torch.zeros(total, dtype=torch.uint8)
        for t, o in zip(tensors, offs):
            arena[o:o + t.numel() * 4] = t.contiguous().reshape(-1).view(torch.uint8)
        self._arena = arena.to(self.device)
        cfg = _lib.SpkConfig(mel_n_channels, model_hidden_size, model_num_layers, model_embedding_size, n_fft, hop,
                             partials_n_frames, int(max_seconds * sampling_rate))
        self.cfg, self._lib = cfg, _lib.load()
        wsb = self._lib.mvb_spk_workspace_bytes(C.byref(cfg))
        if wsb == 0:
            _lib.check(_lib.MVB_ERR_UNSUPPORTED)
        self._ws = torch.zeros(wsb, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mvb_spk_create(C.byref(cfg), self._arena.data_ptr(), self._arena.numel(),
                                                (C.c_uint64 * len(offs))(*offs), self._ws.data_ptr(), C.byref(h)))
        self._h = h
        if verbose:
            print(f"Loaded the speaker embedding model on {self.device.type}.")

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._lib.mvb_spk_destroy(self._h)
            self._h = None

    __del__ = close

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def compute_partial_slices(n_samples: int, rate, min_coverage):
        """model.py:55-79: windows of 160 mel frames every round(16000 / rate / 160) frames; the last one is dropped when
        it covers less than ``min_coverage`` of real audio."""
        spf = int(sampling_rate * mel_window_step / 1000)
        n_frames = int(np.ceil((n_samples + 1) / spf))
        step = int(np.round((sampling_rate / rate) / spf))
        starts = list(range(0, max(1, n_frames - partials_n_frames + step + 1), step))
        mel_slices = [slice(i, i + partials_n_frames) for i in starts]
        wav_slices = [slice(i * spf, (i + partials_n_frames) * spf) for i in starts]
        last = wav_slices[-1]
        if (n_samples - last.start) / (last.stop - last.start) < min_coverage and len(mel_slices) > 1:
            mel_slices, wav_slices = mel_slices[:-1], wav_slices[:-1]
        return wav_slices, mel_slices

    @torch.no_grad()
    def wav_to_mel_spectrogram(self, wav: np.ndarray) -> torch.Tensor:
        """audio.py:10-22 on the device: fp32 [n_frames, 40]."""
        w = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32)).to(self.device)
        mel = torch.empty((1 + w.numel() // self.cfg.hop, mel_n_channels), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.mvb_spk_mel(self._h, w.data_ptr(), w.numel(), mel.data_ptr(), self._st()))
        return mel

    @torch.no_grad()
    def embed_utterance(self, wav: np.ndarray, return_partials=False, rate=1.3, min_coverage=0.75, numpy: bool = True):
        wav_slices, mel_slices = self.compute_partial_slices(len(wav), rate, min_coverage)
        max_wave_length = wav_slices[-1].stop
        if max_wave_length >= len(wav):
            wav = np.pad(wav, (0, max_wave_length - len(wav)), "constant")
        if len(wav) > self.cfg.max_samples:
            raise ValueError(f"speaker reference longer than {self.cfg.max_samples / sampling_rate:.0f} s")
        w = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32)).to(self.device)
        starts = np.asarray([s.start for s in mel_slices], dtype=np.int32)
        embed = torch.empty(model_embedding_size, dtype=torch.float32, device=self.device)
        partials = torch.empty((len(starts), model_embedding_size), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.mvb_spk_embed(self._h, w.data_ptr(), w.numel(), starts.ctypes.data_as(C.c_void_p), len(starts),
                                           embed.data_ptr(), partials.data_ptr(), self._st()))
        if numpy:
            embed_o, partials_o = embed.cpu().numpy(), partials.cpu().numpy()
        else:
            embed_o, partials_o = embed, partials
        if return_partials:
            return embed_o, partials_o, wav_slices
        return embed_o

    def embed_utterance_from_file(self, fpath: str, numpy: bool) -> torch.Tensor:
        wav, sr = read_wav(fpath)                             # librosa.load(fpath, sr=16000)
        wav = trim_silence(resample(wav, sr, sampling_rate), top_db=20)
        return self.embed_utterance(wav, numpy=numpy)


def check_audio_file(path: str, threshold_s: float = 30.0):
    """fam/llm/utils.py:55-70: the speaker reference must hold at least 30 s of audio."""
    wav, sr = read_wav(path)
    if wav.size / sr < threshold_s:
        raise Exception(f"The audio file is too short. Please provide an audio file that is at least {threshold_s} seconds long to proceed.")


def get_cached_embedding(local_file_path: str, spkemb_model: SpeakerEncoder) -> torch.Tensor:
    """fam/llm/inference.py:419-435: disk cache ~/.cache/fam/embedding_<md5(path)>.pt keyed by the path string."""
    if not os.path.exists(local_file_path):
        raise FileNotFoundError(f"File {local_file_path} not found!")
    name = "embedding_" + hashlib.md5(local_file_path.encode("utf-8")).hexdigest() + ".pt"
    os.makedirs(os.path.expanduser("~/.cache/fam/"), exist_ok=True)
    cache_path = os.path.expanduser(f"~/.cache/fam/{name}")
    if not os.path.exists(cache_path):
        spk_emb = spkemb_model.embed_utterance_from_file(local_file_path, numpy=False).unsqueeze(0).cpu()
        torch.save(spk_emb, cache_path)
    else:
        spk_emb = torch.load(cache_path, map_location="cpu")
    return spk_emb
