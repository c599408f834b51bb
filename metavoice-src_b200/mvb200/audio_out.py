"""Host-side audio post-processing of the reference's `_save_audio` (fam/llm/decoders.py:40-47): it calls
`audiocraft.data.audio.audio_write(name, wav, 24000, strategy="loudness", loudness_compressor=True)`.

audiocraft==1.2.0 is not in this image, so this restates the published algorithm of
`audiocraft/data/audio_utils.py::normalize_loudness` / `normalize_audio` / `_clip_wav` and
`audiocraft/data/audio.py::audio_write` (parity unpinned: nothing in the reference's tests covers it); the loudness
meter itself is torchaudio's ITU-R BS.1770 implementation -- the very function audiocraft calls -- which IS installed.
SURVEY.md §8 row a18 keeps this step on the host (N2 moves it on device).
"""
from __future__ import annotations

import struct

import numpy as np
import torch

LOUDNESS_HEADROOM_DB = 14.0      # audio_write default `loudness_headroom_db`
ENERGY_FLOOR = 2e-3              # normalize_loudness default `energy_floor`


def normalize_loudness(wav: torch.Tensor, sample_rate: int, loudness_headroom_db: float = LOUDNESS_HEADROOM_DB,
                       loudness_compressor: bool = False, energy_floor: float = ENERGY_FLOOR) -> torch.Tensor:
    """wav [C, T] fp32 -> same shape, integrated loudness moved to -`loudness_headroom_db` LUFS, optional tanh
    compressor.  Signals below the energy floor are returned unchanged (as audiocraft does)."""
    import torchaudio
    energy = wav.pow(2).mean().sqrt().item()
    if energy < energy_floor:
        return wav
    input_loudness_db = torchaudio.functional.loudness(wav, sample_rate).item()
    delta_loudness = -loudness_headroom_db - input_loudness_db
    gain = 10.0 ** (delta_loudness / 20.0)
    output = gain * wav
    if loudness_compressor:
        output = torch.tanh(output)
    assert output.isfinite().all(), (input_loudness_db, energy)
    return output


def audio_write_wav(path_stem: str, wav: torch.Tensor, sample_rate: int, strategy: str = "loudness",
                    loudness_compressor: bool = True) -> str:
    """`audio_write(..., format="wav")`: normalise ([C, T] or [T] fp32 on any device), clip to [-1, 1], write PCM16.
    Returns the path (stem + ".wav")."""
    assert wav.dtype.is_floating_point, "wav is not floating point"
    wav = wav.detach().to("cpu", torch.float32)
    if wav.dim() == 1:
        wav = wav[None]
    elif wav.dim() > 2:
        raise ValueError("Input wav should be at most 2 dimension.")
    assert wav.isfinite().all()
    if strategy == "loudness":
        wav = normalize_loudness(wav, sample_rate, LOUDNESS_HEADROOM_DB, loudness_compressor)
    elif strategy != "clip":
        raise ValueError(f"unsupported normalisation strategy {strategy!r}")
    wav = wav.clamp(-1.0, 1.0)                                              # _clip_wav
    path = path_stem + ".wav"
    # interleaved little-endian PCM16 (torchaudio.save(..., encoding="PCM_S", bits_per_sample=16))
    pcm = (wav.t().contiguous().numpy() * 32767.0).round().astype("<i2").tobytes()
    ch = wav.shape[0]
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt "
                + struct.pack("<IHHIIHH", 16, 1, ch, sample_rate, sample_rate * 2 * ch, 2 * ch, 16))
        f.write(b"data" + struct.pack("<I", len(pcm)) + pcm)
    return path


def read_wav_pcm16(path: str):
    """Minimal reader for the files written above (tests)."""
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    ch, sr = struct.unpack("<H", b[22:24])[0], struct.unpack("<I", b[24:28])[0]
    n = struct.unpack("<I", b[40:44])[0]
    x = np.frombuffer(b[44:44 + n], dtype="<i2").reshape(-1, ch).T.astype(np.float32) / 32767.0
    return torch.from_numpy(x.copy()), sr


# ---- on-device variant: loudness meter, gain, compressor and PCM16 conversion in libmvb200 -----------
_POST_WS = {}


def wav_bytes_on_device(wav: torch.Tensor, sample_rate: int, loudness_headroom_db: float = LOUDNESS_HEADROOM_DB,
                        loudness_compressor: bool = True, return_stats: bool = False):
    """Device fp32 mono waveform [T] (or [1, T]) -> the bytes of the wav FILE `_save_audio` would have written
    (decoders.py:40-47): only 2 bytes per sample cross the PCIe bus, no temp file (serving.py:96-97 re-reads the file it
    just wrote; a server can return these bytes directly)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    w = wav.detach().reshape(-1).to(torch.float32).contiguous()
    assert w.is_cuda, "wav_bytes_on_device expects a device tensor (use audio_write_wav for host tensors)"
    n = w.numel()
    key = (w.device, max(n, 1 << 17))
    ws = _POST_WS.get(w.device)
    if ws is None or ws[0] < n:
        cap = max(n, 1 << 20)
        ws = (cap, torch.empty(lib.mvb_audio_post_workspace_bytes(cap), dtype=torch.uint8, device=w.device))
        _POST_WS[w.device] = ws
    pcm = torch.empty(n, dtype=torch.int16, device=w.device)
    stats = torch.empty(2, dtype=torch.float32, device=w.device)
    st = C.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)
    _lib.check(lib.mvb_audio_post(w.data_ptr(), n, int(sample_rate), float(loudness_headroom_db), int(bool(loudness_compressor)),
                                  ws[1].data_ptr(), pcm.data_ptr(), None, stats.data_ptr(), st))
    body = pcm.cpu().numpy().astype("<i2").tobytes()
    header = (b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVEfmt " +
              struct.pack("<IHHIIHH", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16) + b"data" + struct.pack("<I", len(body)))
    if return_stats:
        lk, gain = stats.cpu().tolist()
        return header + body, lk, gain
    return header + body
