"""Host-side mirror of the façade ``fam/llm/fast_inference.py::TTS`` over libmvb200.

Same constructor keywords and the same ``synthesise(text, spk_ref_path, top_p, guidance_scale, temperature) -> path``
contract (fast_inference.py:41-50, 111-119, 195).  Differences, all stated where they occur:
  * checkpoints are read from a local directory (the reference calls ``snapshot_download``, :71; no network here);
  * the speaker reference may be a RIFF/WAVE file (>= 30 s, utils.py:55-70), embedded by the on-device speaker encoder
    (mvb200/speaker_encoder.py) and disk-cached like the reference (inference.py:419-435), or an already
    computed embedding (``.pt`` tensor / ``.npy``); mp3 / flac / URLs need a decoder / network that this image lacks;
  * the vocoder stage is the EnCodec decoder followed by the multi-band-diffusion refinement (mvb200/mbd.py; parity
    UNPINNED, its configuration is read from ``multiband_diffusion.pt`` in the model directory or passed in) when such a
    checkpoint is available; the DeepFilterNet enhancer (fast_inference.py:158-163) is not implemented (N1);
  * the wav goes through the reference's ``audio_write`` post-processing (loudness normalisation to -14 LUFS, tanh
    compressor, PCM16; decoders.py:40-47) in mvb200/audio_out.py.
"""
from __future__ import annotations

import os
import re
import time
import uuid
from datetime import datetime
from pathlib import Path
from typing import Literal, Optional

import numpy as np
import torch

from .fast_inference_utils import build_model, main
from .second_stage import SecondStage, flattened_interleaved_decode
from .tokenise import TrainedBPETokeniser
from .vocoder import EncodecDecodeEngine

_UNICODE = {8175: "'", 8189: "'", 8190: "'", 8208: "-", 8209: "-", 8210: "-", 8211: "-", 8212: "-", 8213: "-", 8214: "||",
            8216: "'", 8217: "'", 8218: ",", 8219: "`", 8220: '"', 8221: '"', 8222: ",,", 8223: '"', 8228: ".", 8229: "..",
            8230: "...", 8242: "'", 8243: '"', 8245: "'", 8246: '"', 180: "'", 2122: "TM"}


from .audio_out import audio_write_wav  # noqa: E402


def normalize_text(text: str) -> str:
    """fam/llm/utils.py:12-52: map typographic punctuation, reject code points >= 256, collapse whitespace."""
    text = text.translate(_UNICODE)
    bad = {c for c in text if ord(c) >= 256}
    if bad:
        raise ValueError(f"Non-supported character found: {[(c, ord(c)) for c in bad]}")
    text = text.replace("\t", " ").replace("\n", " ").replace("\r", " ").replace("*", " ").strip()
    return re.sub(r"\s\s+", " ", text)


def load_speaker_embedding(path: str, smodel=None) -> torch.Tensor:
    if not os.path.exists(path):
        raise FileNotFoundError(f"File {path} not found!")      # inference.py:412-415, 420-421
    if path.endswith(".npy"):
        e = torch.from_numpy(np.load(path))
    elif path.endswith(".pt"):
        e = torch.load(path, map_location="cpu")
    else:
        if smodel is None:
            raise FileNotFoundError("speaker_encoder.pt not found in the model directory: pass a precomputed embedding (.pt / .npy)")
        from .speaker_encoder import check_audio_file, get_cached_embedding
        check_audio_file(path)                                  # fast_inference.py:123: >= 30 s of reference audio
        e = get_cached_embedding(path, smodel)                  # fast_inference.py:124-127
    return e.reshape(1, -1).float()


class TTS:
    END_OF_AUDIO_TOKEN = 1024

    def __init__(self, model_name: str = "metavoiceio/metavoice-1B-v0.1", *, seed: int = 1337, output_dir: str = "outputs",
                 quantisation_mode: Optional[Literal["int4", "int8"]] = None, first_stage_path: Optional[str] = None,
                 telemetry_origin: Optional[str] = None, encodec_state_dict=None, device: str = "cuda", max_utts: int = 1,
                 mbd_checkpoint: Optional[dict] = None, mbd_settings=None):
        self._device = device
        self._model_dir = model_name
        if not os.path.isdir(model_name):
            raise FileNotFoundError(f"{model_name}: pass a local snapshot directory holding first_stage.pt / second_stage.pt "
                                    "(there is no network access for snapshot_download)")
        self.output_dir = output_dir
        os.makedirs(self.output_dir, exist_ok=True)
        self._first_stage_ckpt = first_stage_path or f"{self._model_dir}/first_stage.pt"
        torch.manual_seed(seed)                                                        # inference.py:69-70
        ck2 = torch.load(f"{self._model_dir}/second_stage.pt", map_location="cpu", weights_only=False)
        tok2 = TrainedBPETokeniser(**ck2["meta"]["tokenizer"])
        self.llm_second_stage = SecondStage(ck2, device=device, tokenizer=tok2)
        if encodec_state_dict is None:
            encodec_state_dict = torch.load(f"{self._model_dir}/encodec_24khz.pt", map_location="cpu", weights_only=False)
        self.codec = EncodecDecodeEngine(encodec_state_dict, device=device)
        # decoders.py:13: MultiBandDiffusion.get_mbd_24khz(bw=6) -- here: {"models", "proc", "settings"} (see mvb200/mbd.py)
        self.mbd = None
        mbd_path = f"{self._model_dir}/multiband_diffusion.pt"
        if mbd_checkpoint is None and os.path.isfile(mbd_path):
            mbd_checkpoint = torch.load(mbd_path, map_location="cpu", weights_only=False)
            mbd_settings = mbd_settings or mbd_checkpoint.get("settings")
        if mbd_checkpoint is not None:
            from .mbd import MBDSettings, MultiBandDiffusionEngine
            self.mbd = MultiBandDiffusionEngine(mbd_checkpoint, mbd_settings or MBDSettings(), device=device, max_seconds=30.0)
        self.precision = torch.bfloat16
        self.model, self.tokenizer, self.smodel, self.model_size = build_model(
            precision=self.precision, checkpoint_path=Path(self._first_stage_ckpt),
            spk_emb_ckpt_path=Path(f"{self._model_dir}/speaker_encoder.pt"), device=device, compile=True,
            compile_prefill=True, quantisation_mode=quantisation_mode, max_utts=max_utts)
        self._seed = seed
        spk_ckpt = f"{self._model_dir}/speaker_encoder.pt"
        if os.path.isfile(spk_ckpt):                                                   # fast_inference_utils.py:314-318
            from .speaker_encoder import SpeakerEncoder
            self.smodel = SpeakerEncoder(weights_fpath=spk_ckpt, device=device, eval=True, verbose=False)

    def synthesise(self, text: str, spk_ref_path: str, top_p=0.95, guidance_scale=3.0, temperature=1.0) -> str:
        text = normalize_text(text)
        spk_emb = load_speaker_embedding(spk_ref_path, self.smodel)
        start = time.time()
        tokens = main(model=self.model, tokenizer=self.tokenizer, model_size=self.model_size, prompt=text, spk_emb=spk_emb,
                      top_p=top_p, guidance_scale=guidance_scale, temperature=temperature, device=self._device)
        _, extracted = flattened_interleaved_decode(tokens, self.END_OF_AUDIO_TOKEN)
        codes = self.llm_second_stage.non_causal_sample(
            texts=[text], encodec_tokens=[torch.tensor(extracted, dtype=torch.int32).unsqueeze(0)],
            speaker_embs=spk_emb.unsqueeze(0), batch_size=1, top_k=200, temperature=1.0)[0]
        wav = self._tokens_to_wav(codes)
        if wav.shape[-1] < 9600:
            raise Exception("wav predicted is shorter than 400ms!")                    # decoders.py:88-91
        name = f"synth_{datetime.now().strftime('%y-%m-%d--%H-%M-%S')}_{text.replace(' ', '_')[:25]}_{uuid.uuid4()}"
        # decoders.py:40-47: audio_write(strategy="loudness", loudness_compressor=True) -> <name>.wav
        path = audio_write_wav(str(Path(self.output_dir).resolve() / name), wav.reshape(1, -1), 24000,
                               strategy="loudness", loudness_compressor=True)
        print(f"\nSaved audio to {path}")
        dt = time.time() - start
        dur = wav.shape[-1] / 24000.0
        print(f"\nTotal time to synth (s): {dt}")
        print(f"Real-time factor: {dt / dur:.2f}")
        return path

    def _tokens_to_wav(self, codes: torch.Tensor) -> torch.Tensor:
        """decoders.py:84-85 ``mbd.tokens_to_wav(tokens)``: codec decode, then (when a diffusion checkpoint is loaded) the
        multi-band-diffusion refinement conditioned on the codec latent and re-equalised against the codec waveform."""
        wav = self.codec.decode(codes)
        if self.mbd is not None:
            wav = self.mbd.tokens_to_wav(self.codec.decode_latent(codes), wav)
        return wav

    def synthesise_long(self, text: str, spk_ref_path: str, top_p=0.95, guidance_scale=3.0, temperature=1.0,
                        max_chars: int = 220) -> str:
        """Long-form synthesis (the reference truncates at 220 characters, inference.py:535-541: "Long form synthesis
        coming soon"): the text is cut into <= ``max_chars`` chunks at sentence boundaries, the chunks are decoded as one
        continuous batch over the engine's KV slots (``TTS(max_utts=...)``), every chunk goes through stage 2 and the
        vocoder, and the waveforms are concatenated into one file."""
        from .fast_inference_utils import encode_tokens
        from .serving import ContinuousBatcher, chunk_text
        chunks = chunk_text(normalize_text(text), max_chars)
        if not chunks:
            raise ValueError("empty text")
        spk_emb = load_speaker_embedding(spk_ref_path, self.smodel)
        start = time.time()
        cb = ContinuousBatcher(self.model)
        ids = [cb.submit(encode_tokens(self.tokenizer, c, device="cpu"), spk_emb, top_p=top_p, guidance_scale=guidance_scale,
                         temperature=temperature) for c in chunks]
        toks = cb.run_until_done()
        wavs = []
        for c, rid in zip(chunks, ids):
            _, extracted = flattened_interleaved_decode(toks[rid].tolist(), self.END_OF_AUDIO_TOKEN)
            codes = self.llm_second_stage.non_causal_sample(
                texts=[c], encodec_tokens=[torch.tensor(extracted, dtype=torch.int32).unsqueeze(0)],
                speaker_embs=spk_emb.unsqueeze(0), batch_size=1, top_k=200, temperature=1.0)[0]
            wavs.append(self._tokens_to_wav(codes).reshape(-1))
        wav = torch.cat(wavs)
        if wav.shape[-1] < 9600:
            raise Exception("wav predicted is shorter than 400ms!")                    # decoders.py:88-91
        name = f"synth_{datetime.now().strftime('%y-%m-%d--%H-%M-%S')}_{chunks[0].replace(' ', '_')[:25]}_{uuid.uuid4()}"
        path = audio_write_wav(str(Path(self.output_dir).resolve() / name), wav.reshape(1, -1), 24000,
                               strategy="loudness", loudness_compressor=True)
        dt = time.time() - start
        print(f"\nSaved audio to {path} ({len(chunks)} chunks, {wav.shape[-1] / 24000.0:.1f} s of audio in {dt:.2f} s)")
        return path
