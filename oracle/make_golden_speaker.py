"""ORACLE (test infrastructure): tests/golden/speaker.npz from the REFERENCE'S OWN ``SpeakerEncoder``
(fam/quantiser/audio/speaker_encoder/model.py) on a seeded synthetic ``speaker_encoder.pt``: (a) ``forward`` on random mel
partials, (b) ``embed_utterance`` on a synthetic 6.4 s waveform -- its slicing, padding, batching and averaging are the
reference's statements; only ``audio.wav_to_mel_spectrogram`` (librosa, absent) is replaced by the restatement in
oracle/speaker_port.py, whose output is stored too so the CUDA front-end can be checked against the same numbers.
Build container only:  python oracle/make_golden_speaker.py"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "metavoice-src_b200"))
from mvb200 import synth  # noqa: E402
from oracle import ref_harness as R, speaker_port as P  # noqa: E402


def main():
    R._import_reference()
    from fam.quantiser.audio.speaker_encoder import audio as ref_audio
    from fam.quantiser.audio.speaker_encoder.model import SpeakerEncoder
    state = synth.speaker_encoder_state_dict(3)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "speaker_encoder.pt")
        torch.save({"model_state": state}, path)
        enc = SpeakerEncoder(weights_fpath=path, device="cpu", verbose=False, eval=True)
    g = torch.Generator().manual_seed(11)
    mels = torch.rand(5, 160, 40, generator=g) * 3.0
    with torch.no_grad():
        emb = enc(mels)
    wav = synth.synthetic_waveform(6.4, 16000, seed=5)
    ref_audio.wav_to_mel_spectrogram = P.mel_spectrogram            # librosa is absent: restated front-end (unpinned)
    e_utt, partials, wav_slices = enc.embed_utterance(wav, return_partials=True)
    ws, ms = enc.compute_partial_slices(len(wav), 1.3, 0.75)
    ws2, ms2 = P.compute_partial_slices(len(wav))
    assert [(s.start, s.stop) for s in ms] == ms2 and [(s.start, s.stop) for s in ws] == ws2
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "speaker.npz"),
                        mels=mels.numpy(), emb=emb.numpy(), wav=wav, utt_embed=e_utt, partials=partials.numpy(),
                        mel_slices=np.asarray(ms2, np.int32), mel=P.mel_spectrogram(np.pad(wav, (0, max(0, ws2[-1][1] - len(wav))))),
                        checksum=np.float64(synth.state_dict_checksum(state)))
    print("speaker golden:", emb.shape, partials.shape, e_utt.shape)


if __name__ == "__main__":
    main()
