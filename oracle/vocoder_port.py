"""ORACLE (test infrastructure, never the product path): CPU restatement of the EnCodec-24 kHz decode path that
``mbd.tokens_to_wav`` runs first (fam/llm/decoders.py:85 -> audiocraft 1.2.0 ``MultiBandDiffusion.tokens_to_wav`` ->
``codec_model.decode`` / ``decode_latent``).  audiocraft is NOT under /root/reference and not installed; for the
24 kHz codec it delegates to ``transformers.EncodecModel`` (HFEncodecCompressionModel), whose source IS on disk
(site-packages/transformers/models/encodec/modeling_encodec.py) and is what this file restates and is pinned against:

  * RVQ decode: sum of codebook rows                        modeling_encodec.py EncodecResidualVectorQuantizer.decode
  * causal Conv1d with reflect left padding + weight norm   EncodecConv1d            (:82-172)
  * ConvTranspose1d, right-trimmed by (k - stride)          EncodecConvTranspose1d   (:175-233)
  * 2-layer LSTM with skip                                  EncodecLSTM              (:236-249)
  * residual block (ELU, k3, ELU, k1) + 1x1 shortcut        EncodecResnetBlock       (:252-282)
  * decoder stack, ratios (8, 5, 4, 2)                      EncodecDecoder           (:316-347)

The multi-band diffusion stage itself (4 band UNets x 20 steps, re-EQ) exists only inside audiocraft + mbd_comp_8.pt:
PARITY UNPINNED, not restated here.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F


def fold_weight_norm(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """w = g * v / ||v|| with the norm over every dim but 0 (torch weight_norm, dim=0); accepts both the
    parametrizations.* naming of current torch and the legacy weight_g / weight_v naming."""
    if prefix + ".parametrizations.weight.original0" in sd:
        g, v = sd[prefix + ".parametrizations.weight.original0"], sd[prefix + ".parametrizations.weight.original1"]
    elif prefix + ".weight_g" in sd:
        g, v = sd[prefix + ".weight_g"], sd[prefix + ".weight_v"]
    else:
        return sd[prefix + ".weight"].float()
    v = v.float()
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g.float() / n)


def causal_conv1d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """stride 1, dilation 1: left reflect padding of k-1 samples (EncodecConv1d.forward, causal branch)."""
    k = w.shape[-1]
    if k > 1:
        length = x.shape[-1]
        extra = 0
        if length <= k - 1:          # _pad1d: reflect needs length > pad
            extra = k - 1 - length + 1
            x = F.pad(x, (0, extra))
        x = F.pad(x, (k - 1, 0), mode="reflect")
        if extra:
            x = x[..., : x.shape[-1] - extra]
    return F.conv1d(x, w, b)


def conv_transpose1d_trim(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int) -> torch.Tensor:
    y = F.conv_transpose1d(x, w, b, stride=stride)
    return y[..., : y.shape[-1] - (w.shape[-1] - stride)]      # causal: all (k - stride) trimmed on the right


class EncodecDecodeOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], ratios=(8, 5, 4, 2), n_q: int = 8):
        self.ratios = ratios
        self.codebooks = [sd[f"quantizer.layers.{q}.codebook.embed"].float() for q in range(n_q)]
        cw = lambda p: (fold_weight_norm(sd, p + ".conv"), sd[p + ".conv.bias"].float())
        self.conv_in = cw("decoder.layers.0")
        self.lstm = [{k: sd[f"decoder.layers.1.lstm.{k}_l{l}"].float() for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")}
                     for l in range(2)]
        self.ups, self.res = [], []
        i = 3
        for _ in ratios:
            self.ups.append(cw(f"decoder.layers.{i}"))
            p = f"decoder.layers.{i + 1}"
            self.res.append(dict(c1=cw(p + ".block.1"), c2=cw(p + ".block.3"), sc=cw(p + ".shortcut")))
            i += 3
        self.conv_out = cw(f"decoder.layers.{i}")

    def decode_latent(self, codes: torch.Tensor) -> torch.Tensor:
        """codes int [B, n_q, T] -> [B, 128, T]."""
        return sum(cb[codes[:, q]] for q, cb in enumerate(self.codebooks)).transpose(1, 2)

    def _lstm(self, x: torch.Tensor) -> torch.Tensor:       # x [B, C, T]
        seq = x.permute(2, 0, 1)
        out = seq
        for L in self.lstm:
            H = L["weight_hh"].shape[1]
            h = torch.zeros(seq.shape[1], H); c = torch.zeros(seq.shape[1], H)
            ys = []
            for t in range(out.shape[0]):
                g = out[t] @ L["weight_ih"].t() + L["bias_ih"] + h @ L["weight_hh"].t() + L["bias_hh"]
                i, f, gg, o = g.chunk(4, dim=-1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                ys.append(h)
            out = torch.stack(ys)
        return (out + seq).permute(1, 2, 0)

    @torch.no_grad()
    def decode(self, codes: torch.Tensor, taps: dict = None) -> torch.Tensor:
        """codes int [B, n_q, T] -> waveform [B, 1, 320*T]."""
        x = self.decode_latent(codes)
        x = causal_conv1d(x, *self.conv_in)
        x = self._lstm(x)
        if taps is not None:
            taps["lstm"] = x.clone()
        for (uw, ub), r, blk in zip(self.ups, self.ratios, self.res):
            x = conv_transpose1d_trim(F.elu(x), uw, ub, r)
            h = causal_conv1d(F.elu(x), *blk["c1"])
            h = causal_conv1d(F.elu(h), *blk["c2"])
            x = causal_conv1d(x, *blk["sc"]) + h
        return causal_conv1d(F.elu(x), *self.conv_out)
