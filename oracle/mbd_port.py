"""ORACLE (test infrastructure, never the product path): CPU restatement of the multi-band-diffusion vocoder that
``EncodecDecoder.decode`` calls (fam/llm/decoders.py:13 ``MultiBandDiffusion.get_mbd_24khz(bw=6)``, :84-85
``mbd.tokens_to_wav(tokens)``).

PARITY UNPINNED.  The arithmetic lives in third-party ``audiocraft==1.2.0`` (requirements.txt:21) + ``julius==0.2.7``
and in the checkpoint ``facebook/multiband-diffusion::mbd_comp_8.pt`` (per-band model / schedule / processor configs and
weights); neither the packages nor the checkpoint exist in this image, and the reference has no test or golden vector
for this boundary.  This file restates the published algorithm -- arXiv 2308.02560 and the
audiocraft 1.2.0 modules ``models/multibanddiffusion.py``, ``models/unet.py``, ``modules/diffusion_schedule.py``,
``solvers/diffusion.py`` (processors) and julius ``bands.py`` / ``lowpass.py`` -- against a PARAMETRISED configuration
(``MBDConfig``); the real widths, depths and schedule constants are whatever the checkpoint's ``cfg`` holds.

    tokens_to_wav(tokens):  wav_encodec = codec.decode(tokens);  cond = codec.decode_latent(tokens)
                            wav = sum over the 4 band models of  DiffusionProcess.generate(cond, randn)      (20 steps each)
                            return re_eq(wav, ref=wav_encodec, n_bands=32)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class UnetCfg:            # audiocraft DiffusionUnet kwargs (config/model/score/basic.yaml as recalled; codec_dim set by the solver)
    chin: int = 1
    hidden: int = 48
    depth: int = 4
    growth: float = 4.0
    kernel: int = 8
    stride: int = 4
    res_blocks: int = 1
    norm_groups: int = 4
    emb_all_layers: bool = True
    codec_dim: int = 128
    num_steps: int = 1000
    max_channels: int = 10_000

    def channels(self) -> List[int]:
        ch, h = [], self.hidden
        for _ in range(self.depth):
            ch.append(h)
            h = min(int(h * self.growth), self.max_channels)
        return ch


@dataclass
class ScheduleCfg:        # audiocraft NoiseSchedule kwargs
    beta_t0: float = 1e-5
    beta_t1: float = 2.9e-2
    beta_exp: float = 7.5
    num_steps: int = 1000
    clip: float = 5.0
    rescale: float = 1.0
    noise_scale: float = 1.0


@dataclass
class MBDConfig:
    sample_rate: int = 24000
    n_models: int = 4                 # band models summed by MultiBandDiffusion.generate
    unet: UnetCfg = field(default_factory=UnetCfg)
    schedule: ScheduleCfg = field(default_factory=ScheduleCfg)
    proc_bands: int = 8               # MultiBandProcessor.n_bands
    power_std: float = 1.0
    eq_bands: int = 32                # tokens_to_wav(n_bands=32)
    step_list: Optional[List[int]] = None

    def steps(self) -> List[int]:
        return self.step_list if self.step_list is not None else list(range(self.schedule.num_steps))[::-50] + [0]


# ---- julius.SplitBands (bands.py) over julius.LowPassFilters (lowpass.py) -------------------------------------------
def mel_cutoffs(n_bands: int, sample_rate: int) -> List[float]:
    hz2mel = lambda f: 2595 * math.log10(1 + f / 700)
    mel2hz = lambda m: 700 * (10 ** (m / 2595) - 1)
    lo, hi = hz2mel(0.0), hz2mel(sample_rate / 2)
    mels = [lo + (hi - lo) * i / n_bands for i in range(n_bands + 1)]
    return [mel2hz(m) for m in mels][1:-1]


def lowpass_bank(n_bands: int, sample_rate: int, zeros: float = 8.0) -> torch.Tensor:
    """[n_bands - 1, 2 * half + 1] windowed-sinc low-pass filters at the mel-spaced cutoffs."""
    cut = [c / sample_rate for c in mel_cutoffs(n_bands, sample_rate)]
    half = int(zeros / min(cut) / 2)
    window = torch.hann_window(2 * half + 1, periodic=False, dtype=torch.float64)
    t = torch.arange(-half, half + 1, dtype=torch.float64)
    fs = []
    for c in cut:
        x = 2 * c * math.pi * t
        sinc = torch.where(x == 0, torch.ones_like(x), torch.sin(x) / x)
        f = 2 * c * window * sinc
        fs.append(f / f.sum())
    return torch.stack(fs).float()


def split_bands(x: torch.Tensor, n_bands: int, sample_rate: int) -> torch.Tensor:
    """x [T] -> [n_bands, T]: differences of successive low-passed copies (replicate padding), last band = x - low."""
    bank = lowpass_bank(n_bands, sample_rate)
    half = (bank.shape[1] - 1) // 2
    xp = F.pad(x.view(1, 1, -1), (half, half), mode="replicate")
    lows = F.conv1d(xp, bank[:, None, :])[0]          # [n_bands - 1, T]
    bands, low = [lows[0]], lows[0]
    for i in range(1, lows.shape[0]):
        bands.append(lows[i] - low)
        low = lows[i]
    bands.append(x - low)
    return torch.stack(bands)


# ---- audiocraft.models.unet.DiffusionUnet ------------------------------------------------------------------------
def _res_block(sd, p, x, groups, dilation):
    h = F.conv1d(F.relu(F.group_norm(x, groups, sd[p + "norm1.weight"], sd[p + "norm1.bias"])), sd[p + "conv1.weight"],
                 sd[p + "conv1.bias"], padding=dilation, dilation=dilation)
    h = F.conv1d(F.relu(F.group_norm(h, groups, sd[p + "norm2.weight"], sd[p + "norm2.bias"])), sd[p + "conv2.weight"],
                 sd[p + "conv2.bias"], padding=dilation, dilation=dilation)
    return x + h


def unet_forward(sd: Dict[str, torch.Tensor], c: UnetCfg, x: torch.Tensor, step: int, condition: torch.Tensor) -> torch.Tensor:
    """x [1, chin, T], condition [1, codec_dim, T_frames] -> estimate [1, chin, T]."""
    pad_k = (c.kernel - c.stride) // 2
    skips, z = [], x
    for i in range(c.depth):
        p = f"encoders.{i}."
        T = z.shape[-1]
        z = F.pad(z, (0, (c.stride - T % c.stride) % c.stride))
        z = F.conv1d(z, sd[p + "conv.weight"], None, stride=c.stride, padding=pad_k)
        z = F.relu(F.group_norm(z, c.norm_groups, sd[p + "norm.weight"], sd[p + "norm.bias"]))
        for j in range(c.res_blocks):
            z = _res_block(sd, f"{p}res_blocks.{j}.", z, c.norm_groups, 2 ** j)
        emb = sd["embedding.weight"] if i == 0 else (sd[f"embeddings.{i - 1}.weight"] if c.emb_all_layers else None)
        if emb is not None:
            z = z + emb[step].view(1, -1, 1)
        skips.append(z)
    cond = F.conv1d(condition, sd["conv_codec.weight"], sd["conv_codec.bias"])
    assert cond.shape[-1] <= 2 * z.shape[-1]
    z = z + F.interpolate(cond, z.shape[-1])             # default mode: nearest
    # bottleneck: no BLSTM / transformer in this parametrisation (identity pass-through)
    for i in range(c.depth):
        p = f"decoders.{i}."
        s = skips.pop(-1)
        z = z[:, :, : s.shape[2]] + s
        for j in range(c.res_blocks):
            z = _res_block(sd, f"{p}res_blocks.{j}.", z, c.norm_groups, 2 ** j)
        z = F.relu(F.group_norm(z, c.norm_groups, sd[p + "norm.weight"], sd[p + "norm.bias"]))
        z = F.conv_transpose1d(z, sd[p + "convtr.weight"], None, stride=c.stride, padding=pad_k)
    return z[:, :, : x.shape[2]]


# ---- audiocraft.modules.diffusion_schedule.NoiseSchedule.generate_subsampled -------------------------------------------
def schedule_coefficients(s: ScheduleCfg, step_list: List[int]):
    """Per model call: (a, b, sigma) with previous = (current - a * estimate) * b (+ sigma * noise), then clamp."""
    betas = torch.linspace(s.beta_t0 ** (1 / s.beta_exp), s.beta_t1 ** (1 / s.beta_exp), s.num_steps, dtype=torch.float32) ** s.beta_exp
    abar_all = (1 - betas).cumprod(dim=0)
    sub = abar_all[list(reversed(step_list))]
    betas_sub = 1 - torch.cat([sub[:1], sub[1:] / sub[:-1]])
    alpha_bar = abar_all[s.num_steps - 1]
    out = []
    for idx, step in enumerate(step_list[:-1]):
        alpha = 1 - betas_sub[-1 - idx]
        a = (1 - alpha) / (1 - alpha_bar).sqrt()
        b = 1 / alpha.sqrt()
        prev_abar = abar_all[step_list[idx + 1]]
        if step == step_list[-2]:
            sigma2, prev_abar = torch.tensor(0.0), torch.tensor(1.0)
        else:
            sigma2 = (1 - prev_abar) / (1 - alpha_bar) * (1 - alpha)
        out.append((float(a), float(b), float(sigma2.clamp(min=0).sqrt())))
        alpha_bar = prev_abar
    return out


class MBDOracle:
    def __init__(self, ckpt: dict, cfg: MBDConfig):
        """ckpt: {"models": [state_dict x n_models], "proc": [{"mean", "std", "target_std"} x n_models]} (fp32)."""
        self.cfg, self.models, self.proc = cfg, [{k: v.float() for k, v in m.items()} for m in ckpt["models"]], ckpt["proc"]

    @torch.no_grad()
    def generate_band(self, m: int, cond: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """cond [codec_dim, T_f]; noise [n_calls, T] (row 0 = initial sample, row i = the draw added after call i)."""
        c, s = self.cfg, self.cfg.schedule
        steps = c.steps()
        cur = (noise[0] * s.noise_scale).view(1, 1, -1)
        for i, (a, b, sigma) in enumerate(schedule_coefficients(s, steps)):
            est = unet_forward(self.models[m], c.unet, cur, steps[i], cond[None]) * s.noise_scale
            prev = (cur - a * est) * b
            if sigma > 0:
                prev = prev + sigma * noise[i + 1].view(1, 1, -1) * s.noise_scale
            if s.clip:
                prev = prev.clamp(-s.clip, s.clip)
            cur = prev
        x = cur.view(-1)        # (`rescale` only applies when step 0 is itself evaluated: never with the default step list)
        # MultiBandProcessor.return_sample: per processor band  x_b * (std / target_std) ** power_std + mean, summed
        P = self.proc[m]
        bands = split_bands(x, c.proc_bands, c.sample_rate)
        scale = (P["std"] / P["target_std"]) ** c.power_std
        return (bands * scale.view(-1, 1) + P["mean"].view(-1, 1)).sum(0)

    @torch.no_grad()
    def tokens_to_wav(self, wav_encodec: torch.Tensor, cond: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """wav_encodec [T] (codec decode, the re-EQ reference), cond [codec_dim, T_f], noise [n_models, n_calls, T] -> [T]."""
        c = self.cfg
        wav = torch.zeros_like(wav_encodec)
        for m in range(c.n_models):
            wav = wav + self.generate_band(m, cond, noise[m])
        return re_eq(wav, wav_encodec, c.eq_bands, c.sample_rate)


def re_eq(wav: torch.Tensor, ref: torch.Tensor, n_bands: int, sample_rate: int, strictness: float = 1.0) -> torch.Tensor:
    """MultiBandDiffusion.re_eq: match the energy of every mel band of `wav` to `ref`."""
    b, br = split_bands(wav, n_bands, sample_rate), split_bands(ref, n_bands, sample_rate)
    out = torch.zeros_like(ref)
    for i in range(n_bands):
        out = out + b[i] * (br[i].std() / b[i].std()) ** strictness
    return out
