"""Host-side mirror of ``fam/llm/fast_model.py`` over libmvb200.

Same names and call signatures as the reference (``ModelArgs``, ``Transformer.from_name``,
``setup_spk_cond_mask``, ``setup_caches``, ``forward(idx, spk_emb, input_pos)``) so that
``fam/llm/fast_inference.py`` works by changing one import; the arithmetic runs in
``libmvb200.so``.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib

_ALIGN = 256  # byte alignment of every tensor inside the weight arena (TMA needs >= 128)


def find_multiple(n: int, k: int) -> int:
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class ModelArgs:
    """fam/llm/fast_model.py:52-84."""
    block_size: int = 2048
    vocab_size: int = 32000
    n_layer: int = 32
    n_head: int = 32
    dim: int = 4096
    speaker_emb_dim: int = 256
    intermediate_size: Optional[int] = None
    n_local_heads: int = -1
    head_dim: int = 64
    norm_eps: float = 1e-5
    dtype: torch.dtype = torch.bfloat16

    def __post_init__(self):
        if self.n_local_heads == -1:
            self.n_local_heads = self.n_head
        if self.intermediate_size is None:
            self.intermediate_size = find_multiple(int(2 * 4 * self.dim / 3), 256)
        self.head_dim = self.dim // self.n_head

    @classmethod
    def from_name(cls, name: str):
        if name in transformer_configs:
            return cls(**transformer_configs[name])
        config = [c for c in transformer_configs if c in str(name).upper() or c in str(name)]
        assert len(config) == 1, name
        return cls(**transformer_configs[config[0]])


transformer_configs = {"metavoice-1B": dict(n_layer=24, n_head=16, dim=2048, vocab_size=2562)}

# arena order == MVB_S1_GLOBAL_TENSORS / MVB_S1_LAYER_TENSORS in include/mvb200.h; names are the
# checkpoint's (fast_inference_utils.py:250-278 maps them onto the fast model's names)
_GLOBAL_KEYS = ["transformer.wtes.0.weight", "transformer.wpe.weight", "speaker_cond_pos.weight",
                "transformer.ln_f.weight", "lm_heads.0.weight"]
_LAYER_KEYS = ["ln_1.weight", "attn.c_attn.weight", "attn.c_proj.weight", "ln_2.weight",
               "mlp.swiglu.w1.weight", "mlp.swiglu.w3.weight", "mlp.c_proj.weight"]


def pack_arena(state_dict: Dict[str, torch.Tensor], n_layer: int):
    """Checkpoint state dict -> one contiguous bf16 byte arena (CPU) + byte offsets.
    Replaces the module construction of ``_load_model`` (fast_inference_utils.py:236-281)."""
    sd = {(k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k): v for k, v in state_dict.items()}
    keys = list(_GLOBAL_KEYS)
    for i in range(n_layer):
        keys += [f"transformer.h.{i}.{k}" for k in _LAYER_KEYS]
    offsets: List[int] = []
    total = 0
    for k in keys:
        if k not in sd:
            raise KeyError(f"checkpoint is missing {k}")
        offsets.append(total)
        total = (total + sd[k].numel() * 2 + _ALIGN - 1) // _ALIGN * _ALIGN
    arena = torch.zeros(total, dtype=torch.uint8)
    for k, o in zip(keys, offsets):
        t = sd[k].detach().to(torch.bfloat16).contiguous().reshape(-1)
        arena[o:o + t.numel() * 2] = t.view(torch.uint8)
    return arena, offsets


class KVCache:
    """Placeholder with the reference's attribute names (fast_model.py:97-113); the cache itself is a
    single device buffer owned by the Transformer (layout in DESIGN.md)."""

    def __init__(self, k_cache: torch.Tensor, v_cache: torch.Tensor):
        self.k_cache, self.v_cache = k_cache, v_cache


class Transformer:
    """Mirror of fam/llm/fast_model.py:116-167 backed by an ``mvb_s1`` engine handle."""

    def __init__(self, config: ModelArgs, arena: torch.Tensor, offsets: List[int], device="cuda"):
        self.config = config
        self.device = torch.device(device)
        self._lib = _lib.load()
        self._arena = arena.to(self.device) if arena.device != self.device else arena
        self._offsets = list(offsets)
        self._handle = None
        self._kv = self._ws = None
        self.max_batch_size = -1
        self.max_seq_length = -1
        self.max_utts = 0
        self._spk_key = None
        self.kv_dtype = "bf16"

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_name(cls, name: str):
        raise RuntimeError("mvb200.Transformer holds device weights; use Transformer.from_state_dict(...)")

    @classmethod
    def from_state_dict(cls, state_dict, config: ModelArgs, device="cuda"):
        arena, offsets = pack_arena(state_dict, config.n_layer)
        return cls(config, arena, offsets, device)

    def model_size_bytes(self) -> int:
        """Σ numel*itemsize of parameters == the reference's ``model_size`` (utils:348) before caches."""
        c = self.config
        per_layer = 4 * c.dim * c.dim + 3 * c.dim * c.intermediate_size + 2 * c.dim
        n = c.n_layer * per_layer + 2 * c.vocab_size * c.dim + c.block_size * c.dim + c.speaker_emb_dim * c.dim + c.dim
        return n * 2

    def setup_spk_cond_mask(self):
        # fast_model.py:132-134: row 0 conditioned, row 1 unconditioned -- built into the kernels.
        self.spk_cond_mask = torch.zeros((2, 1, self.config.dim), dtype=torch.bool)
        self.spk_cond_mask[0] = 1

    def setup_caches(self, max_batch_size: int, max_seq_length: int, kv_dtype: str = "bf16", max_new: Optional[int] = None,
                     tensor_core_path: Optional[bool] = None):
        """fast_model.py:136-148.  ``max_batch_size`` counts rows: 2 per utterance (CFG pair).
        ``tensor_core_path`` (test hook): force the tcgen05 rows path on/off for prefill and batched decode."""
        if (self._handle is not None and self.max_seq_length >= max_seq_length and self.max_batch_size >= max_batch_size
                and kv_dtype == self.kv_dtype and tensor_core_path is None):
            return
        c = self.config
        if max_batch_size % 2:
            raise ValueError("max_batch_size must be even: every utterance owns a {cond, uncond} row pair")
        if find_multiple(max_seq_length, 8) != c.block_size:
            raise ValueError("the engine sizes the KV cache to block_size slots per row")
        self.close()
        self.max_seq_length, self.max_batch_size = c.block_size, max_batch_size
        self.max_utts = max_batch_size // 2
        self.kv_dtype = kv_dtype
        cfg = _lib.S1Config(c.n_layer, c.n_head, c.head_dim, c.dim, c.intermediate_size, c.vocab_size, c.block_size,
                            c.speaker_emb_dim, c.norm_eps, self.max_utts,
                            _lib.MVB_KV_FP32 if kv_dtype == "fp32" else _lib.MVB_KV_BF16, max_new or c.block_size)
        self._cfg = cfg
        kvb = self._lib.mvb_s1_kv_bytes(C.byref(cfg))
        wsb = self._lib.mvb_s1_workspace_bytes(C.byref(cfg))
        if kvb == 0 or wsb == 0:
            _lib.check(_lib.MVB_ERR_UNSUPPORTED)
        self._kv = torch.zeros(kvb, dtype=torch.uint8, device=self.device)
        self._ws = torch.zeros(wsb, dtype=torch.uint8, device=self.device)
        offs = (C.c_uint64 * len(self._offsets))(*self._offsets)
        h = C.c_void_p()
        torch.cuda.synchronize(self.device)
        import os
        # test hook: True/"BC" = all engine paths, False/"" = CUDA-core kernels only, "B" = tensor-core rows path
        # without the persistent decode kernel, "C" = persistent decode kernel with CUDA-core prefill
        saved = {k: os.environ.get(k) for k in ("MVB_PATHB", "MVB_PATHC")}
        if tensor_core_path is not None:
            paths = ("BC" if tensor_core_path else "") if isinstance(tensor_core_path, bool) else str(tensor_core_path)
            os.environ["MVB_PATHB"] = "1" if "B" in paths else "0"
            os.environ["MVB_PATHC"] = "1" if "C" in paths else "0"
        try:
            with torch.cuda.device(self.device):
                rc = self._lib.mvb_s1_create(C.byref(cfg), self._arena.data_ptr(), self._arena.numel(), offs,
                                             self._kv.data_ptr(), self._ws.data_ptr(), C.byref(h))
        finally:
            if tensor_core_path is not None:
                for k, v in saved.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
        _lib.check(rc)
        self._handle = h
        self._spk_key = None

    def close(self):
        if self._handle is not None:
            self._lib.mvb_s1_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers --------------------------------------------------------------------------------
    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError("call setup_caches() first")
        return self._handle

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_speaker(self, utt: int, spk_emb: torch.Tensor):
        e = spk_emb.detach().reshape(-1).to(device=self.device, dtype=torch.float32).contiguous()
        assert e.numel() == self.config.speaker_emb_dim
        _lib.check(self._lib.mvb_s1_set_speaker(self.handle, utt, e.data_ptr(), self._stream()))
        self._keepalive = e

    # ---- Transformer.forward (fast_model.py:150-163) -------------------------------------------
    def forward(self, idx: torch.Tensor, spk_emb: torch.Tensor, input_pos: torch.Tensor, utt: int = 0) -> torch.Tensor:
        """idx int [2,S]; spk_emb [1,speaker_emb_dim]; input_pos int [S] (must be consecutive).
        Returns fp32 logits [2,S,vocab] and updates the KV cache at ``input_pos``."""
        B, S = idx.shape
        if B != 2:
            raise ValueError("idx must hold the {cond, uncond} row pair of one utterance")
        ip = input_pos.detach().to("cpu", torch.int64)
        pos0 = int(ip[0])
        if S > 1 and not torch.equal(ip, torch.arange(pos0, pos0 + S)):
            raise ValueError("input_pos must be consecutive positions")
        self.set_speaker(utt, spk_emb)
        idx_d = idx.detach().to(device=self.device, dtype=torch.int32).contiguous()
        logits = torch.empty((2, S, self.config.vocab_size), dtype=torch.float32, device=self.device)
        _lib.check(self._lib.mvb_s1_forward(self.handle, utt, idx_d.data_ptr(), S, pos0, logits.data_ptr(), 1,
                                            self._stream()))
        return logits

    __call__ = forward
