"""Host side of the encoder forward (K1/K2/K3): weights in HBM as bf16, packed
(unpadded) token batches, one C-ABI call per batch.

Mirrors what the reference does inside BGEEmbeddingModel._encode after
tokenisation (BGEEmbedding.py:119-127): BertModel forward -> mean_pooling ->
F.normalize.  Weight names follow HF's BertModel state dict.
"""
from __future__ import annotations

import ctypes as C
import itertools
import json
import os
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native


@dataclass
class EncoderConfig:
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    intermediate_size: int
    vocab_size: int
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    position_offset: int = 0  # 2 for XLM-R style checkpoints

    @classmethod
    def from_hf_json(cls, path: str) -> "EncoderConfig":
        with open(path) as f:
            d = json.load(f)
        mt = d.get("model_type", "bert")
        if mt not in ("bert", "xlm-roberta", "roberta"):
            raise ValueError(f"unsupported model_type {mt!r} (BERT-family encoders only)")
        if d.get("hidden_act", "gelu") != "gelu":
            raise ValueError("only exact-erf GELU encoders are supported")
        if d.get("position_embedding_type", "absolute") != "absolute":
            raise ValueError("only absolute position embeddings are supported")
        return cls(hidden_size=d["hidden_size"], num_hidden_layers=d["num_hidden_layers"],
                   num_attention_heads=d["num_attention_heads"], intermediate_size=d["intermediate_size"],
                   vocab_size=d["vocab_size"], max_position_embeddings=d.get("max_position_embeddings", 512),
                   type_vocab_size=d.get("type_vocab_size", 2), layer_norm_eps=d.get("layer_norm_eps", 1e-12),
                   position_offset=(d.get("pad_token_id", 1) + 1) if mt != "bert" else 0)

    # named shapes of BASELINE.json's configs
    @classmethod
    def bge_small(cls):
        return cls(384, 12, 12, 1536, 30522)

    @classmethod
    def bge_base(cls):
        return cls(768, 12, 12, 3072, 30522)

    @classmethod
    def bge_large(cls):
        return cls(1024, 24, 16, 4096, 30522)

    def flops_per_chunk(self, seq_len: int) -> float:
        """Algorithmic flops of one L-token chunk: Lyr*L*(8H^2 + 4HI + 4LH)."""
        H, I, L = self.hidden_size, self.intermediate_size, seq_len
        return float(self.num_hidden_layers) * L * (8.0 * H * H + 4.0 * H * I + 4.0 * L * H)


class _Layer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_qkv", "b_qkv", "w_o", "b_o", "ln1_g", "ln1_b", "w_ff1", "b_ff1",
                                          "w_ff2", "b_ff2", "ln2_g", "ln2_b")]


class _Model(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("n_layers", C.c_int32), ("heads", C.c_int32), ("intermediate", C.c_int32),
                ("vocab", C.c_int32), ("max_pos", C.c_int32), ("pos_offset", C.c_int32), ("ln_eps", C.c_float),
                ("word_emb", C.c_void_p), ("pos_emb", C.c_void_p), ("type_emb", C.c_void_p),
                ("emb_ln_g", C.c_void_p), ("emb_ln_b", C.c_void_p), ("layers", C.POINTER(_Layer))]


class _Head(C.Structure):
    _fields_ = [("w_dense", C.c_void_p), ("b_dense", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p),
                ("n_labels", C.c_int32)]


def _bind_encoder_abi(lib) -> None:
    if getattr(lib, "_crag_encoder_bound", False):
        return
    lib.crag_encoder_workspace_bytes.restype = C.c_size_t
    lib.crag_encoder_workspace_bytes.argtypes = [C.POINTER(_Model), C.c_int]
    lib.crag_encoder_forward.restype = C.c_int
    lib.crag_encoder_forward.argtypes = [C.POINTER(_Model), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.crag_encoder_classify.restype = C.c_int
    lib.crag_encoder_classify.argtypes = [C.POINTER(_Model), C.POINTER(_Head), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib._crag_encoder_bound = True


class BertEncoderB200:
    """BERT-family encoder resident on one GPU; thread-safe forward."""

    def __init__(self, config: EncoderConfig, state: Dict[str, torch.Tensor], device: Optional[torch.device] = None):
        if device is None:
            if not torch.cuda.is_available():
                raise _native.NativeError("BertEncoderB200 needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.config = config
        self.device = torch.device(device)
        self._lib = _native.load()
        _bind_encoder_abi(self._lib)
        self._tensors: List[torch.Tensor] = []  # keeps device storage alive
        self._build(state)
        self._lock = threading.Lock()

    # -------------------------------------------------------------- weights
    def _dev(self, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        d = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._tensors.append(d)
        return d

    def _build(self, sd: Dict[str, torch.Tensor]) -> None:
        cfg = self.config
        # sequence-classification checkpoints (bge-reranker-*): head weights sit beside the prefixed encoder
        self._head = None
        self.n_labels = 0
        if "classifier.dense.weight" in sd and "classifier.out_proj.weight" in sd:
            h = _Head()
            h.w_dense = self._dev(sd["classifier.dense.weight"], torch.bfloat16).data_ptr()
            h.b_dense = self._dev(sd["classifier.dense.bias"], torch.float32).data_ptr()
            h.w_out = self._dev(sd["classifier.out_proj.weight"], torch.bfloat16).data_ptr()
            h.b_out = self._dev(sd["classifier.out_proj.bias"], torch.float32).data_ptr()
            h.n_labels = self.n_labels = int(sd["classifier.out_proj.weight"].shape[0])
            self._head = h
        # accept both "bert."/"roberta."-prefixed and bare BertModel keys
        for pref in ("bert.", "roberta.", "model."):
            if any(k.startswith(pref + "embeddings.") for k in sd):
                sd = {k[len(pref):]: v for k, v in sd.items() if k.startswith(pref)}
                break
        bf, f32 = torch.bfloat16, torch.float32
        g = lambda k: sd[k]
        m = _Model()
        m.hidden, m.n_layers, m.heads = cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads
        m.intermediate, m.vocab, m.max_pos = cfg.intermediate_size, cfg.vocab_size, cfg.max_position_embeddings
        m.pos_offset, m.ln_eps = cfg.position_offset, cfg.layer_norm_eps
        m.word_emb = self._dev(g("embeddings.word_embeddings.weight"), bf).data_ptr()
        m.pos_emb = self._dev(g("embeddings.position_embeddings.weight"), bf).data_ptr()
        m.type_emb = self._dev(g("embeddings.token_type_embeddings.weight"), bf).data_ptr()
        m.emb_ln_g = self._dev(g("embeddings.LayerNorm.weight"), f32).data_ptr()
        m.emb_ln_b = self._dev(g("embeddings.LayerNorm.bias"), f32).data_ptr()
        layers = (_Layer * max(cfg.num_hidden_layers, 1))()
        for i in range(cfg.num_hidden_layers):
            p = f"encoder.layer.{i}."
            wq, wk, wv = (g(p + f"attention.self.{n}.weight") for n in ("query", "key", "value"))
            bq, bk, bv = (g(p + f"attention.self.{n}.bias") for n in ("query", "key", "value"))
            L = layers[i]
            L.w_qkv = self._dev(torch.cat([wq, wk, wv], 0), bf).data_ptr()
            L.b_qkv = self._dev(torch.cat([bq, bk, bv], 0), f32).data_ptr()
            L.w_o = self._dev(g(p + "attention.output.dense.weight"), bf).data_ptr()
            L.b_o = self._dev(g(p + "attention.output.dense.bias"), f32).data_ptr()
            L.ln1_g = self._dev(g(p + "attention.output.LayerNorm.weight"), f32).data_ptr()
            L.ln1_b = self._dev(g(p + "attention.output.LayerNorm.bias"), f32).data_ptr()
            L.w_ff1 = self._dev(g(p + "intermediate.dense.weight"), bf).data_ptr()
            L.b_ff1 = self._dev(g(p + "intermediate.dense.bias"), f32).data_ptr()
            L.w_ff2 = self._dev(g(p + "output.dense.weight"), bf).data_ptr()
            L.b_ff2 = self._dev(g(p + "output.dense.bias"), f32).data_ptr()
            L.ln2_g = self._dev(g(p + "output.LayerNorm.weight"), f32).data_ptr()
            L.ln2_b = self._dev(g(p + "output.LayerNorm.bias"), f32).data_ptr()
        self._layers = layers
        m.layers = C.cast(layers, C.POINTER(_Layer))
        self._model = m

    @classmethod
    def from_pretrained(cls, path: str, device: Optional[torch.device] = None) -> "BertEncoderB200":
        """Load an HF checkpoint directory (config.json + model.safetensors / pytorch_model.bin)."""
        cfg = EncoderConfig.from_hf_json(os.path.join(path, "config.json"))
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        return cls(cfg, sd, device)

    @classmethod
    def random_init(cls, cfg: EncoderConfig, seed: int = 0, std: float = 0.02,
                    device: Optional[torch.device] = None) -> "BertEncoderB200":
        """HF-style random initialisation (N(0, std) matrices, zero biases, unit LayerNorm), generated on the device."""
        return cls(cfg, random_state_dict(cfg, seed, std, device or torch.device("cuda", torch.cuda.current_device())), device)

    # -------------------------------------------------------------- forward
    def workspace_bytes(self, total_tokens: int) -> int:
        return int(self._lib.crag_encoder_workspace_bytes(C.byref(self._model), int(total_tokens)))

    def forward_packed(self, token_ids: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, normalize: bool = True,
                       out_f32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None,
                       stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """token_ids int32 [T] and cu_seqlens int32 [n+1] on the device -> fp32 [n, H] (device)."""
        n = cu_seqlens.numel() - 1
        T = token_ids.numel()
        H = self.config.hidden_size
        dev = self.device
        with torch.cuda.device(dev):
            st = stream if stream is not None else torch.cuda.current_stream(dev)
            with torch.cuda.stream(st):
                if out_f32 is None and out_bf16 is None:
                    out_f32 = torch.empty((n, H), dtype=torch.float32, device=dev)
                ws_bytes = self.workspace_bytes(T)
                ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
                rc = self._lib.crag_encoder_forward(
                    C.byref(self._model), token_ids.data_ptr(), cu_seqlens.data_ptr(), n, T, int(max_seqlen),
                    1 if normalize else 0, _native.ptr(out_f32), _native.ptr(out_bf16),
                    out_bf16.stride(0) if out_bf16 is not None else 0, ws.data_ptr(), ws_bytes, st.cuda_stream)
                _native.check(rc, "crag_encoder_forward")
        return out_f32 if out_f32 is not None else out_bf16

    def encode_token_lists(self, seqs: Sequence[Sequence[int]], normalize: bool = True) -> torch.Tensor:
        """List of token-id lists (already with [CLS]/[SEP]) -> fp32 [n, H] on the device."""
        if len(seqs) == 0:
            return torch.empty((0, self.config.hidden_size), dtype=torch.float32, device=self.device)
        flat, cu, longest = _pack(seqs, self.device)
        return self.forward_packed(flat, cu, longest, normalize)

    # ------------------------------------------------- cross-encoder scoring
    def classify_packed(self, token_ids: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int,
                        stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """Packed pair sequences -> classification-head logits fp32 [n, n_labels] (device).  Needs a checkpoint with
        `classifier.*` weights (XLMRobertaForSequenceClassification layout, e.g. bge-reranker-large)."""
        if self._head is None:
            raise _native.NativeError("this checkpoint has no classifier head (classifier.dense / classifier.out_proj)")
        n, T, dev = cu_seqlens.numel() - 1, token_ids.numel(), self.device
        with torch.cuda.device(dev):
            st = stream if stream is not None else torch.cuda.current_stream(dev)
            with torch.cuda.stream(st):
                logits = torch.empty((n, self.n_labels), dtype=torch.float32, device=dev)
                ws_bytes = self.workspace_bytes(T)
                ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
                rc = self._lib.crag_encoder_classify(C.byref(self._model), C.byref(self._head), token_ids.data_ptr(),
                                                     cu_seqlens.data_ptr(), n, T, int(max_seqlen), logits.data_ptr(),
                                                     ws.data_ptr(), ws_bytes, st.cuda_stream)
                _native.check(rc, "crag_encoder_classify")
        return logits

    def classify_token_lists(self, seqs: Sequence[Sequence[int]]) -> torch.Tensor:
        if len(seqs) == 0:
            return torch.empty((0, self.n_labels), dtype=torch.float32, device=self.device)
        flat, cu, longest = _pack(seqs, self.device)
        return self.classify_packed(flat, cu, longest)


def _flatten(seqs: Sequence[Sequence[int]]):
    """Token-id lists -> (flat int32 [T], cu_seqlens int32 [n + 1], longest) as numpy, without a Python-level pass
    over the tokens (the per-token list comprehension cost ~1 ms per 16k tokens, 5-8 % of an encode step)."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    if lens.size == 0 or int(lens.min()) < 1:
        raise ValueError("empty token sequence")
    flat = np.fromiter(itertools.chain.from_iterable(seqs), dtype=np.int32, count=int(lens.sum()))
    cu = np.zeros(len(seqs) + 1, dtype=np.int32)
    np.cumsum(lens, out=cu[1:])
    return flat, cu, int(lens.max())


def _pack(seqs: Sequence[Sequence[int]], device: torch.device):
    flat, cu, longest = _flatten(seqs)
    return (torch.from_numpy(flat).pin_memory().to(device, non_blocking=True),
            torch.from_numpy(cu).pin_memory().to(device, non_blocking=True), longest)


def random_head_state_dict(cfg: EncoderConfig, n_labels: int = 1, seed: int = 0, std: float = 0.02, device="cpu"):
    """XLMRobertaClassificationHead-shaped random weights (classifier.dense / classifier.out_proj)."""
    g = torch.Generator(device=device).manual_seed(seed + 7919)
    H = cfg.hidden_size
    w = lambda *shape: torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std
    return {"classifier.dense.weight": w(H, H), "classifier.dense.bias": w(H),
            "classifier.out_proj.weight": w(n_labels, H), "classifier.out_proj.bias": w(n_labels)}


def random_state_dict(cfg: EncoderConfig, seed: int = 0, std: float = 0.02, device="cpu") -> Dict[str, torch.Tensor]:
    """BertModel-shaped random weights (HF init: N(0, std), zero bias, LN gamma 1 / beta 0)."""
    g = torch.Generator(device=device).manual_seed(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size

    def w(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std

    z = lambda n: torch.zeros(n, device=device)
    o = lambda n: torch.ones(n, device=device)
    sd = {
        "embeddings.word_embeddings.weight": w(cfg.vocab_size, H),
        "embeddings.position_embeddings.weight": w(cfg.max_position_embeddings, H),
        "embeddings.token_type_embeddings.weight": w(cfg.type_vocab_size, H),
        "embeddings.LayerNorm.weight": o(H), "embeddings.LayerNorm.bias": z(H),
    }
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        for n in ("query", "key", "value"):
            sd[p + f"attention.self.{n}.weight"] = w(H, H)
            sd[p + f"attention.self.{n}.bias"] = z(H)
        sd[p + "attention.output.dense.weight"] = w(H, H)
        sd[p + "attention.output.dense.bias"] = z(H)
        sd[p + "attention.output.LayerNorm.weight"] = o(H)
        sd[p + "attention.output.LayerNorm.bias"] = z(H)
        sd[p + "intermediate.dense.weight"] = w(I, H)
        sd[p + "intermediate.dense.bias"] = z(I)
        sd[p + "output.dense.weight"] = w(H, I)
        sd[p + "output.dense.bias"] = z(H)
        sd[p + "output.LayerNorm.weight"] = o(H)
        sd[p + "output.LayerNorm.bias"] = z(H)
    return sd
