"""CPU restatement (plain torch fp32 ops) of the reference's encode path.

TEST INFRASTRUCTURE ONLY: the product path never imports this module.

The reference's encoder arithmetic lives in a third-party dependency that is not
vendored under the reference tree: `transformers` (pinned ==4.51.3 in
requirements.txt:176; 5.5.0 is what this image has) -- `AutoModel` resolves to
`BertModel` for BGE-en checkpoints and is called at BGEEmbedding.py:51-52,120.
This file restates BertModel's published forward (post-LN encoder, learned
absolute positions, exact-erf GELU, softmax(QK^T/sqrt(dh)+mask)V) together with
the reference's own mean_pooling (BGEEmbedding.py:15-28) and F.normalize
(BGEEmbedding.py:126-127).  It is pinned by tests/golden/encoder_*.npz, which
tests/golden/make_golden_encoder.py produced by running the reference's
BGEEmbeddingModel (through HF BertModel) on a synthetic checkpoint in the build
container; tests/test_oracle_encoder.py replays them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def mean_pooling(token_embeddings: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """BGEEmbedding.py:15-28."""
    token_embeddings = token_embeddings.masked_fill(~mask[..., None].bool(), 0.0)
    return token_embeddings.sum(dim=1) / mask.sum(dim=1)[..., None]


def bert_forward(sd: Dict[str, torch.Tensor], cfg, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """HF BertModel.forward -> last_hidden_state [b, L, H]; fp32, padded batch + additive key mask.

    `cfg` needs hidden_size, num_hidden_layers, num_attention_heads, layer_norm_eps, position_offset.
    """
    H, nh = cfg.hidden_size, cfg.num_attention_heads
    dh = H // nh
    eps = cfg.layer_norm_eps
    b, L = input_ids.shape
    g = lambda k: sd[k].float()
    pos = torch.arange(L, device=input_ids.device) + getattr(cfg, "position_offset", 0)
    x = F.embedding(input_ids, g("embeddings.word_embeddings.weight")) \
        + g("embeddings.position_embeddings.weight")[pos][None] \
        + g("embeddings.token_type_embeddings.weight")[0][None, None]
    x = F.layer_norm(x, (H,), g("embeddings.LayerNorm.weight"), g("embeddings.LayerNorm.bias"), eps)
    bias = (1.0 - attention_mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        q = F.linear(x, g(p + "attention.self.query.weight"), g(p + "attention.self.query.bias"))
        k = F.linear(x, g(p + "attention.self.key.weight"), g(p + "attention.self.key.bias"))
        v = F.linear(x, g(p + "attention.self.value.weight"), g(p + "attention.self.value.bias"))
        q, k, v = (t.view(b, L, nh, dh).transpose(1, 2) for t in (q, k, v))
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + bias, dim=-1)
        ctx = (att @ v).transpose(1, 2).reshape(b, L, H)
        y = F.linear(ctx, g(p + "attention.output.dense.weight"), g(p + "attention.output.dense.bias"))
        x = F.layer_norm(y + x, (H,), g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias"), eps)
        h = F.gelu(F.linear(x, g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")))
        y = F.linear(h, g(p + "output.dense.weight"), g(p + "output.dense.bias"))
        x = F.layer_norm(y + x, (H,), g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"), eps)
    return x


def encode_token_lists(sd: Dict[str, torch.Tensor], cfg, seqs: Sequence[Sequence[int]], normalize: bool = True,
                       batch_size: int = 32, pad_id: int = 0) -> torch.Tensor:
    """BGEEmbeddingModel._encode after tokenisation (BGEEmbedding.py:112-127): pad to longest, forward,
    masked mean pool, L2 normalise; batches of `batch_size` as batch_encode does (BGEEmbedding.py:168-176)."""
    dev = next(iter(sd.values())).device
    outs: List[torch.Tensor] = []
    with torch.no_grad():
        for s0 in range(0, len(seqs), batch_size):
            chunk = seqs[s0:s0 + batch_size]
            L = max(len(s) for s in chunk)
            ids = torch.full((len(chunk), L), pad_id, dtype=torch.long, device=dev)
            mask = torch.zeros((len(chunk), L), dtype=torch.long, device=dev)
            for i, s in enumerate(chunk):
                ids[i, :len(s)] = torch.tensor(list(s), dtype=torch.long, device=dev)
                mask[i, :len(s)] = 1
            e = mean_pooling(bert_forward(sd, cfg, ids, mask), mask)
            if normalize:
                e = F.normalize(e, p=2, dim=1)
            outs.append(e)
    return torch.cat(outs, 0) if outs else torch.empty((0, cfg.hidden_size), device=dev)


def classifier_logits(sd: Dict[str, torch.Tensor], head: Dict[str, torch.Tensor], cfg, seqs: Sequence[Sequence[int]],
                      batch_size: int = 32, pad_id: int = 1) -> torch.Tensor:
    """Cross-encoder score of packed (query, passage) sequences: HF XLMRobertaForSequenceClassification's published
    forward -- encoder, then XLMRobertaClassificationHead on the first token (<s>): out_proj(tanh(dense(h[:, 0]))).
    The reference has no such arithmetic (rerank.py:97-123 is an LLM prompt), so this is pinned
    against the `transformers` implementation itself (tests/test_oracle_encoder.py), not against the reference."""
    dev = next(iter(sd.values())).device
    outs: List[torch.Tensor] = []
    with torch.no_grad():
        for s0 in range(0, len(seqs), batch_size):
            chunk = seqs[s0:s0 + batch_size]
            L = max(len(s) for s in chunk)
            ids = torch.full((len(chunk), L), pad_id, dtype=torch.long, device=dev)
            mask = torch.zeros((len(chunk), L), dtype=torch.long, device=dev)
            for i, s in enumerate(chunk):
                ids[i, :len(s)] = torch.tensor(list(s), dtype=torch.long, device=dev)
                mask[i, :len(s)] = 1
            h0 = bert_forward(sd, cfg, ids, mask)[:, 0]
            y = torch.tanh(F.linear(h0, head["classifier.dense.weight"].float(), head["classifier.dense.bias"].float()))
            outs.append(F.linear(y, head["classifier.out_proj.weight"].float(), head["classifier.out_proj.bias"].float()))
    return torch.cat(outs, 0)
