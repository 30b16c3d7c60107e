"""The torch-fp32 encoder oracle reproduces the embeddings the reference's BGEEmbeddingModel produced (through HF
BertModel) for the synthetic checkpoint: tests/golden/encoder_golden.npz, made by make_golden_encoder.py."""
import os

import numpy as np
import pytest
import torch

from comorag_b200.encoder import EncoderConfig
from oracle import encoder_oracle as eo

HERE = os.path.dirname(__file__)
CKPT = os.path.join(HERE, "golden", "bge-tiny-synth")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "encoder_golden.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def model():
    from safetensors.torch import load_file
    return EncoderConfig.from_hf_json(os.path.join(CKPT, "config.json")), load_file(os.path.join(CKPT, "model.safetensors"))


def test_batch_encode_embeddings_match_reference(gold, model):
    cfg, sd = model
    seqs = [t.tolist() for t in gold["token_ids"]]
    emb = eo.encode_token_lists(sd, cfg, seqs, batch_size=4).numpy()   # the golden run used embedding_batch_size=4
    np.testing.assert_allclose(emb, gold["emb_batch"], rtol=0, atol=2e-6)
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)


def test_positional_encode_has_no_instruction_prefix(gold, model):
    cfg, sd = model
    seqs = [t.tolist() for t in gold["token_ids_plain"]]
    emb = eo.encode_token_lists(sd, cfg, seqs).numpy()
    np.testing.assert_allclose(emb, gold["emb_encode"], rtol=0, atol=2e-6)


def test_reference_ignores_the_instruction_kwarg(gold):
    """batch_encode(q, instruction=...) and encode_queries(q) equal the passage-instruction rows."""
    n_c, n_q = int(gold["n_chunks"]), int(gold["n_questions"])
    np.testing.assert_allclose(gold["emb_query"], gold["emb_batch"][n_c:n_c + n_q], atol=1e-6)
    np.testing.assert_allclose(gold["emb_encode_queries"], gold["emb_batch"][n_c:n_c + n_q], atol=1e-6)


def test_padding_does_not_change_a_row(model):
    cfg, sd = model
    a = eo.encode_token_lists(sd, cfg, [[2, 10, 11, 12, 3]])
    b = eo.encode_token_lists(sd, cfg, [[2, 10, 11, 12, 3], [2] + list(range(5, 200)) + [3]])
    assert torch.allclose(a[0], b[0], atol=1e-6)


def test_oracle_matches_hf_xlm_roberta_positions():
    """XLM-R family (bge-m3 / bge-reranker backbones): position ids start at padding_idx + 1 = 2 and there is one
    token type.  The oracle with position_offset=2 equals HF's XLMRobertaModel on a random small checkpoint."""
    from transformers import XLMRobertaConfig, XLMRobertaModel
    hf_cfg = XLMRobertaConfig(vocab_size=300, hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                              max_position_embeddings=130, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5)
    torch.manual_seed(3)
    hf = XLMRobertaModel(hf_cfg, add_pooling_layer=False).eval()
    sd = {k: v for k, v in hf.state_dict().items()}
    cfg = EncoderConfig(64, 2, 2, 128, 300, max_position_embeddings=130, type_vocab_size=1, layer_norm_eps=1e-5, position_offset=2)
    seqs = [[0, 17, 45, 99, 2], [0] + list(range(10, 60)) + [2]]
    got = eo.encode_token_lists(sd, cfg, seqs, pad_id=1)
    L = max(len(s) for s in seqs)
    ids = torch.full((2, L), 1, dtype=torch.long)
    mask = torch.zeros((2, L), dtype=torch.long)
    for i, s_ in enumerate(seqs):
        ids[i, :len(s_)] = torch.tensor(s_)
        mask[i, :len(s_)] = 1
    with torch.no_grad():
        out = hf(input_ids=ids, attention_mask=mask).last_hidden_state
    want = torch.nn.functional.normalize(eo.mean_pooling(out, mask), dim=1)
    assert float((got - want).abs().max()) < 2e-6


def test_classifier_oracle_matches_hf_sequence_classification():
    """The cross-encoder oracle (encoder + head on <s>) equals HF's XLMRobertaForSequenceClassification (the
    bge-reranker-* architecture) on a random small checkpoint, padded batch included."""
    from transformers import XLMRobertaConfig, XLMRobertaForSequenceClassification
    hf_cfg = XLMRobertaConfig(vocab_size=300, hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                              max_position_embeddings=130, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5, num_labels=1)
    torch.manual_seed(11)
    hf = XLMRobertaForSequenceClassification(hf_cfg).eval()
    full = hf.state_dict()
    sd = {k[len("roberta."):]: v for k, v in full.items() if k.startswith("roberta.")}
    head = {k: v for k, v in full.items() if k.startswith("classifier.")}
    cfg = EncoderConfig(64, 2, 2, 128, 300, max_position_embeddings=130, type_vocab_size=1, layer_norm_eps=1e-5, position_offset=2)
    seqs = [[0, 17, 45, 2, 2, 99, 120, 2], [0] + list(range(10, 40)) + [2, 2] + list(range(50, 90)) + [2]]
    got = eo.classifier_logits(sd, head, cfg, seqs, pad_id=1)
    L = max(len(s) for s in seqs)
    ids = torch.full((2, L), 1, dtype=torch.long)
    mask = torch.zeros((2, L), dtype=torch.long)
    for i, s_ in enumerate(seqs):
        ids[i, :len(s_)] = torch.tensor(s_)
        mask[i, :len(s_)] = 1
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask).logits
    assert got.shape == want.shape == (2, 1)
    assert float((got - want).abs().max()) < 2e-6
