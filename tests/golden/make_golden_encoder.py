"""Generate the encoder / EmbeddingStore golden fixtures by RUNNING THE REFERENCE in the build container.

    python tests/golden/make_golden_encoder.py

Writes (all small, committed):
  tests/golden/bge-tiny-synth/      synthetic BERT checkpoint (config.json, model.safetensors, tokenizer files);
                                    the directory name contains "bge-" so the reference's factory picks
                                    BGEEmbeddingModel (embedding_model/__init__.py:10-12)
  tests/golden/encoder_golden.npz   texts, the token ids the reference's tokenizer call produced
                                    (BGEEmbedding.py:112-117), and the embeddings returned by the reference's
                                    BGEEmbeddingModel.batch_encode / .encode (BGEEmbedding.py:131-185, 92-129)
  tests/golden/store_golden.json    hash ids / lookup results of the reference's EmbeddingStore
                                    (embedding_store.py:14-167) after insert_strings on the cinderella chunks
The cinderella chunk texts (the reference's only bundled data, BASELINE config 1) are stored inside the
fixture because the GPU box has no /root/reference.
"""
import json
import os
import re
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _reference_harness import REFERENCE_ROOT, import_reference  # noqa: E402

CKPT = os.path.join(HERE, "bge-tiny-synth")
INSTRUCTION = "Generate a representation for this sentence to retrieve relevant articles:"


def build_checkpoint(chunks):
    from transformers import BertConfig, BertModel, BertTokenizerFast
    if os.path.isdir(CKPT):
        shutil.rmtree(CKPT)
    os.makedirs(CKPT)
    words = {}
    for t in chunks + [INSTRUCTION]:
        for w in re.findall(r"[a-z]+|[^a-z\s]", t.lower()):
            words[w] = words.get(w, 0) + 1
    specials = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = sorted({c for w in words for c in w} | set("abcdefghijklmnopqrstuvwxyz0123456789.,;:!?'\"-()"))
    pieces = chars + ["##" + c for c in chars if c.isalnum()]
    common = [w for w, c in sorted(words.items(), key=lambda kv: (-kv[1], kv[0])) if len(w) > 1][:600]
    vocab = specials + pieces + [w for w in common if w not in pieces]
    with open(os.path.join(CKPT, "vocab.txt"), "w") as f:
        f.write("\n".join(vocab) + "\n")
    tok = BertTokenizerFast(vocab={w: i for i, w in enumerate(vocab)}, do_lower_case=True)  # transformers 5.x signature
    tok.save_pretrained(CKPT)
    cfg = BertConfig(vocab_size=len(vocab), hidden_size=128, num_hidden_layers=2, num_attention_heads=4,
                     intermediate_size=256, max_position_embeddings=512, type_vocab_size=2)
    torch.manual_seed(0)
    model = BertModel(cfg)
    # HF's default init (std 0.02) collapses all embeddings to one direction (any two texts: cosine > 0.98), which
    # makes every ranking noise.  Give the WORD embeddings unit scale and keep the position / layer weights small:
    # the residual stream is then dominated by what the text says, document-document cosines spread over 0.6-0.9
    # and query-document raw score ranges are 0.1-0.3 -- rankings that the end-to-end parity test can actually hold
    # an implementation to (tests/test_e2e_cinderella.py).
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "word_embeddings" in n:
                p.normal_(0.0, 1.0)
            elif "position_embeddings" in n or "token_type_embeddings" in n:
                p.normal_(0.0, 0.02)
            elif p.dim() == 2:
                p.normal_(0.0, 0.03)
            elif "LayerNorm.weight" in n:
                p.fill_(1.0).add_(0.1 * torch.randn_like(p))
            else:
                p.normal_(0.0, 0.02)
    model.save_pretrained(CKPT, safe_serialization=True)
    return len(vocab)


def main():
    ref = import_reference()
    data = os.path.join(REFERENCE_ROOT, "dataset/cinderella/cinderella_1")
    chunks = [json.loads(l)["contents"] for l in open(os.path.join(data, "corpus.jsonl")) if l.strip()]
    questions = [json.loads(l)["question"] for l in open(os.path.join(data, "qas.jsonl")) if l.strip()]
    build_checkpoint(chunks)
    cfg = ref.BaseConfig(embedding_model_name=CKPT, embedding_batch_size=4, embedding_max_seq_len=512)
    model = ref.OracleBGE(global_config=cfg, embedding_model_name=CKPT)

    texts = chunks + questions + ["a", "the king's son", "Cinderella went to the ball. " * 40]
    emb_batch = model.batch_encode(texts)                                   # BGEEmbedding.py:131-185 (loops over batches of 4)
    emb_query = np.concatenate([model.batch_encode(q, instruction="ignored by the reference", norm=True) for q in questions])
    emb_encode = model.encode(texts[:3]).cpu().numpy()                      # positional surface, no instruction prefix
    emb_q_api = model.encode_queries(questions)
    # token ids exactly as the reference's _encode builds them (instruction + text, no separator)
    tok_ids = [model.tokenizer(INSTRUCTION + t, truncation=True, max_length=512)["input_ids"] for t in texts]
    tok_ids_plain = [model.tokenizer(t, truncation=True, max_length=512)["input_ids"] for t in texts[:3]]
    np.savez_compressed(
        os.path.join(HERE, "encoder_golden.npz"),
        texts=np.array(texts, dtype=object), n_chunks=len(chunks), n_questions=len(questions),
        token_ids=np.array([np.array(t, dtype=np.int32) for t in tok_ids], dtype=object),
        token_ids_plain=np.array([np.array(t, dtype=np.int32) for t in tok_ids_plain], dtype=object),
        emb_batch=emb_batch.astype(np.float32), emb_query=emb_query.astype(np.float32),
        emb_encode=emb_encode.astype(np.float32), emb_encode_queries=emb_q_api.astype(np.float32))

    # EmbeddingStore golden: reference store over the cinderella chunks
    tmp = tempfile.mkdtemp()
    try:
        store = ref.EmbeddingStore(model, os.path.join(tmp, "chunk_embeddings"), 4, "chunk")
        r1 = store.insert_strings(chunks)
        r2 = store.insert_strings(chunks[:2] + ["a brand new chunk"])
        r3 = store.insert_strings(chunks[:2])
        missing = store.get_missing_string_hash_ids([chunks[0], "never seen"])
        import pandas as pd
        import pyarrow.parquet as pq
        schema = pq.read_schema(store.filename)
        gold = {
            "namespace": "chunk", "filename": os.path.basename(store.filename),
            "insert_returns": [repr(r1), repr(r2), repr(r3)],
            "hash_ids": store.get_all_ids(), "texts": store.texts,
            "missing": missing, "row0": store.get_row(store.hash_ids[0]),
            "hash_id_to_order": store.get_hash_id_to_order(),
            "parquet_schema": {n: str(schema.field(n).type) for n in schema.names},
            "embedding_row0_first8": [float(x) for x in store.get_embedding(store.hash_ids[0])[:8]],
            "get_embeddings_shape": list(store.get_embeddings(store.hash_ids[:3]).shape),
            "get_embeddings_dtype": str(store.get_embeddings(store.hash_ids[:3]).dtype),
            "reloaded_embedding_type": None,
        }
        store2 = ref.EmbeddingStore(model, os.path.join(tmp, "chunk_embeddings"), 4, "chunk")
        gold["reloaded_hash_ids"] = store2.get_all_ids()
        gold["reloaded_embedding_type"] = type(store2.embeddings[0]).__name__ + ":" + str(store2.embeddings[0].dtype)
        np.save(os.path.join(HERE, "store_golden_embeddings.npy"), np.array(store.embeddings, dtype=np.float32))
        with open(os.path.join(HERE, "store_golden.json"), "w") as f:
            json.dump(gold, f, indent=1)
    finally:
        shutil.rmtree(tmp)
    print("wrote encoder_golden.npz, store_golden.json; vocab", len(model.tokenizer), "emb", emb_batch.shape)


if __name__ == "__main__":
    main()
