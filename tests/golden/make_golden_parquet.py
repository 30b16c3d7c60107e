"""Parquet interop fixture, produced in the build container with the REFERENCE's EmbeddingStore:

  1. the reference writes vdb_chunk.parquet (embedding_store.py:109-120, pandas -> pyarrow) for the golden chunk
     texts/embeddings -> committed as tests/golden/vdb_chunk_reference.parquet; tests/test_store_host.py loads it with
     OUR store and must see the same ids / texts / rows;
  2. our store writes the same rows; the reference's EmbeddingStore loads that file here and must see the same
     ids / texts / rows -> recorded in tests/golden/parquet_interop.json (the reference cannot run on the GPU box).

    python tests/golden/make_golden_parquet.py
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from _reference_harness import import_reference  # noqa: E402


class Replay:
    def __init__(self, texts, emb):
        self.table = dict(zip(texts, emb))
        self.embedding_dim = emb.shape[1]

    def batch_encode(self, texts, **kw):
        return np.stack([self.table[t] for t in texts]).astype(np.float32)


def main():
    ref = import_reference()
    from comorag_b200.embedding_store import EmbeddingStore as OurStore
    gold = json.load(open(os.path.join(HERE, "store_golden.json")))
    emb = np.load(os.path.join(HERE, "store_golden_embeddings.npy"))
    texts = gold["texts"]
    tmp = tempfile.mkdtemp()
    try:
        # 1. reference writes
        rs = ref.EmbeddingStore(Replay(texts, emb), os.path.join(tmp, "ref"), 4, "chunk")
        rs.insert_strings(texts)
        shutil.copy(rs.filename, os.path.join(HERE, "vdb_chunk_reference.parquet"))
        # 2. ours writes, reference reads
        ours = OurStore(Replay(texts, emb), os.path.join(tmp, "ours"), 4, "chunk")
        ours.insert_strings(texts)
        back = ref.EmbeddingStore(Replay(texts, emb), os.path.join(tmp, "ours"), 4, "chunk")
        result = {
            "reference_read_ours": {
                "hash_ids_equal": back.get_all_ids() == gold["hash_ids"],
                "texts_equal": back.texts == texts,
                "embeddings_equal": bool(np.array_equal(np.array(back.embeddings, dtype=np.float32), emb)),
                "embedding_row_type": type(back.embeddings[0]).__name__ + ":" + str(back.embeddings[0].dtype),
                "get_embeddings_shape": list(np.asarray(back.get_embeddings(back.hash_ids[:3])).shape),
                "second_insert_is_noop": back.insert_strings(texts[:2]) == {} and len(back.hash_ids) == len(texts),
            },
            "reference_file": "vdb_chunk_reference.parquet",
            "rows": len(texts), "dim": int(emb.shape[1]),
        }
        assert all(v is True for k, v in result["reference_read_ours"].items() if k.endswith("equal") or k.endswith("noop")), result
        with open(os.path.join(HERE, "parquet_interop.json"), "w") as f:
            json.dump(result, f, indent=1)
        print(json.dumps(result, indent=1))
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
