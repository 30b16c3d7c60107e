"""The N>1 exchange step on CPU: world_size-2 gloo processes each hold a row shard, compute their local exact top-k
with the oracle, all-gather the packed partials (the same payload ShardedIndex sends over NCCL) and merge; the result
must equal the oracle's top-k over the unsharded corpus."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from comorag_b200.dist import merge_partials_reference, pack_partial, shard_bounds, unpack_partials
from oracle import search_oracle as so


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _corpus(n, d):
    g = torch.Generator().manual_seed(11)
    E = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).bfloat16().float().numpy()
    Q = torch.nn.functional.normalize(torch.randn(6, d, generator=g), dim=1).bfloat16().float().numpy()
    return E, Q


def _worker(rank, world, port, n, d, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E, Q = _corpus(n, d)
    offs = shard_bounds(n, world)
    ids, scores, mm, _ = so.topk_exact(E[offs[rank]:offs[rank + 1]], Q, k, row_offset=offs[rank])
    mine = pack_partial(torch.from_numpy(ids), torch.from_numpy(scores.astype(np.float32)), torch.from_numpy(mm.astype(np.float32)))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    g_ids, g_scores, g_mm = unpack_partials(torch.cat(gathered), world, Q.shape[0], k)
    oi, os_, om = merge_partials_reference(g_ids, g_scores, g_mm, k)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), ids=oi.numpy(), scores=os_.numpy(), mm=om.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(3001, 10), (9, 10)])
def test_two_rank_shard_merge_equals_global_topk(tmp_path, n, k):
    d, world = 64, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, d, k, str(tmp_path)), nprocs=world, join=True)
    E, Q = _corpus(n, d)
    want_i, want_s, want_mm, gaps = so.topk_exact(E, Q, k)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        so.assert_topk_matches(got["ids"], got["scores"].astype(np.float64), want_i, want_s, gaps, score_tol=1e-6)
        np.testing.assert_allclose(got["mm"], want_mm, atol=1e-6)


def _rerank_worker(rank, world, port, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from comorag_b200.dist import sharded_rerank
    pairs = [[0] + list(range(5, 5 + (i % 7) + 1)) + [2] for i in range(n_pairs)]
    seen = []

    def score(token_lists):                      # stand-in for CrossEncoderReranker.score_token_lists
        seen.append(len(token_lists))
        return np.array([[float(sum(t)) + 0.5 * len(t)] for t in token_lists], dtype=np.float32).reshape(-1, 1)

    got = sharded_rerank(score, pairs)
    np.savez(os.path.join(out_dir, f"rr{rank}.npz"), logits=got.numpy(), seen=np.array(seen))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [7, 8, 1])
def test_two_rank_pair_sharded_rerank(tmp_path, n_pairs):
    """Config 5's rerank across ranks: each rank scores only its slice of the pairs; one all-gather later every
    rank holds every logit in pair order (ragged last slice and a rank with nothing to score included)."""
    world = 2
    mp.spawn(_rerank_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
    pairs = [[0] + list(range(5, 5 + (i % 7) + 1)) + [2] for i in range(n_pairs)]
    want = np.array([[float(sum(t)) + 0.5 * len(t)] for t in pairs], dtype=np.float32)
    sizes = []
    for r in range(world):
        got = np.load(tmp_path / f"rr{r}.npz")
        np.testing.assert_array_equal(got["logits"], want)
        sizes.append(int(got["seen"][0]))
    assert sum(sizes) == n_pairs and max(sizes) - min(sizes) <= 1


def _rerank_worker_labels(rank, world, port, n_pairs, labels, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from comorag_b200.dist import sharded_rerank
    pairs = [[0, 5 + i, 2] for i in range(n_pairs)]

    def score(token_lists):   # a rank with nothing to score returns shape (0,): it cannot know the label count
        if not token_lists:
            return np.zeros((0,), dtype=np.float32)
        return np.array([[float(t[1]) * (j + 1) for j in range(labels)] for t in token_lists], dtype=np.float32)

    got = sharded_rerank(score, pairs)
    np.save(os.path.join(out_dir, f"rl{rank}.npy"), got.numpy())
    dist.destroy_process_group()


def test_rerank_label_count_is_agreed_across_ranks(tmp_path):
    """fewer pairs than ranks + more than one label -- the empty rank must not size the collective from
    its own (0, 1) output."""
    world, n_pairs, labels = 2, 1, 3
    mp.spawn(_rerank_worker_labels, args=(world, _free_port(), n_pairs, labels, str(tmp_path)), nprocs=world, join=True)
    want = np.array([[5.0, 10.0, 15.0]], dtype=np.float32)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"rl{r}.npy"), want)
