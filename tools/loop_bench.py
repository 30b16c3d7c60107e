"""BASELINE config 5: the probe -> retrieve -> rerank cycle of ComoRAG's iterative loop, on one GPU
(`python tools/loop_bench.py`) or on N GPUs of one box (`python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 tools/loop_bench.py`: the index is row-sharded with one all-gather per search, every rank
encodes the same probes, the rerank pairs are split by rank with one all-gather of the logits).

Per cycle (ComoRAG.py:456-554 `tri_retrieve` is called once per probe; here the 32 probes of a cycle arrive as one
wave):
  1. encode the probes         (bge-large shape, short queries)            -> [nq, 1024] on the device
  2. search the passage index  (rows x 1024 bf16, exact top-k)             -> ids to the host
  3. cross-encoder rerank      (bge-reranker-large shape: XLM-R-large + classification head, nq*k pairs of
                                query + passage tokens up to 512)          -> logits to the host
Weights and corpus are synthetic (random init / seeded unit rows); token ids are random.  Prints one JSON object
with per-stage device milliseconds (CUDA events on the launching stream) and whole-loop probes/s measured by the
host clock around all cycles (host<->device copies included).  Not a bench.py line: config 5's metric has no
reference arithmetic for stage 3 (the reference's rerank is an LLM prompt).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=32)
    ap.add_argument("--cycles", type=int, default=5)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--query-tokens", type=int, default=24)
    ap.add_argument("--pair-tokens", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--rerank-vocab", type=int, default=250002)
    args = ap.parse_args()

    import numpy as np
    import torch
    from bench import make_shard
    from comorag_b200.encoder import BertEncoderB200, EncoderConfig, random_head_state_dict, random_state_dict
    import torch.distributed as dist
    from comorag_b200.dist import ShardedIndex, shard_bounds, sharded_rerank
    from comorag_b200.index import DenseIndex
    from comorag_b200.rerank import CrossEncoderReranker

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    offs = shard_bounds(args.rows, world)
    index = ShardedIndex(DenseIndex.from_tensor(make_shard(offs[rank + 1] - offs[rank], args.dim, 1234 + rank, dev),
                                                row_offset=offs[rank]))
    enc = BertEncoderB200.random_init(EncoderConfig.bge_large(), seed=0, device=dev)
    rcfg = EncoderConfig(1024, 24, 16, 4096, args.rerank_vocab, max_position_embeddings=514, type_vocab_size=1,
                         layer_norm_eps=1e-5, position_offset=2)
    rsd = random_state_dict(rcfg, seed=1, device=dev)
    rsd.update(random_head_state_dict(rcfg, n_labels=1, seed=1, std=0.02, device=dev))
    reranker = CrossEncoderReranker("synthetic", encoder=BertEncoderB200(rcfg, rsd, dev), tokenizer=object(),
                                    max_length=512, token_budget=16384)
    del rsd
    torch.cuda.empty_cache()

    rng = np.random.default_rng(0)
    st = torch.cuda.current_stream(dev)

    def cycle(ev):
        probes = [[101] + rng.integers(1000, 30000, args.query_tokens - 2).tolist() + [102] for _ in range(args.nq)]
        ev[0].record(st)
        q = enc.encode_token_lists(probes)                                   # H2D of the token ids inside
        ev[1].record(st)
        ids, scores, minmax = index.search_device(q.to(torch.bfloat16), args.k)
        ids_host = ids.cpu()                                                 # the caller needs the passage ids
        ev[2].record(st)
        # the passages' tokens would come from the chunk store; synthetic ids of the configured pair length
        pairs = [[0] + rng.integers(5, args.rerank_vocab, args.pair_tokens - 2).tolist() + [2]
                 for _ in range(args.nq * args.k)]
        ev[3].record(st)
        if world == 1:
            logits = reranker.score_token_lists(pairs)                       # H2D ids, D2H logits inside
        else:
            logits = sharded_rerank(reranker.score_token_lists, pairs, device=dev).cpu().numpy()
        ev[4].record(st)
        return ids_host, logits

    def events():
        return [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    for _ in range(max(args.warmup, 1)):
        cycle(events())
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    all_ev = [events() for _ in range(args.cycles)]
    t0 = time.perf_counter()
    for ev in all_ev:
        ids_host, logits = cycle(ev)
    barrier()
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    stage = np.zeros(4)
    for ev in all_ev:
        stage += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    stage /= args.cycles
    assert ids_host.shape == (args.nq, args.k) and logits.shape == (args.nq * args.k, 1) and np.isfinite(logits).all()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({
        "n_gpus": world,
        "workload": f"{args.nq} probes x {args.cycles} cycles, {args.rows}x{args.dim} bf16 index top-{args.k}, "
                    f"rerank {args.nq * args.k} pairs x {args.pair_tokens} tokens (XLM-R-large shape), {world} GPU(s)",
        "ms_per_cycle": {"encode_probes": round(float(stage[0]), 3), "search_and_ids_d2h": round(float(stage[1]), 3),
                         "host_pair_assembly": round(float(stage[2]), 3), "rerank": round(float(stage[3]), 3)},
        "wall_ms_per_cycle": round(wall * 1e3 / args.cycles, 3),
        "probes_per_s": round(args.nq * args.cycles / wall, 1),
        "rerank_pairs_per_s": round(args.nq * args.k / (float(stage[3]) * 1e-3), 1),
        "data": "synthetic (random-init weights, seeded unit-norm corpus, random token ids)",
    }))


if __name__ == "__main__":
    main()
