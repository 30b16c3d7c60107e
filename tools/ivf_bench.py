"""BASELINE config 4: IVF residual-IP search timing + recall against the flat exact search, on one GPU

    python tools/ivf_bench.py [--rows 12500000 --dim 768 --nlist 4096 --nprobe 32 --k 100 --nq 32]

or row-sharded over the N GPUs of one box (`--rows` is then PER RANK: 8 x 12.5M = the 100M x 768 of config 4):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/ivf_bench.py

Synthetic clustered unit vectors (every rank draws its own rows around the same cluster directions); the centroid
table is trained on rank 0's sample and broadcast, every rank lays out the residual lists of its own rows; `steps`
searches of one 32-query block are timed with CUDA events (max over ranks), next to the flat exact search of the same
rows (DenseIndex / ShardedIndex) for the speed-up and recall@k.  Rank 0 prints one JSON object.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def clustered(rows, dim, n_dirs, seed, device, sigma=0.5):
    import torch
    gd = torch.Generator(device=device).manual_seed(12345)        # the cluster directions are the same on every rank
    dirs = torch.nn.functional.normalize(torch.randn(n_dirs, dim, generator=gd, device=device), dim=1)
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((rows, dim), dtype=torch.float32, device=device)
    slab = 1 << 19
    for s in range(0, rows, slab):
        n = min(slab, rows - s)
        pick = torch.randint(0, n_dirs, (n,), generator=g, device=device)
        out[s:s + n] = torch.nn.functional.normalize(
            dirs[pick] + (sigma / dim ** 0.5) * torch.randn(n, dim, generator=g, device=device), dim=1)
    return out, dirs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12_500_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--nq", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--kmeans-iters", type=int, default=6)
    ap.add_argument("--train-rows", type=int, default=1 << 20)
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    from comorag_b200.dist import ShardedIndex
    from comorag_b200.index import DenseIndex
    from comorag_b200.ivf import IVFIndex, ShardedIVF, spherical_kmeans

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    x, dirs = clustered(args.rows, args.dim, 4 * args.nlist, 1 + rank, dev)
    g = torch.Generator(device=dev).manual_seed(2)
    q = torch.nn.functional.normalize(
        dirs[torch.randint(0, dirs.shape[0], (args.nq,), generator=g, device=dev)]
        + (0.5 / args.dim ** 0.5) * torch.randn(args.nq, args.dim, generator=g, device=dev), dim=1).to(torch.bfloat16)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t0[0].record()
    centroids = None
    if world > 1:   # one table for the whole index: trained on rank 0's sample, broadcast
        g = torch.Generator(device=dev).manual_seed(1)
        centroids = torch.empty((args.nlist, args.dim), dtype=torch.float32, device=dev)
        if rank == 0:
            sample = x[torch.randperm(x.shape[0], generator=g, device=dev)[:args.train_rows]]
            centroids.copy_(spherical_kmeans(sample, args.nlist, iters=args.kmeans_iters, seed=0))
        dist.broadcast(centroids, 0)
    local = IVFIndex.build(x, args.nlist, iters=args.kmeans_iters, seed=0, train_rows=args.train_rows, centroids=centroids,
                           row_offset=rank * args.rows)
    t0[1].record()
    sync()
    build_ms = max_ranks(t0[0].elapsed_time(t0[1]))
    ivf = ShardedIVF(local) if world > 1 else None
    flat_local = DenseIndex.from_tensor(x.to(torch.bfloat16), row_offset=rank * args.rows)
    flat = ShardedIndex(flat_local) if world > 1 else flat_local
    del x
    torch.cuda.empty_cache()

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        for _ in range(args.steps):
            out = fn()
        e1.record()
        sync()
        return max_ranks(e0.elapsed_time(e1)) / args.steps, tuple(t.clone() if torch.is_tensor(t) else t for t in out)

    _, _, _, (p_ids, _) = local.search_device(q, args.nprobe, args.k)
    if world > 1:
        ivf_ms, (ids, scores, mm) = timed(lambda: ivf.search_device(q, args.nprobe, args.k))
    else:
        ivf_ms, (ids, scores, mm) = timed(lambda: local.search_device(q, args.nprobe, args.k)[:3])
    flat_ms, (f_ids, f_scores, _) = timed(lambda: flat.search_device(q, args.k))
    ids, f_ids = ids.cpu().numpy(), f_ids.cpu().numpy()
    recall = float(np.mean([len(set(a[a >= 0].tolist()) & set(b.tolist())) / args.k for a, b in zip(ids, f_ids)]))
    list_rows = local.list_rows.cpu().numpy().astype(np.int64)
    probed = np.unique(p_ids.cpu().numpy())
    probed_rows = int(((list_rows[probed[probed >= 0]] + 127) // 128 * 128).sum())     # this rank's probed rows
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({
        "workload": f"IVF-{args.nlist} residual-IP, {world * args.rows}x{args.dim} bf16 ({args.rows} rows per rank), nprobe {args.nprobe}, "
                    f"top-{args.k}, {args.nq} queries, {world} GPU(s)",
        "n_gpus": world,
        "ivf_ms_per_step": round(ivf_ms, 4), "ivf_queries_per_s": round(args.nq / (ivf_ms * 1e-3), 1),
        "flat_ms_per_step": round(flat_ms, 4), "flat_queries_per_s": round(args.nq / (flat_ms * 1e-3), 1),
        "speedup_vs_flat": round(flat_ms / ivf_ms, 2), f"recall_at_{args.k}_vs_flat": round(recall, 4),
        "probed_rows_rank0": probed_rows, "probed_fraction": round(probed_rows / max(args.rows, 1), 4),
        "probed_bytes_per_step_rank0": probed_rows * args.dim * 2,
        "achieved_GBps_on_probed_rows_per_gpu": round(probed_rows * args.dim * 2 / (ivf_ms * 1e-3) / 1e9, 1),
        "build_ms": round(build_ms, 1), "largest_list_rows": int(list_rows.max()), "empty_lists": int((list_rows == 0).sum()),
        "data": "synthetic clustered unit vectors",
    }))


if __name__ == "__main__":
    main()
