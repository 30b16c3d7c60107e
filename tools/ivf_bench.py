"""BASELINE config 4 on one GPU: IVF residual-IP search timing + recall against the flat exact search.

    python tools/ivf_bench.py [--rows 12500000 --dim 768 --nlist 4096 --nprobe 32 --k 100 --nq 32]

(12.5M x 768 is one rank's share of the 100M x 768 corpus over 8 GPUs.)  Synthetic clustered unit vectors; the
IVF index is built on the device (k-means on a sample), then `steps` searches of one 32-query block are timed with
CUDA events, next to the flat scan of the same rows (DenseIndex) for the speed-up and the recall@k.  Prints one JSON
object.  Written in round 1 after the GPU budget ended: NOT yet run at these sizes.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def clustered(rows, dim, n_dirs, seed, device, sigma=0.5):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    dirs = torch.nn.functional.normalize(torch.randn(n_dirs, dim, generator=g, device=device), dim=1)
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
    from comorag_b200.index import DenseIndex
    from comorag_b200.ivf import IVFIndex

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    x, dirs = clustered(args.rows, args.dim, 4 * args.nlist, 1, dev)
    g = torch.Generator(device=dev).manual_seed(2)
    q = torch.nn.functional.normalize(
        dirs[torch.randint(0, dirs.shape[0], (args.nq,), generator=g, device=dev)]
        + (0.5 / args.dim ** 0.5) * torch.randn(args.nq, args.dim, generator=g, device=dev), dim=1).to(torch.bfloat16)

    t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0[0].record()
    ivf = IVFIndex.build(x, args.nlist, iters=args.kmeans_iters, seed=0, train_rows=args.train_rows)
    t0[1].record()
    torch.cuda.synchronize()
    build_ms = t0[0].elapsed_time(t0[1])
    flat = DenseIndex.from_tensor(x.to(torch.bfloat16))
    del x
    torch.cuda.empty_cache()

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, out

    ivf_ms, (ids, scores, mm, (p_ids, _)) = timed(lambda: ivf.search_device(q, args.nprobe, args.k))
    flat_ms, (f_ids, f_scores, _) = timed(lambda: flat.search_device(q, args.k))
    ids, f_ids = ids.cpu().numpy(), f_ids.cpu().numpy()
    recall = float(np.mean([len(set(a[a >= 0].tolist()) & set(b.tolist())) / args.k for a, b in zip(ids, f_ids)]))
    list_rows = ivf.list_rows.cpu().numpy().astype(np.int64)
    probed = np.unique(p_ids.cpu().numpy())
    probed_rows = int(((list_rows[probed[probed >= 0]] + 127) // 128 * 128).sum())
    print(json.dumps({
        "workload": f"IVF-{args.nlist} residual-IP, {args.rows}x{args.dim} bf16, nprobe {args.nprobe}, top-{args.k}, {args.nq} queries, 1 GPU",
        "ivf_ms_per_step": round(ivf_ms, 4), "ivf_queries_per_s": round(args.nq / (ivf_ms * 1e-3), 1),
        "flat_ms_per_step": round(flat_ms, 4), "flat_queries_per_s": round(args.nq / (flat_ms * 1e-3), 1),
        "speedup_vs_flat": round(flat_ms / ivf_ms, 2), f"recall_at_{args.k}_vs_flat": round(recall, 4),
        "probed_rows": probed_rows, "probed_fraction": round(probed_rows / max(args.rows, 1), 4),
        "probed_bytes_per_step": probed_rows * args.dim * 2,
        "achieved_GBps_on_probed_rows": round(probed_rows * args.dim * 2 / (ivf_ms * 1e-3) / 1e9, 1),
        "build_ms": round(build_ms, 1), "largest_list_rows": int(list_rows.max()), "empty_lists": int((list_rows == 0).sum()),
        "data": "synthetic clustered unit vectors",
    }))


if __name__ == "__main__":
    main()
