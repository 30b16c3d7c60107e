"""Self-consistency of the IVF oracle (oracle/ivf_oracle.py).  The reference has no IVF code, so these are the
invariants that define the semantic the engine's IVF path will be tested against: layout, the nprobe == nlist
degenerate case against exact search, monotone recall, and that the residual form is the more accurate one."""
import numpy as np
import torch

from oracle import ivf_oracle as ivf
from oracle import search_oracle as so


def _data(n=3000, d=64, nq=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    # clustered unit vectors (what an embedding corpus looks like): 40 directions + noise of norm ~0.5
    dirs = torch.nn.functional.normalize(torch.randn(40, d, generator=g), dim=1)
    sigma = 0.5 / d ** 0.5
    x = torch.nn.functional.normalize(dirs[torch.randint(0, 40, (n,), generator=g)] + sigma * torch.randn(n, d, generator=g), dim=1)
    q = torch.nn.functional.normalize(dirs[torch.randint(0, 40, (nq,), generator=g)] + sigma * torch.randn(nq, d, generator=g), dim=1)
    return x.numpy(), q.numpy()


def test_bf16_round_matches_torch():
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 3
    x[:4] = [0.0, -0.0, 1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8]            # exact ties: round to even
    np.testing.assert_array_equal(ivf.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_layout_partitions_the_rows():
    x, _ = _data()
    c = ivf.spherical_kmeans(x, 16, iters=5)
    np.testing.assert_allclose(np.linalg.norm(c, axis=1), 1.0, atol=1e-5)
    L = ivf.IVFLists(x, c)
    assert L.offsets[0] == 0 and L.offsets[-1] == x.shape[0] and np.all(np.diff(L.offsets) >= 0)
    assert sorted(L.ids.tolist()) == list(range(x.shape[0]))
    a = ivf.assign(x, c)
    for l in range(16):
        rows = L.ids[L.offsets[l]:L.offsets[l + 1]]
        assert np.all(a[rows] == l) and np.all(np.diff(rows) > 0)          # ascending original id inside a list
    # residuals are small next to the rows they encode, so their bf16 rounding error is too
    err_res = np.abs(L.reconstructed() - (x[L.ids].astype(np.float64))).mean()
    err_direct = np.abs(ivf.bf16_round(x).astype(np.float64) - x.astype(np.float64)).mean()
    assert err_res < err_direct


def test_full_probe_equals_exact_search_over_reconstructed_rows():
    x, q = _data()
    L = ivf.IVFLists(x, ivf.spherical_kmeans(x, 16, iters=5))
    ids, scores, gaps = ivf.search(L, q, nprobe=16, k=10)
    recon = L.reconstructed()
    qq = ivf.bf16_round(q)
    for i in range(q.shape[0]):
        s = recon @ qq[i].astype(np.float64)
        order = np.lexsort((L.ids, -s))[:10]
        j = 0
        while j < 10:                                                      # near ties (float64 noise) compare as sets
            e = j
            while e < 9 and abs(s[order[e]] - s[order[e + 1]]) < 1e-9:
                e += 1
            assert set(ids[i, j:e + 1].tolist()) == set(L.ids[order[j:e + 1]].tolist())
            j = e + 1
        # the coarse term enters as an fp32 value (it is what the coarse kernel hands to the fine kernel)
        np.testing.assert_allclose(scores[i], s[order], atol=1e-6)


def test_recall_grows_with_nprobe_and_small_lists_pad_with_minus_one():
    x, q = _data()
    L = ivf.IVFLists(x, ivf.spherical_kmeans(x, 32, iters=6))
    exact, _, _, _ = so.topk_exact(ivf.bf16_round(x), ivf.bf16_round(q), 10)
    rec = [ivf.recall_at_k(ivf.search(L, q, nprobe=p, k=10)[0], exact) for p in (1, 4, 32)]
    assert rec[0] <= rec[1] <= rec[2] and rec[1] > 0.9 and rec[2] > 0.95
    tiny = ivf.IVFLists(x[:20], ivf.spherical_kmeans(x[:20], 8, iters=3))
    ids, scores, gaps = ivf.search(tiny, q[:2], nprobe=1, k=10)
    n_first = [int(tiny.offsets[l + 1] - tiny.offsets[l]) for l in ivf.probe_lists(tiny, q[:2], 1)[0][:, 0]]
    for i, n in enumerate(n_first):
        assert np.all(ids[i, :n] >= 0) and np.all(ids[i, n:] == -1) and np.all(np.isneginf(scores[i, n:]))
