"""Scan / finalize timing vs shard size (fixed-overhead fit for the multi-GPU strong-scaling regime)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from comorag_b200 import _native
from comorag_b200.index import DenseIndex
from util_search import make_unit_rows


def main():
    lib = _native.load()
    dev = torch.device("cuda:0")
    nq, k, dim = int(os.environ.get("NQ", "32")), int(os.environ.get("K", "10")), 1024
    q = make_unit_rows(nq, dim, 5, device=dev)
    big = make_unit_rows(10_000_000, dim, 6, device=dev)
    st = torch.cuda.current_stream()
    ws_bytes = lib.crag_search_workspace_bytes(nq, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ids = torch.empty(nq, k, dtype=torch.int64, device=dev); sc = torch.empty(nq, k, device=dev); mm = torch.empty(nq, 2, device=dev)
    out = []
    sizes = [int(x) for x in os.environ["ROWS"].split(",")] if "ROWS" in os.environ else [78_125, 156_250, 312_500, 625_000, 1_250_000, 2_500_000, 5_000_000, 10_000_000]
    for rows in sizes:
        corpus = big[:rows]

        def scan():
            _native.check(lib.crag_search_scan(corpus.data_ptr(), rows, dim, dim, q.data_ptr(), nq, k, ws.data_ptr(), ws_bytes, st.cuda_stream), "scan")

        def fin():
            _native.check(lib.crag_search_finalize(ws.data_ptr(), ws_bytes, rows, nq, k, 0, ids.data_ptr(), sc.data_ptr(), mm.data_ptr(), st.cuda_stream), "fin")

        def timeit(fns, reps=20):
            for _ in range(3):
                for f in fns: f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for f in fns: f()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3
        oi = torch.empty(nq, k, dtype=torch.int64, device=dev); osc = torch.empty(nq, k, device=dev); omm = torch.empty(nq, 2, device=dev)

        def full():
            _native.check(lib.crag_search_topk(corpus.data_ptr(), rows, dim, dim, 0, q.data_ptr(), nq, k, oi.data_ptr(), osc.data_ptr(),
                                               omm.data_ptr(), ws.data_ptr(), ws_bytes, st.cuda_stream), "topk")
        sess = DenseIndex.from_tensor(corpus).session(nq, k)
        sess.queries.copy_(q)

        def graph():
            sess.run(sess.queries)
        t_scan, t_both, t_full, t_graph = timeit([scan]), timeit([scan, fin]), timeit([full]), timeit([graph])
        ideal = rows * dim * 2 / 7.15e12 * 1e6
        out.append({"rows": rows, "k": k, "nq": nq, "scan_us": round(t_scan, 1), "scan+finalize_us": round(t_both, 1), "topk_call_us": round(t_full, 1),
                    "graph_step_us": round(t_graph, 1), "ideal_us@7.15TB/s": round(ideal, 1), "overhead_us": round(t_graph - ideal, 1),
                    "frac_of_ideal": round(ideal / t_graph, 3)})
        print(json.dumps(out[-1]), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"time_search_k{k}_nq{nq}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
