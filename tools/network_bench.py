"""Times the network pieces (gficf_b200.modularity) on a device-resident SNN graph and checks them
against the oracle at the same size.  One GPU.  Usage: python tools/network_bench.py [n] [k] [cluster_size]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gficf_b200 import device as D, modularity, snn, synth  # noqa: E402
from oracle.binding import NetworkOracle  # noqa: E402


def timed(fn, reps=3):
    out = fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return out, best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    csize = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    idx0 = synth.knn_index(n, k, family="planted", scramble=False)
    padded, _ = D.pad_rows(idx0.cuda())
    colptr, rows, w, flags = snn.snn_lower_triangle(padded, n, k)
    nv, nnz = colptr.numel() - 1, rows.numel()
    rec = {"n": n, "k": k, "vertices": nv, "lower_entries": nnz, "flags": int(flags[0])}
    net, rec["network_ms"] = timed(lambda: modularity.matrix_to_network(colptr, rows, w))
    cl = (torch.arange(nv, device="cuda", dtype=torch.int32) // csize).contiguous()
    nc = int(cl.max()) + 1
    res = 0.8 / (2 * net.get_total_edge_weight())
    q, rec["quality_ms"] = timed(lambda: net.calc_quality_function(cl, res, n_clusters=nc))
    red, rec["reduce_ms"] = timed(lambda: net.create_reduced_network(cl, n_clusters=nc))
    rec.update(clusters=nc, directed_edges=net.n_edges, reduced_edges=red.n_edges, quality=q)
    # the oracle (single-threaded C restatement of the reference's loops) on the same input
    O = NetworkOracle()
    cp = colptr.cpu().numpy()
    cols = np.repeat(np.arange(nv, dtype=np.int32), np.diff(cp))
    r_h, w_h, cl_h = rows.cpu().numpy(), w.cpu().numpy(), cl.cpu().numpy()
    t = time.perf_counter()
    want = O.network(cols, r_h, w_h)
    rec["cpu_network_ms"] = (time.perf_counter() - t) * 1e3
    t = time.perf_counter()
    q_want, cw_want = O.quality(want, cl_h, res)
    rec["cpu_quality_ms"] = (time.perf_counter() - t) * 1e3
    t = time.perf_counter()
    red_want = O.reduce(want, cl_h)
    rec["cpu_reduce_ms"] = (time.perf_counter() - t) * 1e3
    rec["parity"] = {
        "network_arrays_equal": bool(np.array_equal(net.first_neighbor_index.cpu().numpy(), want["first"]) and
                                     np.array_equal(net.neighbor.cpu().numpy(), want["neighbor"]) and
                                     np.array_equal(net.edge_weight.cpu().numpy(), want["edge_w"]) and
                                     np.array_equal(net.node_weight.cpu().numpy(), want["node_w"])),
        "cluster_weights_equal": bool(np.array_equal(net.cluster_weights(cl, nc).cpu().numpy(), cw_want)),
        "reduced_arrays_equal": bool(np.array_equal(red.first_neighbor_index.cpu().numpy(), red_want["first"]) and
                                     np.array_equal(red.neighbor.cpu().numpy(), red_want["neighbor"]) and
                                     np.array_equal(red.edge_weight.cpu().numpy(), red_want["edge_w"]) and
                                     np.array_equal(red.node_weight.cpu().numpy(), red_want["node_w"])),
        "quality_equal": bool(q == q_want),
        "total_weight_equal": bool(net.get_total_edge_weight() == want["total_w"]),
        "self_links_equal": bool(red.total_edge_weight_self_links == red_want["self_links"]),
        "reduced_total_weight_equal": bool(red.get_total_edge_weight() == red_want["total_w"]),
    }
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
