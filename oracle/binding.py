"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

``Oracle``    -> oracle/libgficf_oracle.so  (oracle/jaccard_oracle.c, our restatement of
                 /root/reference/src/rcpp_parallel_jaccard_coeff.cpp:24-55 and
                 /root/reference/src/jaccard_coeff.cpp:28-42)
``Reference`` -> oracle/_ref/libgficf_ref.so (those reference files themselves, built by
                 oracle/Makefile; absent if the container never built them)

All matrices use the reference's conventions: ``idx`` is n x k float64,
Fortran (column-major) order, 1-based; results are (n*k) x 3 float64 Fortran order.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libgficf_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgficf_ref.so")
REF_WMU_SO = os.path.join(_HERE, "_ref", "libgficf_ref_wmu.so")
MODOPT_BIN = os.path.join(_HERE, "_ref", "modopt")
REF_MODOPT_SO = os.path.join(_HERE, "_ref", "libgficf_ref_modopt.so")

_dp = C.POINTER(C.c_double)


def build(verbose: bool = False) -> None:
    """Compile the checkers (the C restatement always; oracle/_ref when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _as_idx(idx) -> np.ndarray:
    a = np.asarray(idx)
    if a.ndim != 2:
        raise ValueError("idx must be a 2-D matrix")
    return np.asfortranarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class Oracle:
    """Our plain-C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.gficf_oracle_parallel_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int32]
        self.lib.gficf_oracle_parallel_jaccard.restype = None
        self.lib.gficf_oracle_parallel_jaccard_rows.argtypes = [
            _dp, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _dp, C.c_int32]
        self.lib.gficf_oracle_parallel_jaccard_rows.restype = None
        self.lib.gficf_oracle_serial_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp]
        self.lib.gficf_oracle_serial_jaccard.restype = C.c_int64

    def parallel(self, idx, nthreads: int | None = None) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.lib.gficf_oracle_parallel_jaccard(_ptr(a), n, k, _ptr(out), nthreads or os.cpu_count() or 1)
        return out

    def parallel_rows(self, idx, lo: int, hi: int, nthreads: int | None = None) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.zeros(((hi - lo) * k, 3), dtype=np.float64, order="F")
        self.lib.gficf_oracle_parallel_jaccard_rows(_ptr(a), n, k, lo, hi, _ptr(out),
                                                    nthreads or os.cpu_count() or 1)
        return out

    def serial(self, idx) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.lib.gficf_oracle_serial_jaccard(_ptr(a), n, k, _ptr(out))
        return out


class Reference:
    """The reference's own Jaccard sources (oracle/_ref), R runtime stubbed."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (build it in the container: make -C oracle ref)")
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.gficf_ref_parallel_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int32, C.c_int32]
        L.gficf_ref_parallel_jaccard.restype = C.c_double
        L.gficf_ref_parallel_jaccard_rows.argtypes = [
            _dp, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _dp, C.c_int32]
        L.gficf_ref_parallel_jaccard_rows.restype = C.c_double
        L.gficf_ref_serial_jaccard.argtypes = [_dp, C.c_int64, C.c_int32, _dp, C.c_int32]
        L.gficf_ref_serial_jaccard.restype = C.c_double
        L.gficf_ref_capture_begin.argtypes = []
        L.gficf_ref_capture_begin.restype = None
        L.gficf_ref_capture_end.argtypes = [C.c_char_p, C.c_int64]
        L.gficf_ref_capture_end.restype = C.c_int64
        L.gficf_ref_hw_threads.argtypes = []
        L.gficf_ref_hw_threads.restype = C.c_int32
        self.last_seconds = 0.0

    def parallel(self, idx, print_output: bool = False, nthreads: int = 0) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_parallel_jaccard(
            _ptr(a), n, k, _ptr(out), int(print_output), nthreads)
        return out

    def parallel_rows(self, idx, lo: int, hi: int, nthreads: int = 0) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.zeros(((hi - lo) * k, 3), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_parallel_jaccard_rows(
            _ptr(a), n, k, lo, hi, _ptr(out), nthreads)
        return out

    def serial(self, idx, print_output: bool = False) -> np.ndarray:
        a = _as_idx(idx)
        n, k = a.shape
        out = np.empty((n * k, 3), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_serial_jaccard(_ptr(a), n, k, _ptr(out), int(print_output))
        return out

    def captured(self, fn, *args, **kw):
        """Run fn and return (result, text the reference printed through Rprintf)."""
        self.lib.gficf_ref_capture_begin()
        try:
            res = fn(*args, **kw)
        finally:
            buf = C.create_string_buffer(4096)
            self.lib.gficf_ref_capture_end(buf, 4096)
        return res, buf.value.decode()

    def hw_threads(self) -> int:
        return int(self.lib.gficf_ref_hw_threads())


# ---------------------------------------------------------------------------------------------
# Mann-Whitney U per gene: matX n_genes x n1, matY n_genes x n2 (Fortran order
# float64, what R hands over) -> n_genes x 2 (p-value, log2 fold change), Fortran order.
# ---------------------------------------------------------------------------------------------
def _as_f64(m) -> np.ndarray:
    a = np.asarray(m)
    if a.ndim != 2:
        raise ValueError("a matrix is required")
    return np.asfortranarray(a, dtype=np.float64)


class WmuOracle:
    """oracle/wmu_oracle.c: our plain-C restatement of rcpp_parallel_mann_whitney.cpp:27-100."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.gficf_oracle_wmu.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int32]
        self.lib.gficf_oracle_wmu.restype = None
        self.lib.gsl_cdf_ugaussian_P.argtypes = [C.c_double]
        self.lib.gsl_cdf_ugaussian_P.restype = C.c_double
        self.lib.gsl_cdf_ugaussian_Q.argtypes = [C.c_double]
        self.lib.gsl_cdf_ugaussian_Q.restype = C.c_double

    def wmu(self, mat_x, mat_y, nthreads: int | None = None) -> np.ndarray:
        x, y = _as_f64(mat_x), _as_f64(mat_y)
        g = x.shape[0]
        if y.shape[0] != g:
            raise ValueError("matX and matY must have the same number of rows (genes)")
        out = np.empty((g, 2), dtype=np.float64, order="F")
        self.lib.gficf_oracle_wmu(_ptr(x), _ptr(y), g, x.shape[1], y.shape[1], _ptr(out),
                                  nthreads or os.cpu_count() or 1)
        return out


class WmuReference:
    """The reference's own mann_whitney.cpp + rcpp_parallel_mann_whitney.cpp (oracle/_ref), R runtime
    stubbed, GSL's normal cdf restated (parity unpinned on the cdf)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_WMU_SO)

    def __init__(self):
        if not os.path.exists(REF_WMU_SO):
            raise FileNotFoundError(REF_WMU_SO + " (build it in the container: make -C oracle ref)")
        self.lib = C.CDLL(REF_WMU_SO)
        self.lib.gficf_ref_wmu.argtypes = [_dp, _dp, C.c_int64, C.c_int64, C.c_int64, _dp, C.c_int32, C.c_int32]
        self.lib.gficf_ref_wmu.restype = C.c_double
        self.last_seconds = 0.0

    def wmu(self, mat_x, mat_y, nthreads: int = 0) -> np.ndarray:
        x, y = _as_f64(mat_x), _as_f64(mat_y)
        g = x.shape[0]
        out = np.empty((g, 2), dtype=np.float64, order="F")
        self.last_seconds = self.lib.gficf_ref_wmu(_ptr(x), _ptr(y), g, x.shape[1], y.shape[1], _ptr(out), 0, nthreads)
        return out


# ---------------------------------------------------------------------------------------------
# The data-parallel pieces of the community detection: network from the
# lower-triangle edge list, quality function, reduced network.  A network is the dict
#   n_nodes, first (int32[n_nodes+1]), neighbor (int32[E]), edge_w (f64[E]), node_w (f64[n_nodes]),
#   total_w (getTotalEdgeWeight), self_links (totalEdgeWeightSelfLinks)
# ---------------------------------------------------------------------------------------------
_ip = C.POINTER(C.c_int32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _pi(a: np.ndarray):
    return a.ctypes.data_as(_ip)


class NetworkOracle:
    """oracle/modopt_oracle.c: our plain-C restatement of ModularityOptimizer.cpp:761-806, :462-482, :322-373."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.modopt_network.argtypes = [_ip, _ip, _dp, C.c_longlong, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.modopt_network.restype = C.c_longlong
        L.modopt_quality.argtypes = [C.c_int, _ip, _ip, _dp, _dp, C.c_double, C.c_double, _ip, C.c_int,
                                     C.c_double, _dp]
        L.modopt_quality.restype = C.c_double
        L.modopt_reduce.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _ip, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.modopt_reduce.restype = C.c_longlong
        L.modopt_seq_sum.argtypes = [_dp, C.c_longlong, C.c_double]
        L.modopt_seq_sum.restype = C.c_double

    def network(self, node1, node2, w) -> dict:
        a, b, ww = _i32(node1), _i32(node2), _f64(w)
        m = a.size
        n_nodes = int(max(a.max(), b.max())) + 1
        first = np.zeros(n_nodes + 1, np.int32)
        neighbor = np.zeros(2 * m, np.int32)
        edge_w = np.zeros(2 * m, np.float64)
        node_w = np.zeros(n_nodes, np.float64)
        total = C.c_double(0.0)
        e = self.lib.modopt_network(_pi(a), _pi(b), _ptr(ww), m, n_nodes, _pi(first), _pi(neighbor),
                                    _ptr(edge_w), _ptr(node_w), C.byref(total))
        return dict(n_nodes=n_nodes, first=first, neighbor=neighbor[:e].copy(), edge_w=edge_w[:e].copy(),
                    node_w=node_w, total_w=total.value, self_links=0.0)

    def quality(self, net: dict, cluster, resolution: float):
        cl = _i32(cluster)
        nc = int(cl.max()) + 1
        cw = np.zeros(nc, np.float64)
        q = self.lib.modopt_quality(net["n_nodes"], _pi(net["first"]), _pi(net["neighbor"]), _ptr(net["edge_w"]),
                                    _ptr(net["node_w"]), net["self_links"], net["total_w"], _pi(cl), nc,
                                    float(resolution), _ptr(cw))
        return q, cw

    def reduce(self, net: dict, cluster) -> dict:
        cl = _i32(cluster)
        nc = int(cl.max()) + 1
        e = max(1, net["neighbor"].size)
        first = np.zeros(nc + 1, np.int32)
        neighbor = np.zeros(e, np.int32)
        edge_w = np.zeros(e, np.float64)
        node_w = np.zeros(nc, np.float64)
        self_links = C.c_double(net["self_links"])
        n_red = self.lib.modopt_reduce(net["n_nodes"], _pi(net["first"]), _pi(net["neighbor"]), _ptr(net["edge_w"]),
                                       _ptr(net["node_w"]), _pi(cl), nc, _pi(first), _pi(neighbor), _ptr(edge_w),
                                       _ptr(node_w), C.byref(self_links))
        ew = edge_w[:n_red].copy()
        return dict(n_nodes=nc, first=first, neighbor=neighbor[:n_red].copy(), edge_w=ew, node_w=node_w,
                    total_w=self.lib.modopt_seq_sum(_ptr(ew), ew.size, 0.0) / 2.0,
                    self_links=self_links.value)


class NetworkReference:
    """The reference's own Network / Clustering / VOSClusteringTechnique classes (oracle/_ref)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_MODOPT_SO)

    def __init__(self):
        if not os.path.exists(REF_MODOPT_SO):
            raise FileNotFoundError(REF_MODOPT_SO + " (build it in the container: make -C oracle ref)")
        self.lib = C.CDLL(REF_MODOPT_SO)
        L = self.lib
        L.ref_net_build.argtypes = [_ip, _ip, _dp, C.c_longlong, C.c_int]
        L.ref_net_build.restype = C.c_void_p
        L.ref_net_dims.argtypes = [C.c_void_p, _ip, _ip]
        L.ref_net_dims.restype = None
        L.ref_net_get.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp, _dp]
        L.ref_net_get.restype = None
        L.ref_net_quality.argtypes = [C.c_void_p, _ip, C.c_double]
        L.ref_net_quality.restype = C.c_double
        L.ref_net_reduce.argtypes = [C.c_void_p, _ip]
        L.ref_net_reduce.restype = C.c_void_p
        L.ref_net_free.argtypes = [C.c_void_p]
        L.ref_net_free.restype = None
        L.ref_louvain_hooked.argtypes = [_ip, _ip, _dp, C.c_longlong, C.c_double, C.c_int, C.c_int, C.c_int,
                                         C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p, _ip, _dp,
                                         C.POINTER(C.c_longlong)]
        L.ref_louvain_hooked.restype = C.c_int

    def _export(self, h) -> dict:
        nn, ne = C.c_int32(0), C.c_int32(0)
        self.lib.ref_net_dims(h, C.byref(nn), C.byref(ne))
        first = np.zeros(nn.value + 1, np.int32)
        neighbor = np.zeros(max(1, ne.value), np.int32)
        edge_w = np.zeros(max(1, ne.value), np.float64)
        node_w = np.zeros(nn.value, np.float64)
        total, self_links = C.c_double(0.0), C.c_double(0.0)
        self.lib.ref_net_get(h, _pi(first), _pi(neighbor), _ptr(edge_w), _ptr(node_w), C.byref(total),
                             C.byref(self_links))
        return dict(n_nodes=nn.value, first=first, neighbor=neighbor[:ne.value].copy(),
                    edge_w=edge_w[:ne.value].copy(), node_w=node_w, total_w=total.value,
                    self_links=self_links.value, handle=h)

    def network(self, node1, node2, w) -> dict:
        a, b, ww = _i32(node1), _i32(node2), _f64(w)
        return self._export(self.lib.ref_net_build(_pi(a), _pi(b), _ptr(ww), a.size, 1))

    def quality(self, net: dict, cluster, resolution: float) -> float:
        return self.lib.ref_net_quality(net["handle"], _pi(_i32(cluster)), float(resolution))

    def reduce(self, net: dict, cluster) -> dict:
        return self._export(self.lib.ref_net_reduce(net["handle"], _pi(_i32(cluster))))

    def free(self, net: dict) -> None:
        self.lib.ref_net_free(net.pop("handle"))

    # -- the reference's Louvain run with the three bulk steps optionally supplied from outside --------
    _HOOK_NETWORK = C.CFUNCTYPE(None, _ip, _ip, _dp, C.c_longlong, C.c_int, _ip, _ip, _dp, _dp, _dp)
    _HOOK_QUALITY = C.CFUNCTYPE(C.c_double, C.c_int, _ip, _ip, _dp, _dp, C.c_double, _ip, C.c_int, C.c_double)
    _HOOK_REDUCE = C.CFUNCTYPE(C.c_longlong, C.c_int, _ip, _ip, _dp, _dp, C.c_double, _ip, C.c_int, _ip, _ip, _dp,
                               _dp, _dp)

    def louvain(self, node1, node2, w, resolution=0.8, algorithm=1, n_start=10, n_iter=10, seed=180582, hooks=None):
        """Labels (after orderClustersByNNodes), the maximum modularity and how often each hook ran.
        `hooks` (optional) provides network(node1, node2, w, n_nodes) -> dict, quality(net, cluster,
        n_clusters, resolution) -> float and reduce(net, cluster, n_clusters) -> dict on numpy arrays (`net`
        dicts as everywhere in this module); without it the reference's own functions run: the
        reference's algorithm end to end (RModularityOptimizer.cpp:101-172)."""
        a, b, ww = _i32(node1), _i32(node2), _f64(w)
        n_nodes = int(max(a.max(), b.max())) + 1
        as_arr = np.ctypeslib.as_array
        errors = []

        def view_net(n, first, neighbor, edge_w, node_w, self_links):
            f = as_arr(first, shape=(n + 1,))
            e = int(f[n])
            return dict(n_nodes=n, first=f, neighbor=as_arr(neighbor, shape=(max(e, 1),))[:e],
                        edge_w=as_arr(edge_w, shape=(max(e, 1),))[:e], node_w=as_arr(node_w, shape=(n,)),
                        self_links=self_links)

        def cb_network(p1, p2, pw, m, n, first, neighbor, edge_w, node_w, total_w):
            try:
                net = hooks.network(as_arr(p1, shape=(m,)), as_arr(p2, shape=(m,)), as_arr(pw, shape=(m,)), n)
                as_arr(first, shape=(n + 1,))[:] = net["first"]
                as_arr(neighbor, shape=(2 * m,))[:] = net["neighbor"]
                as_arr(edge_w, shape=(2 * m,))[:] = net["edge_w"]
                as_arr(node_w, shape=(n,))[:] = net["node_w"]
                total_w[0] = net["total_w"]
            except Exception as ex:  # exceptions cannot cross the C frames
                errors.append(ex)

        def cb_quality(n, first, neighbor, edge_w, node_w, self_links, cluster, nc, res):
            try:
                return float(hooks.quality(view_net(n, first, neighbor, edge_w, node_w, self_links),
                                           as_arr(cluster, shape=(n,)), nc, res))
            except Exception as ex:
                errors.append(ex)
                return 0.0

        def cb_reduce(n, first, neighbor, edge_w, node_w, self_links, cluster, nc, r_first, r_neighbor, r_edge_w,
                      r_node_w, r_self_links):
            try:
                red = hooks.reduce(view_net(n, first, neighbor, edge_w, node_w, self_links),
                                   as_arr(cluster, shape=(n,)), nc)
                e = red["neighbor"].size
                as_arr(r_first, shape=(nc + 1,))[:] = red["first"]
                if e:
                    as_arr(r_neighbor, shape=(e,))[:] = red["neighbor"]
                    as_arr(r_edge_w, shape=(e,))[:] = red["edge_w"]
                as_arr(r_node_w, shape=(nc,))[:] = red["node_w"]
                r_self_links[0] = red["self_links"]
                return e
            except Exception as ex:
                errors.append(ex)
                return 0

        keep = (self._HOOK_NETWORK(cb_network), self._HOOK_QUALITY(cb_quality), self._HOOK_REDUCE(cb_reduce))
        ptrs = [C.cast(k, C.c_void_p) if hooks is not None else None for k in keep]
        labels = np.zeros(n_nodes, np.int32)
        max_mod = C.c_double(0.0)
        calls = (C.c_longlong * 3)()
        rc = self.lib.ref_louvain_hooked(_pi(a), _pi(b), _ptr(ww), a.size, float(resolution), int(algorithm),
                                         int(n_start), int(n_iter), int(seed), ptrs[0], ptrs[1], ptrs[2],
                                         _pi(labels), C.byref(max_mod), calls)
        if errors:
            raise errors[0]
        if rc < 0:
            raise RuntimeError("the reference's optimiser threw")
        return labels, max_mod.value, list(calls)
