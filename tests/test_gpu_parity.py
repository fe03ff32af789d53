"""GPU (-m gpu): the CUDA path, called through the C ABI (gficf_b200 -> libgficf_cuda.so),
against the oracle on the same inputs -- bit-exact (np.array_equal on the float64 matrices:
identical intersection counts, identical IEEE doubles, same row order)."""
import glob
import os

import numpy as np
import pytest
import torch

from gficf_b200 import synth
from tests.conftest import random_knn

pytestmark = pytest.mark.gpu

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("wmu_", "net_")))  # other rows' fixtures: test_wmu_oracle.py, test_network_oracle.py


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_vectors(cuda, path):
    """Outputs of the reference itself (tests/golden/make_golden.py)."""
    g = np.load(path)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(g["idx"]), g["parallel"])
    assert np.array_equal(cuda.jaccard_coeff(g["idx"]), g["serial"])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 29, 30, 31, 32])
def test_small_k_kernel_all_widths(cuda, oracle, k):
    rng = np.random.default_rng(k)
    idx = random_knn(rng, 1537, k)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx), oracle.parallel(idx))
    assert np.array_equal(cuda.jaccard_coeff(idx), oracle.serial(idx))


@pytest.mark.parametrize("k", [33, 40, 48, 49, 63, 64, 65, 99, 100, 127, 128])
def test_wide_k_kernel(cuda, oracle, k):
    rng = np.random.default_rng(k)
    idx = random_knn(rng, 700, k)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx), oracle.parallel(idx))


@pytest.mark.parametrize("k", [129, 200, 255, 256, 300, 513, 1000, 1024])
def test_large_k_kernel(cuda, oracle, k):
    """128 < k <= 1024: CTA per row, open-addressing table; uint16 counts above 255."""
    rng = np.random.default_rng(k)
    idx = random_knn(rng, k + 75, k)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx), oracle.parallel(idx))
    if k in (129, 256, 300):
        assert np.array_equal(cuda.jaccard_coeff(idx), oracle.serial(idx))


def test_large_k_device_counts(cuda, oracle):
    from gficf_b200 import device as D

    n, k = 3000, 400
    idx0 = synth.knn_index(n, k, family="uniform")
    padded, flags = D.pad_rows(idx0.cuda())
    assert padded.shape == (n, 400)
    cnt, flags = D.jaccard_counts(padded, n, k, flags=flags)
    assert cnt.dtype == torch.int16 and int(flags[0]) == 0
    out, _ = D.expand(padded, k, cnt, mode=0)
    assert np.array_equal(out.cpu().numpy().T, oracle.parallel(synth.to_r_matrix(idx0)))
    fused, _ = D.jaccard_edges(padded, n, k)
    assert torch.equal(fused, out)


def test_exact_kernel_beyond_fast_range_is_reachable(cuda, oracle):
    """k > 1024 has no fast kernel: the device entry says so, the host entry uses the exact path."""
    from gficf_b200 import device as D

    n, k = 1100, 1030
    idx0 = synth.knn_index(n, k, family="uniform")
    padded, flags = D.pad_rows(idx0.cuda())
    with pytest.raises(cuda.GficfCudaError) as e:
        D.jaccard_counts(padded, n, k)
    assert e.value.code == 5
    # the host ABI (what R calls) at k = 1030: exact kernel + expand, both exports
    r = synth.to_r_matrix(idx0)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(r), oracle.parallel(r))
    assert np.array_equal(cuda.jaccard_coeff(r), oracle.serial(r))


@pytest.mark.parametrize("mode", ["dma", "host", "hybrid", None])
def test_output_modes_of_the_host_abi(cuda, oracle, mode, monkeypatch):
    """The three ways the (n*k) x 3 doubles reach the caller's matrix (include/gficf_cuda.h,
    gficf_cuda_last_output) give the reference's bytes: device-written columns moved by the copy
    engine, host threads writing the columns from the 1-byte device counts, and both at once --
    on page-locked and on pageable buffers, for f64 and int32 input, incl. u == 0 rows, a repeated
    id (exact path) and sizes that leave partial pieces."""
    if mode is None:
        monkeypatch.delenv("GFICF_CUDA_OUT_MODE", raising=False)
    else:
        monkeypatch.setenv("GFICF_CUDA_OUT_MODE", mode)
    rng = np.random.default_rng(12)
    cases = [(synth.to_r_matrix(synth.knn_index(200_003, 30, scramble=True)), "planted"),
             (synth.to_r_matrix(synth.knn_index(60_001, 15, family="uniform")), "uniform: u == 0 almost everywhere"),
             (random_knn(rng, 3001, 100), "k=100"), (random_knn(rng, 700, 200), "k=200"),
             (random_knn(rng, 37, 3, with_self=True), "tiny")]
    for r, what in cases:
        n, k = r.shape
        want = oracle.parallel(r)
        for pinned in (False, True):
            for as_int in (False, True):
                src = r.astype(np.int32) if as_int else r
                if pinned:
                    buf = cuda.pinned_empty((n, k), dtype=src.dtype)
                    buf[...] = src
                    src = buf
                    out = cuda.pinned_empty((n * k, 3))
                else:
                    src = np.asfortranarray(src)
                    out = np.empty((n * k, 3), dtype=np.float64, order="F")
                out[...] = -7.0
                got = cuda.rcpp_parallel_jaccard_coef(src, False, 1, out=out)
                assert np.array_equal(np.asarray(got), want), (what, mode, pinned, as_int)
                om = cuda.last_output()
                if mode == "dma":
                    assert om["mode"] == "dma" and om["d2h_bytes"] == 24 * n * k
                elif mode == "host" or (mode is None and not pinned):
                    assert om["mode"] == "host" and om["host_share"] == 1.0 and om["d2h_bytes"] == n * k
                elif pinned:
                    assert om["mode"] == "hybrid" and n * k <= om["d2h_bytes"] <= 25 * n * k
    # a repeated id: the fast counts are discarded, the exact path rewrites everything
    r2 = random_knn(rng, 5000, 30)
    r2[77, 3] = r2[77, 9]
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(r2), oracle.parallel(r2))


@pytest.mark.parametrize("n,k", [(500, 8), (300, 30), (200, 64), (150, 100), (260, 200)])
def test_repeated_ids_take_the_exact_path(cuda, oracle, n, k):
    """Rows that list an id twice: multiset semantics in the parallel export
    (std::set_intersection), unique-set semantics in the serial one (Rcpp::intersect)."""
    rng = np.random.default_rng(n + k)
    idx = random_knn(rng, n, k, distinct=False)
    par, ser = oracle.parallel(idx), oracle.serial(idx)
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx), par)
    assert np.array_equal(cuda.jaccard_coeff(idx), ser)
    # one single repeated id in an otherwise clean matrix must be noticed too
    idx2 = random_knn(rng, n, k)
    idx2[n // 2, 0] = idx2[n // 2, k - 1]
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx2), oracle.parallel(idx2))
    assert np.array_equal(cuda.jaccard_coeff(idx2), oracle.serial(idx2))


def test_self_in_own_list_and_identical_lists(cuda, oracle):
    rng = np.random.default_rng(1)
    idx = random_knn(rng, 400, 30, with_self=True)
    idx[:, 0] = np.arange(1, 401)  # t == i  ->  u == k  ->  w == 1.0
    out = cuda.rcpp_parallel_jaccard_coef(idx)
    assert np.array_equal(out, oracle.parallel(idx))
    assert (out[0::30, 2] == 1.0).all()


def test_integer_matrix_takes_the_int32_entry(cuda, oracle):
    """uwot returns an INTEGER matrix; it is consumed as int32 (gficf_cuda_jaccard_i32), other
    integer widths are coerced to double like Rcpp does."""
    rng = np.random.default_rng(2)
    for n, k in ((300, 15), (50_000, 30), (2_000, 100)):
        idx = random_knn(rng, n, k)
        want = oracle.parallel(idx)
        as_int = np.ascontiguousarray(idx.astype(np.int32))  # C order, int32
        assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(as_int), want)
        assert np.array_equal(cuda.jaccard_coeff(np.asfortranarray(as_int)), oracle.serial(idx))
        assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(idx.astype(np.int64)), want)
    bad = random_knn(rng, 100, 5).astype(np.int32)
    bad[7, 2] = np.iinfo(np.int32).min  # NA_integer_
    with pytest.raises(cuda.GficfCudaError) as e:
        cuda.rcpp_parallel_jaccard_coef(bad)
    assert e.value.code == 2


def test_invalid_ids_are_rejected(cuda):
    rng = np.random.default_rng(3)
    base = random_knn(rng, 200, 15)
    for bad in (0.0, 201.0, -3.0, 2.5, np.nan, np.inf):
        idx = base.copy()
        idx[17, 4] = bad
        with pytest.raises(cuda.GficfCudaError) as e:
            cuda.rcpp_parallel_jaccard_coef(idx)
        assert e.value.code == 2  # GFICF_E_RANGE
    # and the library is still usable afterwards
    assert cuda.rcpp_parallel_jaccard_coef(base).shape == (3000, 3)


def test_empty_and_tiny(cuda, oracle):
    assert cuda.rcpp_parallel_jaccard_coef(np.empty((0, 5))).shape == (0, 3)
    assert cuda.rcpp_parallel_jaccard_coef(np.empty((5, 0))).shape == (0, 3)
    one = np.array([[1.0]])  # a single cell that lists itself
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(one), oracle.parallel(one))
    two = np.array([[2.0], [1.0]])
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(two), oracle.parallel(two))


def test_banners(cuda, capsys):
    idx = np.array([[2.0, 3.0], [1.0, 3.0], [1.0, 2.0]])
    cuda.rcpp_parallel_jaccard_coef(idx, True)
    assert capsys.readouterr().out == "Running Parallell Jaccard Coefficient Estimation...\nDone!!\n"
    cuda.jaccard_coeff(idx, True)
    assert capsys.readouterr().out == "Running Jaccard Coefficient Estimation...\n"


def test_pinned_buffers_and_out_argument(cuda, oracle):
    idx0 = synth.knn_index(20000, 30, seed=5)
    r = synth.to_r_matrix(idx0)
    pin_in = cuda.pinned_empty(r.shape)
    pin_in[...] = r
    pin_out = cuda.pinned_empty((r.size, 3))
    pin_out[...] = -1.0  # must be fully overwritten
    res = cuda.rcpp_parallel_jaccard_coef(pin_in, out=pin_out)
    assert res is pin_out
    assert np.array_equal(np.asarray(res), oracle.parallel(r))


def test_config1_10k_k15(cuda, oracle):
    """BASELINE.json configs[0]: 10k cells, k=15."""
    r = synth.to_r_matrix(synth.knn_index(10_000, 15))
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(r), oracle.parallel(r))
    assert np.array_equal(cuda.jaccard_coeff(r), oracle.serial(r))


@pytest.mark.parametrize("family,scramble", [("planted", False), ("planted", True), ("uniform", False)])
def test_config2_100k_k30(cuda, oracle, family, scramble):
    """BASELINE.json configs[1]: 100k cells, k=30 (E = 3e6): whole-matrix memcmp."""
    r = synth.to_r_matrix(synth.knn_index(100_000, 30, family=family, scramble=scramble))
    assert np.array_equal(cuda.rcpp_parallel_jaccard_coef(r), oracle.parallel(r))


def test_phenograph_edges_call_site(cuda, oracle):
    """R/clustCells.R:63-66: drop the self column, weight, keep weight > 0."""
    idx0 = synth.knn_index(5000, 15)
    neigh = np.concatenate([np.arange(1, 5001)[:, None], idx0.numpy() + 1], axis=1)  # col 1 = self
    rel = cuda.phenograph_edges(neigh)
    full = oracle.parallel(synth.to_r_matrix(idx0))
    assert np.array_equal(rel, full[full[:, 2] > 0])


def test_device_entry_points(cuda, oracle):
    """Resident-data path: pad -> fused kernel; counts -> expand (both modes); exact kernel."""
    from gficf_b200 import device as D

    n, k = 30_000, 30
    idx0 = synth.knn_index(n, k, scramble=True)
    r = synth.to_r_matrix(idx0)
    want = oracle.parallel(r)
    d_dense = idx0.cuda()
    padded, flags = D.pad_rows(d_dense)
    assert padded.shape == (n, 32) and int(flags[0]) == 0
    assert torch.equal(padded[:, :k], d_dense) and bool((padded[:, k:] == -2).all())
    # the f64 column-major route gives the same padded matrix
    d_r = torch.from_numpy(np.ascontiguousarray(r.T)).cuda()  # (k, n) C-order == column-major n x k
    padded2, flags2 = D.layout_from_r_matrix(d_r, n, k)
    assert torch.equal(padded, padded2) and int(flags2[0]) == 0
    out, flags = D.jaccard_edges(padded, n, k)
    torch.cuda.synchronize()
    assert int(flags[0]) == 0
    assert np.array_equal(out.cpu().numpy().T, want)
    # a slab
    lo, hi = 1234, 20_001
    out_s, _ = D.jaccard_edges(padded, n, k, lo, hi)
    assert np.array_equal(out_s.cpu().numpy().T, want[lo * k:hi * k])
    # counts + expand, fixed and compacted
    cnt, _ = D.jaccard_counts(padded, n, k)
    exp0, _ = D.expand(padded, k, cnt, mode=0)
    assert np.array_equal(exp0.cpu().numpy().T, want)
    exp1, nw = D.expand(padded, k, cnt, mode=1)
    ser = oracle.serial(r)
    assert int(nw[0]) == int((ser[:, 2] > 0).sum())
    assert np.array_equal(exp1.cpu().numpy().T, ser)
    # exact kernel agrees with the fast one on clean input
    ex = D.jaccard_counts_exact(padded, n, k, set_semantics=False, row_lo=0, row_hi=2000)
    assert torch.equal(ex, cnt[: 2000 * k])


def test_weight_division_matches_host_for_all_u(cuda):
    """w = u/(2.0*k - u) computed on the device is the host's IEEE double for every (k,u)."""
    from gficf_b200 import device as D

    for k in (1, 3, 15, 30, 100, 255):
        n = k + 1
        idx = torch.arange(n, dtype=torch.int32)[:, None].repeat(1, k)
        idx = ((idx + torch.arange(1, k + 1, dtype=torch.int32)[None, :]) % n).cuda().contiguous()
        padded, _ = D.pad_rows(idx)
        cnt = (torch.arange(n * k, dtype=torch.int64) % (k + 1)).to(torch.uint8).cuda()
        out, _ = D.expand(padded, k, cnt, mode=0)
        w = out[2].cpu().numpy()
        u = cnt.cpu().numpy().astype(np.int64)
        assert np.array_equal(w, u / (2.0 * k - u))


@pytest.mark.parametrize("n,k", [(50_000, 30), (20_003, 15), (9_999, 7), (4_001, 32), (3_000, 100), (2_500, 31), (700, 3)])
def test_streaming_gather_kernels_on_one_gpu(cuda, n, k):
    """The two kernels of the multi-GPU peer gather, exercised on ONE device: the count kernel that
    stores parity-tagged bytes (grouped 16-byte vector stores for k <= 32, row-by-row above) into a
    buffer, and the streaming expand that polls the parity of every byte -- over several segments
    with odd boundaries, for both parities, must reproduce the fused kernel bit for bit."""
    from gficf_b200 import device as D

    idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
    padded, flags = D.pad_rows(idx0)
    want, _ = D.jaccard_edges(padded, n, k)
    buf = torch.zeros(n * k + 64, dtype=torch.uint8, device="cuda")
    cuts = sorted({0, 16 * (n // 48), 16 * (n // 24) + 5, n // 2, n})  # aligned and unaligned slab starts
    segs = [(a, b) for a, b in zip(cuts, cuts[1:]) if b > a]
    for epoch in (1, 2, 3, 4):
        tag = (epoch & 1) << 7
        for lo, hi in segs:  # epochs 3, 4: row-by-row byte stores (GFICF_TAG_ROW_STORES)
            D.jaccard_counts_tagged_to(padded, n, k, lo, hi, buf.data_ptr() + lo * k, tag | (0x100 if epoch > 2 else 0),
                                       flags)
        torch.cuda.synchronize()
        assert bool(((buf[: n * k] & 0x80) == tag).all())
        out = torch.full((3, n * k), -1.0, dtype=torch.float64, device="cuda")
        D.expand_stream(padded, k, segs, buf.data_ptr(), out, tag, flags, timeout_ms=2000)
        torch.cuda.synchronize()
        assert int(flags[0]) == 0
        assert torch.equal(out, want), (n, k, epoch)
    # a byte that never arrives: the bounded spin gives up and says so instead of hanging the GPU
    buf[7 * k + 1] ^= 0x80
    out = torch.empty((3, n * k), dtype=torch.float64, device="cuda")
    D.expand_stream(padded, k, segs, buf.data_ptr(), out, (4 & 1) << 7, flags, timeout_ms=50)
    torch.cuda.synchronize()
    assert int(flags[0]) & 8
