"""GPU (-m gpu): BASELINE.json's full sizes, checked through size-independent properties and
oracle row samples (the whole-matrix oracle would take minutes there)."""
import numpy as np
import pytest
import torch

from gficf_b200 import synth

pytestmark = pytest.mark.gpu


def _check_properties(out3, idx0, k, oracle, r_matrix=None, sample_rows=()):
    """out3: torch [3, E] float64 on the GPU; idx0: int32 [n,k] 0-based on the GPU."""
    n = idx0.shape[0]
    frm, to, w = out3[0], out3[1], out3[2]
    lut = torch.tensor([u / (2.0 * k - u) for u in range(k + 1)], dtype=torch.float64, device=w.device)
    # every weight is one of the k+1 LUT doubles, bit for bit
    u = torch.searchsorted(lut, w.contiguous())
    assert bool((u <= k).all()) and torch.equal(lut[u], w)
    nz = w > 0
    rows = torch.arange(n, device=w.device, dtype=torch.float64).repeat_interleave(k) + 1.0
    # fixed slots: from == i+1 and to == idx+1 where u>0, zeros elsewhere
    assert torch.equal(frm, torch.where(nz, rows, torch.zeros_like(rows)))
    tgt = idx0.reshape(-1).to(torch.float64) + 1.0
    assert torch.equal(to, torch.where(nz, tgt, torch.zeros_like(tgt)))
    # symmetry: if j in N(i) and i in N(j) then u(i,j) == u(j,i)
    uu = u.reshape(n, k)
    probe = torch.randint(0, n, (200_000,), device=w.device)
    jj = torch.randint(0, k, (200_000,), device=w.device)
    t = idx0[probe, jj].long()
    back = (idx0[t] == probe[:, None].to(idx0.dtype))
    has = back.any(dim=1)
    pos = back.float().argmax(dim=1)
    assert torch.equal(uu[probe[has], jj[has]], uu[t[has], pos[has]])
    # brute-force recount of a random edge sample on the GPU with torch set ops
    s = torch.randint(0, n, (50_000,), device=w.device)
    sj = torch.randint(0, k, (50_000,), device=w.device)
    a = idx0[s]
    b = idx0[idx0[s, sj].long()]
    cnt = (a[:, :, None] == b[:, None, :]).any(dim=2).sum(dim=1)
    assert torch.equal(cnt, uu[s, sj])
    # and the oracle itself on a few row ranges
    if r_matrix is not None:
        o = out3.cpu().numpy().T if out3.shape[1] < 5e7 else None
        for lo, hi in sample_rows:
            want = oracle.parallel_rows(r_matrix, lo, hi)
            got = o[lo * k:hi * k] if o is not None else out3[:, lo * k:hi * k].cpu().numpy().T
            assert np.array_equal(got, want)


@pytest.mark.parametrize("n,k,scramble", [(1_000_000, 30, True), (4_000_000, 30, True), (4_000_000, 30, False)])
def test_k30_full_sizes_device_path(cuda, oracle, n, k, scramble):
    """configs[2] (1M) and configs[3] (4M), resident-data entry."""
    from gficf_b200 import device as D

    idx0 = synth.knn_index(n, k, scramble=scramble, device="cuda")
    padded, flags = D.pad_rows(idx0)
    out, flags = D.jaccard_edges(padded, n, k, flags=flags)
    torch.cuda.synchronize()
    assert int(flags[0]) == 0
    # the oracle itself on head / middle / tail rows, at 4M too (the rows gather from the whole matrix)
    r = synth.to_r_matrix(idx0)
    m = 300 if n <= 1_000_000 else 700
    rows = [(0, m), (n // 2, n // 2 + m), (n - m, n)]
    _check_properties(out, idx0, k, oracle, r, rows)


def test_config3_1m_host_abi_roundtrip(cuda, oracle):
    """configs[2] through the host-buffer ABI (H2D, layout pre-pass, kernel, D2H)."""
    n, k = 1_000_000, 30
    idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
    r = synth.to_r_matrix(idx0)
    out = cuda.rcpp_parallel_jaccard_coef(r)
    for lo, hi in [(0, 500), (333_333, 333_833), (n - 500, n)]:
        assert np.array_equal(out[lo * k:hi * k], oracle.parallel_rows(r, lo, hi))
    # the device path gives the same matrix
    from gficf_b200 import device as D

    padded, _ = D.pad_rows(idx0)
    dev, _ = D.jaccard_edges(padded, n, k)
    assert np.array_equal(dev.cpu().numpy().T, out)


def test_k100_wide_kernel_large(cuda, oracle):
    """configs[4]'s shape (k=100), at a size one GPU test can afford: 500k cells."""
    from gficf_b200 import device as D

    n, k = 500_000, 100
    idx0 = synth.knn_index(n, k, scramble=True, device="cuda")
    padded, flags = D.pad_rows(idx0)
    out, flags = D.jaccard_edges(padded, n, k, flags=flags)
    torch.cuda.synchronize()
    assert int(flags[0]) == 0
    r = synth.to_r_matrix(idx0[:, :])
    _check_properties(out, idx0, k, oracle, r, [(0, 40), (n - 40, n)])


def test_config5_10m_k100_device_path(cuda, oracle):
    """configs[4] at full size on ONE GPU (10M cells x k=100, E = 1e9): 4.2 GB index, 24 GB of
    edges.  Checked by oracle row samples, the LUT property and a GPU recount of random edges."""
    from gficf_b200 import device as D

    free, _ = torch.cuda.mem_get_info()
    if free < 70 << 30:
        pytest.skip("needs ~70 GB of free HBM")
    n, k = 10_000_000, 100
    idx0 = synth.knn_index(n, k, scramble=True, device="cuda", chunk=1 << 19)
    padded, flags = D.pad_rows(idx0)
    out, flags = D.jaccard_edges(padded, n, k, flags=flags)
    torch.cuda.synchronize()
    assert int(flags[0]) == 0
    w = out[2]
    lut = torch.tensor([u / (2.0 * k - u) for u in range(k + 1)], dtype=torch.float64, device="cuda")
    # chunked LUT check (1e9 doubles): every weight is one of the k+1 legal doubles
    for lo in range(0, n * k, 1 << 27):
        ww = w[lo:lo + (1 << 27)]
        u = torch.searchsorted(lut, ww.contiguous())
        assert bool((u <= k).all()) and torch.equal(lut[u], ww)
    # GPU recount of random edges with torch set ops
    s = torch.randint(0, n, (20_000,), device="cuda")
    sj = torch.randint(0, k, (20_000,), device="cuda")
    a = idx0[s]
    b = idx0[idx0[s, sj].long()]
    cnt = (a[:, :, None] == b[:, None, :]).any(dim=2).sum(dim=1)
    got_w = w[s * k + sj]
    assert torch.equal(lut[cnt], got_w)
    # exact recount of a few edges on the host with Python sets (the reference's semantics for
    # distinct ids), first and last rows
    for base in (0, n - 64):
        for i in range(base, base + 64, 16):
            ai = set(idx0[i].tolist())
            for j in (0, 37, 99):
                t = int(idx0[i, j])
                u_ref = len(ai & set(idx0[t].tolist()))
                assert float(w[i * k + j]) == u_ref / (2.0 * k - u_ref)
                assert float(out[0, i * k + j]) == (i + 1 if u_ref else 0)
                assert float(out[1, i * k + j]) == (t + 1 if u_ref else 0)
