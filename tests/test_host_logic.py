"""CPU: host-side logic of the Python mirror that needs no device."""
import numpy as np
import pytest

import gficf_b200
from gficf_b200 import api, sharding


def test_matrix_coercion_follows_rcpp():
    # numeric matrices become column-major float64 (what input_parameter<NumericMatrix> does) ...
    a = api._as_numeric_matrix(np.arange(1, 7, dtype=np.int64).reshape(3, 2))
    assert a.dtype == np.float64 and a.flags.f_contiguous and a.shape == (3, 2)
    b = api._as_numeric_matrix(np.ones((3, 2), dtype=np.float32))
    assert b.dtype == np.float64
    # ... an already column-major float64 matrix is passed without a copy ...
    c = np.asfortranarray(np.ones((4, 3)))
    assert api._as_numeric_matrix(c) is c
    # ... and R's integer matrices (int32) keep their type for the integer entry point
    d = api._as_numeric_matrix(np.ones((4, 3), dtype=np.int32))
    assert d.dtype == np.int32 and d.flags.f_contiguous
    for bad in (np.ones(3), np.ones((2, 2, 2)), np.array([["a"]])):
        with pytest.raises(TypeError):
            api._as_numeric_matrix(bad)


def test_out_argument_is_validated_before_any_device_work():
    idx = np.asfortranarray(np.ones((4, 2)))
    for out in (np.empty((8, 3)), np.empty((7, 3), order="F"), np.empty((8, 3), dtype=np.float32, order="F")):
        with pytest.raises(ValueError):
            api._call(idx, api.MODE_PARALLEL, 1, out)


def test_pinned_allocation_fails_loudly_without_a_device():
    if gficf_b200.lib().gficf_cuda_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(gficf_b200.GficfCudaError):
        gficf_b200.pinned_empty((10, 3))


def test_error_codes_are_named():
    e = gficf_b200.GficfCudaError(2, "bad id")
    assert "GFICF_E_RANGE" in str(e) and e.code == 2 and e.message == "bad id"


def test_share_bounds_cover_every_row_once_for_any_host_rank():
    for n in (7, 1000, 4_000_000):
        for world in (2, 3, 4, 8):
            for share in (0.0, 0.075, 0.39, 1.0):
                for host in (0, 1, world - 1):
                    b = sharding.share_bounds(n, world, share, host)
                    assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
                    assert all(x[1] == y[0] for x, y in zip(b, b[1:])) and all(hi >= lo for lo, hi in b)
                    assert b[host][1] - b[host][0] == int(round(share * n))
                    rest = [hi - lo for r, (lo, hi) in enumerate(b) if r != host]
                    assert max(rest) - min(rest) <= world  # the peers' parts are even (up to the remainder)
                    # the streaming peer gather's ranges start on multiples of 16 rows
                    a16 = sharding.share_bounds(n, world, share, host, align=16)
                    assert a16[0][0] == 0 and a16[-1][1] == n and all(x[1] == y[0] for x, y in zip(a16, a16[1:]))
                    assert all(lo % 16 == 0 for r, (lo, hi) in enumerate(a16) if r != host and hi > lo)  # peers' ranges


def test_weighted_bounds_with_the_host_on_the_last_rank():
    # r01 defect: the last non-host rank was stretched to n before the host's range was appended
    assert sharding.weighted_bounds(1000, 4, 0.1, 3) == [(0, 275), (275, 550), (550, 825), (825, 1000)]
    for host in range(4):
        b = sharding.weighted_bounds(1001, 4, 0.1, host)
        assert b[0][0] == 0 and b[-1][1] == 1001 and all(x[1] == y[0] for x, y in zip(b, b[1:]))


def test_balanced_host_share_equalises_host_and_peer_time():
    tf, tc, te = 3.04, 2.6, 0.62  # ms per 4M rows: fused, count, expand (B200, k=30)
    for world in (2, 4):
        x = sharding.balanced_host_share(world, tf, tc, te)
        host = x * tf + (1 - x) * te
        peer = (1 - x) * tc / (world - 1)
        assert 0 < x < 1 and abs(host - peer) < 1e-9
    # from ~6 ranks on the host rank only expands
    assert sharding.balanced_host_share(8, tf, tc, te) == 0.0
    assert sharding.balanced_host_share(1, tf, tc, te) == 1.0


def test_parity_tag_protocol_model():
    """Model of the streaming peer gather's data-as-flag protocol (expand_stream_kernel): a consumer
    that only accepts a byte whose bit 7 equals the epoch's parity never reads a stale count, whatever
    order the producers' byte stores land in, over several epochs on the same buffer."""
    rng = np.random.default_rng(5)
    e, k = 4096, 100
    buf = np.zeros(e, dtype=np.uint8)            # zero-filled: parity 0, the first epoch uses 0x80
    for epoch in range(1, 6):
        tag = (epoch & 1) << 7
        truth = rng.integers(0, k + 1, e).astype(np.uint8)
        order = rng.permutation(e)
        done = np.zeros(e, dtype=bool)
        for piece in np.array_split(order, 7):   # stores land in arbitrary pieces
            ready = (buf & 0x80) == tag          # what the consumer would accept right now
            assert not (ready & ~done).any()     # never a byte that this epoch has not written yet
            buf[piece] = truth[piece] | tag
            done[piece] = True
        assert ((buf & 0x80) == tag).all() and np.array_equal(buf & 0x7F, truth)
