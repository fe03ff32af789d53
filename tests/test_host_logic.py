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


def test_chunk_major_plan_covers_every_row_once():
    for n in (7, 1000, 4_000_000):
        for world in (2, 4, 8):
            for rho in (0.0, 0.29, 1.0):
                plan = sharding.chunk_major_bounds(n, world, rho, 4)
                flat = [r for chunk in plan for r in chunk]
                assert flat[0][0] == 0 and flat[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
                # a chunk is one contiguous range for the host rank's expand kernel
                for chunk in plan:
                    assert chunk[0][0] <= chunk[-1][1] and all(lo <= hi for lo, hi in chunk)
