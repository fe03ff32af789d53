"""The drop-in R-package sources (gficf_b200/rpkg/src): they compile against the stand-in R
runtime and link with the product library (CPU check); on the GPU they return the reference's
matrices and print the reference's banners."""
import numpy as np
import pytest

from tests.conftest import random_knn
from tests.rpkg_harness import RPkg, build


def test_rpkg_sources_compile_and_link():
    build()
    h = RPkg()
    assert h.lib.rpkg_visible_devices() >= 0
    assert h.lib.rpkg_devices(0) >= 1  # query only


def test_rpkg_error_becomes_r_error_without_gpu(cuda_absent=None):
    import gficf_b200

    if gficf_b200.lib().gficf_cuda_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError) as e:
        RPkg().call(0, np.array([[2.0, 3.0], [1.0, 3.0], [1.0, 2.0]]))
    assert "gficf CUDA Jaccard failed (3)" in str(e.value)


@pytest.mark.gpu
def test_rpkg_bodies_match_reference(cuda, oracle):
    h = RPkg()
    rng = np.random.default_rng(9)
    for n, k in ((2000, 15), (1500, 30), (300, 100)):
        idx = random_knn(rng, n, k)
        par, text = h.call(0, idx, True)
        assert np.array_equal(par, oracle.parallel(idx))
        assert text == "Running Parallell Jaccard Coefficient Estimation...\nDone!!\n"
        ser, text = h.call(1, idx, True)
        assert np.array_equal(ser, oracle.serial(idx))
        assert text == "Running Jaccard Coefficient Estimation...\n"
    bad = random_knn(rng, 100, 5)
    bad[3, 2] = 0
    with pytest.raises(RuntimeError) as e:
        h.call(0, bad)
    assert "(2)" in str(e.value)
