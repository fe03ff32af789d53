import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """A fresh checkout has no binaries (they are git-ignored): compile the product library and the
    checkers once per session, exactly what __graft_entry__.build() does."""
    import gficf_b200
    from oracle import binding

    if not os.path.exists(gficf_b200.library_path()) or not os.path.exists(binding.ORACLE_SO):
        sys.path.insert(0, ROOT)
        import __graft_entry__

        __graft_entry__.build()
    yield


def has_cuda() -> bool:
    try:
        import gficf_b200

        return gficf_b200.lib().gficf_cuda_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.binding import Reference

    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; run make -C oracle ref)")
    return Reference()


@pytest.fixture(scope="session")
def cuda():
    """The product library, with a visible GPU.  GPU tests FAIL (not skip) without it when
    selected with -m gpu on a box that has one; on a CPU-only box they are deselected."""
    import gficf_b200

    L = gficf_b200.lib()
    if L.gficf_cuda_device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests need the B200 box")
    return gficf_b200


def random_knn(rng, n, k, distinct=True, with_self=False):
    """n x k float64 1-based ids (Fortran order), optionally with repeated ids / self ids."""
    if distinct:
        rows = [rng.choice(n - (0 if with_self else 1), k, replace=False) for _ in range(n)]
        a = np.stack(rows)
        if not with_self:
            a = a + (a >= np.arange(n)[:, None])  # skip own id
    else:
        a = rng.integers(0, n, size=(n, k))
    return np.asfortranarray(a.astype(np.float64) + 1.0)
