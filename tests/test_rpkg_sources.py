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
        # uwot's INTEGER matrix goes in as it is (the documented default of the drop-in body)
        par_i, text = h.call(2, idx, True)
        assert np.array_equal(par_i, oracle.parallel(idx))
        assert text == "Running Parallell Jaccard Coefficient Estimation...\nDone!!\n"
    # the Mann-Whitney export of the package, same harness
    from oracle.binding import WmuOracle
    from tests.test_wmu_oracle import sc_matrix

    m = sc_matrix(rng, 40, 300)
    x, y = np.asfortranarray(m[:, :60]), np.asfortranarray(m[:, 60:])
    res, text = h.wmu(x, y, True)
    assert np.array_equal(res, WmuOracle().wmu(x, y), equal_nan=True)
    assert text == "Running Parallell WM-U test...\nDone!!\n"
    bad = random_knn(rng, 100, 5)
    bad[3, 2] = 0
    with pytest.raises(RuntimeError) as e:
        h.call(0, bad)
    assert "(2)" in str(e.value)


def test_makevars_recipe_builds(tmp_path):
    """The build recipe a gficf maintainer gets (gficf_b200/rpkg/src/Makevars.cuda, INTEGRATION.md
    section 1), run as written: tools/stage_rpkg.sh lays the files out as src/ + src/cuda/ of the R
    package, then `make` follows Makevars' own nvcc rule and link line, with the few rules R's
    shlib.mk supplies (compile *.cpp with PKG_CPPFLAGS, link $(SHLIB) with PKG_LIBS) restated here
    and the stand-in R runtime on the include path (R itself is not installed)."""
    import ctypes
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = tmp_path / "gficf"
    subprocess.run([os.path.join(root, "tools", "stage_rpkg.sh"), str(pkg)], check=True, capture_output=True)
    src = pkg / "src"
    staged = sorted(os.listdir(src / "cuda"))
    assert "gficf_cuda.cu" in staged and "snn_kernels.cuh" in staged and "gficf_cuda.h" in staged
    (src / "rshlib.mk").write_text(
        "include Makevars\n"
        "OBJS = rcpp_parallel_jaccard_coeff.o jaccard_coeff.o gficf_cuda_devices.o rcpp_parallel_mann_whitney.o\n"
        "%.o: %.cpp\n\tg++ -std=c++11 -fPIC -O2 -I$(RSHIM) $(PKG_CPPFLAGS) -c $< -o $@\n"
        "$(SHLIB): $(OBJS)\n\tg++ -shared -o $@ $(OBJS) $(PKG_LIBS)\n")
    r = subprocess.run(["make", "-f", "rshlib.mk", "SHLIB=gficf.so", "RSHIM=" + os.path.join(root, "oracle", "rshim"),
                        "gficf.so"], cwd=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "-gencode arch=compute_100a,code=sm_100a" in r.stdout
    so = ctypes.CDLL(str(src / "gficf.so"))
    assert so.gficf_cuda_device_count() >= 0          # the C ABI is inside the package's shared object
    assert b"sm_100a" in ctypes.cast(so.gficf_cuda_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()
    sym = subprocess.run(["nm", "-D", "--defined-only", str(src / "gficf.so")], capture_output=True, text=True).stdout
    assert "rcpp_parallel_jaccard_coef" in sym and "jaccard_coeff" in sym and "rcpp_parallel_WMU_test" in sym
