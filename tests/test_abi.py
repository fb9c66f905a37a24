"""CPU-side checks of the boundary: the built library exports every symbol include/smm_b200.h declares, and compute
calls fail loudly (no silent CPU fallback) when no CUDA device is present."""
import ctypes as C

import pytest


def test_library_exports_every_declared_symbol():
    from sparse_matrix_math_b200 import build
    build.build()
    import sparse_matrix_math_b200 as smm
    L = smm.lib()
    assert len(smm.ABI_SYMBOLS) >= 35
    missing = [s for s in smm.ABI_SYMBOLS if not hasattr(L, s)]
    assert not missing, missing
    assert L.smm_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import numpy as np

    import sparse_matrix_math_b200 as smm
    n = C.c_int(0)
    rc = smm.lib().smm_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(smm.SmmError):
        smm.CSRMatrix.from_arrays(1, 1, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([1], np.float32))
    with pytest.raises(smm.SmmError):
        smm.dot(np.ones(4, np.float32), np.ones(4, np.float32))


def test_product_does_not_reference_the_oracle():
    """The product (package, headers, kernels) must not import, link or call anything under oracle/."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for d in ("sparse_matrix_math_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(root, d)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    text = open(os.path.join(dirpath, f), errors="replace").read()
                    if "smm_oracle" in text or "oracle_lib" in text or "smm_ref_" in text:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_python_triplet_matches_oracle_conversion():
    import numpy as np

    import oracle_lib as ol
    import sparse_matrix_math_b200 as smm
    rng = np.random.default_rng(0)
    rows, cols, n = 50, 40, 700
    tr, tc = rng.integers(0, rows, n), rng.integers(0, cols, n)
    tv = rng.uniform(-1, 1, n).astype(np.float32)
    t = smm.TripletMatrix(rows, cols)
    for r, c, v in zip(tr, tc, tv):
        t.addEntry(r, c, v)
    start, pos, val = t.to_csr_arrays()
    o = ol.triplets_to_csr(rows, cols, tr, tc, tv)
    assert np.array_equal(start, o.start) and np.array_equal(pos, o.positions) and val.tobytes() == o.values.tobytes()


def test_python_loader_matches_oracle(tmp_path):
    import os

    import numpy as np

    import oracle_lib as ol
    import sparse_matrix_math_b200 as smm
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("load_symmetric_test", "mesh1e1", "sherman1"):
        t = smm.TripletMatrix()
        assert smm.loadMatrix(os.path.join(gold, name + ".mtx"), t) == smm.MatrixLoadStatus.SUCCESS
        st, o = ol.load_mtx(os.path.join(gold, name + ".mtx"))
        start, pos, val = t.to_csr_arrays()
        assert st == 0 and np.array_equal(start, o.start) and np.array_equal(pos, o.positions) and val.tobytes() == o.values.tobytes()
    assert smm.loadMatrix(str(tmp_path / "x.bin"), smm.TripletMatrix()) == smm.MatrixLoadStatus.FAILED_TO_OPEN_FILE_UNKNOWN_FORMAT
    assert smm.loadMatrix(str(tmp_path / "x.mtx"), smm.TripletMatrix()) == smm.MatrixLoadStatus.FAILED_TO_OPEN_FILE
