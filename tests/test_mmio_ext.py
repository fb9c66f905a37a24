"""the Matrix Market extension (general / skew-symmetric structure, pattern field).  The reference's
loader rejects these files (H:2559-2574), so the reference-named loaders must keep rejecting them; the extension is
checked against scipy.io.mmread (an independent reader) in Python and in the C++ header (host-only program)."""
import os
import subprocess

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

import sparse_matrix_math_b200 as smm
from test_cpp_dropin import compile_cpp, built_lib  # noqa: F401

FILES = {
    "general_real": """%%MatrixMarket matrix coordinate real general
% a comment
4 5 6
1 1 1.5
1 5 -2.25
2 2 0.1
3 1 7
4 4 1e-3
4 5 0
""",
    "general_integer_dups": """%%MatrixMarket matrix coordinate integer general
3 3 4
1 2 3
1 2 4
3 3 -1
2 1 5
""",
    "skew": """%%MatrixMarket MATRIX Coordinate Real Skew-Symmetric
3 3 2
2 1 2.5
3 2 -0.75
""",
    "pattern_general": """%%MatrixMarket matrix coordinate pattern general
3 4 4
1 1
2 3
3 4
3 1
""",
    "pattern_symmetric": """%%MatrixMarket matrix coordinate pattern symmetric
3 3 3
1 1
3 1
2 2
""",
    "empty_general": """%%MatrixMarket matrix coordinate real general
2 2 0
""",
}


def dense_from_triplet(t):
    d = np.zeros((t.getDenseRowCount(), t.getDenseColCount()), np.float32)
    for r, c, v in t:
        d[r, c] = v
    return d


@pytest.mark.parametrize("name", sorted(FILES))
def test_extended_loader_matches_scipy(name, tmp_path):
    path = tmp_path / (name + ".mtx")
    path.write_text(FILES[name])
    t = smm.TripletMatrix()
    assert smm.loadMatrix(str(path), t, extended=True) == smm.MatrixLoadStatus.SUCCESS
    want = sp.coo_matrix(scipy.io.mmread(str(path), spmatrix=True)).astype(np.float32).toarray()   # duplicates are summed, like addEntry
    np.testing.assert_array_equal(dense_from_triplet(t), want)


@pytest.mark.parametrize("name,status", [
    ("general_real", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE"),
    ("skew", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE"),
    ("pattern_general", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE"),
    ("pattern_symmetric", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE"),
])
def test_reference_named_loader_keeps_rejecting(name, status, tmp_path):
    path = tmp_path / (name + ".mtx")
    path.write_text(FILES[name])
    assert smm.loadMatrix(str(path), smm.TripletMatrix()) == getattr(smm.MatrixLoadStatus, status)   # H:2559-2574


def test_extended_loader_unsupported(tmp_path):
    for body, status in (("%%MatrixMarket matrix array real general\n1 1\n1\n", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_FORMAT"),
                         ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE"),
                         ("%%MatrixMarket matrix coordinate real hermitian\n1 1 1\n1 1 1\n", "PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE"),
                         ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 x 1\n", "FAILED_TO_PARSE_FILE")):
        path = tmp_path / "bad.mtx"
        path.write_text(body)
        assert smm.loadMatrix(str(path), smm.TripletMatrix(), extended=True) == getattr(smm.MatrixLoadStatus, status)


def test_cpp_header_extension(built_lib, tmp_path):  # noqa: F811
    """SMM::ext::loadMatrix in the drop-in header: same entries as scipy; SMM::loadMatrix still rejects (host-only)."""
    for name, body in FILES.items():
        (tmp_path / (name + ".mtx")).write_text(body)
    src = tmp_path / "mm_ext.cpp"
    src.write_text(r'''
#include <cstdio>
#include "sparse_matrix_math.h"
int main(int, char** argv) {
    SMM::TripletMatrix<float> t;
    const SMM::MatrixLoadStatus strict = SMM::loadMatrix(argv[1], t);
    SMM::TripletMatrix<float> e;
    const SMM::MatrixLoadStatus st = SMM::ext::loadMatrix(argv[1], e);
    std::printf("%d %d %d %d\n", (int)strict, (int)st, e.getDenseRowCount(), e.getDenseColCount());
    for (const auto& el : e) std::printf("%d %d %.9g\n", el.getRow(), el.getCol(), (double)el.getValue());
    return 0;
}
''')
    exe = str(tmp_path / "mm_ext")
    r = compile_cpp(str(src), exe)
    assert r.returncode == 0, r.stderr[-3000:]
    randoms = [str(q) for q in random_mm_files(tmp_path)]
    for path in [str(tmp_path / (name + ".mtx")) for name in FILES] + randoms:
        out = subprocess.run([exe, path], capture_output=True, text=True).stdout.split("\n")
        strict, st, rows, cols = (int(v) for v in out[0].split())
        assert st == 0, path
        if path not in randoms:
            assert strict != 0                               # every hand-written file is one the reference's loader rejects
        got = np.zeros((rows, cols), np.float32)
        for line in out[1:]:
            if line.strip():
                r_, c_, v_ = line.split()
                got[int(r_), int(c_)] = np.float32(v_)
        want = sp.coo_matrix(scipy.io.mmread(path, spmatrix=True)).astype(np.float32).toarray()
        np.testing.assert_array_equal(got, want)


def random_mm_files(tmp_path, count=40):
    """Seeded random coordinate files of every supported kind (no duplicate entries: those are summed in float32, in file
    order, H:606-618, while scipy sums in double)."""
    rng = np.random.default_rng(20261018)
    paths = []
    for trial in range(count):
        rows, cols = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        structure = ["general", "symmetric", "skew-symmetric"][trial % 3]
        field = ["real", "integer", "pattern"][(trial // 3) % 3]
        if structure != "general":
            cols = rows
        if structure == "skew-symmetric" and field == "pattern":
            field = "real"                                    # not a Matrix Market combination
        n = int(rng.integers(0, 20))
        lines = []
        seen = set()
        for _ in range(n):
            r, c = int(rng.integers(1, rows + 1)), int(rng.integers(1, cols + 1))
            if structure != "general" and c > r:
                r, c = c, r                                   # lower triangle only
            if structure == "skew-symmetric" and r == c:
                continue                                      # no diagonal in a skew-symmetric file
            if (r, c) in seen:
                continue
            seen.add((r, c))
            if field == "pattern":
                lines.append(f"{r} {c}")
            elif field == "integer":
                lines.append(f"{r} {c} {int(rng.integers(-9, 10))}")
            else:
                lines.append(f"{r} {c} {np.float32(rng.standard_normal()):.9g}")
        path = tmp_path / f"r{trial}.mtx"
        path.write_text(f"%%MatrixMarket matrix coordinate {field} {structure}\n% trial {trial}\n{rows} {cols} {len(lines)}\n" + "\n".join(lines) + ("\n" if lines else ""))
        paths.append(path)
    return paths


def test_extended_loader_random_files_match_scipy(tmp_path):
    for path in random_mm_files(tmp_path):
        t = smm.TripletMatrix()
        assert smm.loadMatrix(str(path), t, extended=True) == smm.MatrixLoadStatus.SUCCESS, path.read_text()
        want = sp.coo_matrix(scipy.io.mmread(str(path), spmatrix=True)).astype(np.float32).toarray()
        np.testing.assert_array_equal(dense_from_triplet(t), want, err_msg=path.read_text())
