"""The drop-in C++17 header (include/sparse_matrix_math.h -> smm_b200.hpp): builds against libsmm_b200.so, refuses
non-float scalar types on the hot path at compile time, and (on a GPU) passes the reference's own test scenarios
re-hosted in tests/cpp/dropin_tests.cpp."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "cpp", "build")
CXX = "/usr/bin/g++"
LIBDIR = os.path.join(ROOT, "sparse_matrix_math_b200")


def compile_cpp(src, out, extra=()):
    os.makedirs(BUILD, exist_ok=True)
    cmd = [CXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", f"-I{HERE}/cpp",
           f'-DASSET_PATH="{HERE}/golden/"', src, "-o", out, f"-L{LIBDIR}", "-lsmm_b200", f"-Wl,-rpath,{LIBDIR}",
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", *extra]
    return subprocess.run(cmd, capture_output=True, text=True)


@pytest.fixture(scope="module")
def built_lib():
    from sparse_matrix_math_b200 import build
    return build.build()


def test_dropin_tests_compile_and_link(built_lib):
    r = compile_cpp(os.path.join(HERE, "cpp", "dropin_tests.cpp"), os.path.join(BUILD, "dropin_tests"))
    assert r.returncode == 0, r.stderr[-3000:]


def test_host_containers_are_scalar_generic(built_lib, tmp_path):
    # containers, iterators and loaders are templates usable with double (as in the reference's TEST_CASE_TEMPLATEs)
    src = tmp_path / "generic.cpp"
    src.write_text('''
#include "sparse_matrix_math.h"
int main() {
    SMM::TripletMatrix<double> t(3, 3);
    t.addEntry(0, 0, 1.0); t.addEntry(2, 1, 2.0); t.addEntry(2, 1, 0.5);
    SMM::CSRMatrix<double> m(t);
    double sum = 0;
    for (const auto& el : m) sum += el.getValue();
    SMM::Vector<double> v(3, 1.0);
    v += v;
    return (m.getNonZeroCount() == 2 && sum == 3.5 && m.getValue(2, 1) == 2.5 && v[0] == 2.0) ? 0 : 1;
}
''')
    out = str(tmp_path / "generic")
    r = compile_cpp(str(src), out)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([out]).returncode == 0        # host-only code: runs without a GPU


def test_smmdt_loader_status_codes_match_the_reference(built_lib, tmp_path):
    """loadSMMDTMatrix (H:2611-2646) on well-formed and cut-off files: same MatrixLoadStatus and entries as the real reference
    (oracle/_ref) -- including FAILED_TO_PARSE_FILE for a file that ends with its last row (the trailing ignore + fail check)."""
    import sys
    sys.path.insert(0, HERE)
    import oracle_lib as ol
    if not ol.ref_available():
        pytest.skip("oracle/_ref not built")
    files = {
        "ok.smmdt": "2 3\n{\n{1.5,0,2},\n{0,-3,0}\n}",
        "ok_newline.smmdt": "2 3\n{\n{1.5,0,2},\n{0,-3,0}\n}\n",
        "no_closing_line.smmdt": "2 3\n{\n{1.5,0,2},\n{0,-3,0}\n",
        "ends_with_last_row.smmdt": "2 3\n{\n{1.5,0,2},\n{0,-3,0}",
        "cut_in_row.smmdt": "2 3\n{\n{1.5,0,2},\n{0,-3",
        "bad_header.smmdt": "x 3\n{\n{1}\n}",
    }
    src = tmp_path / "smmdt.cpp"
    src.write_text('''
#include <cstdio>
#include "sparse_matrix_math.h"
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        SMM::TripletMatrix<float> t;
        const int st = (int)SMM::loadMatrix(argv[i], t);
        std::printf("%d %d\\n", st, st == 0 ? t.getNonZeroCount() : -1);
    }
    return 0;
}
''')
    out = str(tmp_path / "smmdt")
    r = compile_cpp(str(src), out)
    assert r.returncode == 0, r.stderr[-3000:]
    paths = []
    for name, text in files.items():
        p = tmp_path / name
        p.write_text(text)
        paths.append(str(p))
    got = [tuple(int(v) for v in line.split()) for line in subprocess.run([out, *paths], capture_output=True, text=True).stdout.splitlines()]
    assert len(got) == len(paths)
    for path, (st, nnz) in zip(paths, got):
        rst, m = ol.ref_load_matrix(path)
        assert st == rst, (path, st, rst)
        if st == 0:
            assert nnz == m.nnz, path
    assert got[3][0] != 0 and got[0][0] == 0          # the case ADVICE named: ends with its last row -> FAILED_TO_PARSE_FILE


def test_hot_path_refuses_other_scalars_at_compile_time(built_lib, tmp_path):
    src = tmp_path / "dbl.cpp"
    src.write_text('''
#include "sparse_matrix_math.h"
int main() {
    SMM::TripletMatrix<double> t(2, 2);
    SMM::CSRMatrix<double> m(t);
    double b[2] = {1, 1}, x[2] = {0, 0};
    return (int)SMM::ConjugateGradient<double>(m, b, x, x, -1, 1e-8);
}
''')
    r = compile_cpp(str(src), str(tmp_path / "dbl"))
    assert r.returncode != 0 and "float only (no CPU fallback)" in r.stderr


@pytest.mark.gpu
def test_dropin_tests_run_on_gpu(built_lib):
    exe = os.path.join(BUILD, "dropin_tests")
    r = compile_cpp(os.path.join(HERE, "cpp", "dropin_tests.cpp"), exe)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, cwd=ROOT)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout


EXAMPLE = os.path.join(ROOT, "examples", "solve_mtx.cpp")


def test_example_compiles(built_lib):
    r = compile_cpp(EXAMPLE, os.path.join(BUILD, "solve_mtx"))
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("matrix,solver,order", [("mesh1em1.mtx", "cg", 1), ("mesh1e1.mtx", "bicgstab-sgs", 0), ("mesh1em6.mtx", "cg-ic0", 1),
                                                 ("mesh1em1.mtx", "bicgstab-ilu0", 0), ("mesh1e1.mtx", "cgs", 0)])
def test_example_runs_on_gpu(built_lib, matrix, solver, order):
    """examples/solve_mtx.cpp: the reference's asset files through the drop-in header, end to end."""
    exe = os.path.join(BUILD, "solve_mtx")
    r = compile_cpp(EXAMPLE, exe)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, os.path.join(HERE, "golden", matrix), solver, "1e-4", str(order)], capture_output=True, text=True, timeout=120)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "status 0" in r.stdout and "48 rows" in r.stdout
    err = float(r.stdout.split("max |x - 1| = ")[1].split(",")[0])
    assert err < 1e-3
