"""bench.py without a GPU: the byte models of the roofline line and the committed ncu capture they are checked against."""
import importlib.util
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_byte_models():
    b = _bench()
    rows, nnz = 512 ** 3, b.stencil_nnz(512)
    assert nnz == 7 * rows - 6 * 512 * 512
    # values + positions, start, p gathered once, Ap written; and the fused minimum of one CG iteration
    assert b.bytes_spmv_dot(rows, nnz) == 8 * nnz + 4 * (rows + 1) + 8 * rows
    assert b.bytes_cg_iteration(rows, nnz) == 8 * nnz + 48 * rows + 4
    # what the kernels move: 16-bit row starts (257 per group of 256 rows), p read once in the vector passes
    moved = b.bytes_spmv_dot_moved(rows, nnz)
    assert moved == 8 * nnz + 2 * (rows + rows // 256 + 1) + 8 * rows
    assert b.bytes_cg_iteration_two_pass(rows, nnz) == moved + 32 * rows
    assert moved < b.bytes_spmv_dot(rows, nnz) and b.bytes_cg_iteration_two_pass(rows, nnz) < b.bytes_cg_iteration(rows, nnz)


def test_committed_ncu_traffic_matches_the_byte_model():
    """roofline.traffic is read from the committed `ncu --set full` capture of the dominant kernel: DRAM read + write of one
    launch must be what the kernel is said to move (no wasted re-reads), within 2 %."""
    b = _bench()
    rows, nnz = 512 ** 3, b.stencil_nnz(512)
    p = b.traffic_file()
    assert p is not None and os.path.basename(p) == b.TRAFFIC_FILES[0], p
    traffic = b.ncu_traffic_bytes("spmv_rows", 512)
    assert traffic is not None
    assert 0.99 <= traffic / b.bytes_spmv_dot_moved(rows, nnz) <= 1.02, traffic
    assert b.ncu_traffic_bytes("spmv_rows", 256) is None           # the capture belongs to the 512^3 command only


def test_cli_and_recorded_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--reduction"):
        assert flag in out.stdout
    # the last recorded driver-style line carries every key of the contract
    rec = json.load(open(os.path.join(ROOT, "profiles", "r02e_bench.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity", "configs"):
        assert key in rec, key
    assert rec["parity"]["matches_golden"] is True and rec["roofline"]["bound"] == "hbm" and rec["e2e"]["h2d_bytes_per_step"] == 8 * 512 ** 3
    assert len(rec["configs"]) == 5 and all(r.get("parity", {}).get("matches_golden", True) for c in rec["configs"] for r in c["runs"])
