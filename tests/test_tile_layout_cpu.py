"""Host side of the tile-level sweep schedule (csrc/sgs_tiles.cu: proposal, generic verification, layout), without a GPU:
tools/layout_fingerprint.cu includes the source file and runs the set-up code on generated stencils.  The fingerprints
pin every array the kernels read (row order, operand positions, steps, push lists) -- the arrays the GPU parity tests
(`test_sweep_schedules_bit_exact`, the solver tests) were green with; a change of the set-up code must reproduce them."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(ROOT, "tools", "bin", "layout_fingerprint")

# (tile levels, tiles per chain, forward fingerprint, backward fingerprint).  Chains: one warp marches along a grid column of
# tiles; "0" as the fifth argument lays the tiles out in tile-level order instead (single-tile chains, the round-1 order).
GOLDEN = {
    ("64",): (46, 16, "a92db4e914f3e2b9", "c6740e91fb1c60a2"),                # 64^3: 16^3 tiles, 3 * 16 - 2 tile levels
    ("64", "64", "64", "0", "0"): (46, 1, "3af65eb5649ada45", "5de0bdfdeda5a402"),
    ("70", "45", "33"): (37, 18, "e26b9f3a74916599", "1f94717455d29427"),     # ragged 3D grid
    ("200", "150", "1"): (43, 25, "e24e55652b9de722", "625e836d02b6c2bf"),    # 2D, 8 x 8 tiles
    # cluster schedule (sixth argument 1): blocks of 32 chains, pushes through shared memory / DSMEM; the tool also EMULATES it
    ("64", "64", "64", "0", "1", "1"): (46, 16, "02fa02a7b0c79573", "ff4be9ebd2480348"),
    ("70", "45", "33", "0", "1", "1"): (37, 18, "1d933aedc1993fc0", "46952ecfb209f4c8"),
    ("200", "150", "1", "0", "1", "1"): (43, 25, "22027ff9aedf6757", "0b7611ae10617423"),
}


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-ccbin", "/usr/bin/g++", f"-I{ROOT}/include", f"-I{ROOT}/sparse_matrix_math_b200/csrc",
           "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-o", EXE, os.path.join(ROOT, "tools", "layout_fingerprint.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return EXE


@pytest.mark.parametrize("dims", sorted(GOLDEN))
def test_tile_layout_fingerprints(exe, dims):
    out = subprocess.run([exe, *dims], capture_output=True, text=True, timeout=300).stdout
    levels, chain, fwd, bwd = GOLDEN[dims]
    got = re.findall(r"(forward|backward)\s+ok (\d) levels (\d+) time \S+ s fingerprint ([0-9a-f]{16}) chain (\d+) schedule_ok (\d)", out)
    assert [g[0] for g in got] == ["forward", "backward"], out
    assert all(g[1] == "1" and int(g[2]) == levels and int(g[4]) == chain for g in got), out
    # every operand comes from an earlier step of the tile, an earlier tile of the chain or a chain handed out earlier: no deadlock
    assert all(g[5] == "1" for g in got), out
    assert got[0][3] == fwd and got[1][3] == bwd, out
    if len(dims) > 5 and dims[5] == "1":
        # the host emulation of the cluster kernel's data flow (staging, inboxes, published vector) reproduces a plain
        # triangular solve bit for bit and solves every row (no deadlock)
        emu = re.findall(r"(forward|backward)\s+cluster blocks (\d+) tiles (\d+) emulation_ok (\d)", out)
        assert [e[0] for e in emu] == ["forward", "backward"] and all(e[3] == "1" and int(e[1]) > 0 for e in emu), out


@pytest.mark.parametrize("dims", [("24", "20", "16"), ("129", "67", "40"), ("520", "300", "1"), ("33", "9", "70")])
def test_cluster_schedule_emulation(exe, dims):
    """Ragged and thin grids: incomplete blocks (padding chains), short chains, 2D."""
    out = subprocess.run([exe, *dims, "0", "1", "1"], capture_output=True, text=True, timeout=300).stdout
    emu = re.findall(r"(forward|backward)\s+cluster blocks (\d+) tiles (\d+) emulation_ok (\d)", out)
    assert [e[0] for e in emu] == ["forward", "backward"] and all(e[3] == "1" for e in emu), out
    assert len(re.findall(r"schedule_ok 1", out)) == 2, out


def test_tile_proposal_with_cycles_is_rejected(exe):
    """A grid-like band matrix whose +-1 couplings run across the grid lines: tiles (0, J) and (last, J) need each other, the
    verification (Kahn on the tile graph) must refuse the proposal, so the row-level schedule is used instead."""
    out = subprocess.run([exe, "64", "48", "1", "1"], capture_output=True, text=True, timeout=300).stdout
    assert re.search(r"forward\s+ok 0", out), out


# ---------------------------------------------------------------------------------------------------------------
# the host-side factorisations of csrc/sgs.cu (IC(0): bit-identical to the reference's O(rows^2) algorithm; ILU(0):
# extension) against the oracle, without a GPU: tools/factor_fingerprint.cu includes the source file
# ---------------------------------------------------------------------------------------------------------------
FEXE = os.path.join(ROOT, "tools", "bin", "factor_fingerprint")


@pytest.fixture(scope="module")
def fexe():
    os.makedirs(os.path.dirname(FEXE), exist_ok=True)
    cmd = ["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-ccbin", "/usr/bin/g++", f"-I{ROOT}/include", f"-I{ROOT}/sparse_matrix_math_b200/csrc",
           "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "shared", "-o", FEXE, os.path.join(ROOT, "tools", "factor_fingerprint.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return FEXE


def _fnv(buf):
    h = 1469598103934665603
    for b in buf:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


@pytest.mark.parametrize("dims,c", [((12, 10, 9), 0.0), ((12, 10, 9), 0.5), ((31, 17, 1), 0.0), ((7, 7, 7), 0.5)])
def test_host_factorisations_match_the_oracle(fexe, dims, c):
    import matgen
    import oracle_lib as ol
    nx, ny, nz = dims
    g = matgen.convdiff3d(nx, c, ny, nz) if nz > 1 else matgen.poisson2d(nx, ny)
    out = subprocess.run([fexe, str(nx), str(ny), str(nz), str(c)], capture_output=True, text=True, timeout=300).stdout
    lv = re.search(r"levels (\d+) (\d+) order_is_permutation 1", out)       # row-level analysis: nx + ny + nz - 2 levels per sweep
    assert lv and int(lv.group(1)) == int(lv.group(2)) == nx + ny + nz - 2, out
    m = re.search(r"rows (\d+) nnz (\d+) valid 1 diag 1 ic0 rc (\d) ([0-9a-f]{16}) ilu0 rc (\d) ([0-9a-f]{16})", out)
    assert m, out
    assert int(m.group(1)) == g.rows and int(m.group(2)) == g.nnz
    rc, f = ol.ic0_factorize(g)
    assert rc == int(m.group(3)) == 0 and _fnv(f[: g.nnz].tobytes()) == m.group(4)
    rc, lu = ol.ilu0_factorize(g)
    assert rc == int(m.group(5)) == 0 and _fnv(lu[: g.nnz].tobytes()) == m.group(6)
