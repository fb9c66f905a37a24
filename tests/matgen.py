"""Deterministic synthetic inputs for the BASELINE.json configs (numpy; test/bench infrastructure).

Every generator is defined with integer / correctly-rounded float arithmetic only, so that the device-side
generators in sparse_matrix_math_b200/csrc (smm_gen_*) produce bit-identical CSR arrays; tests compare them.

  G1 poisson2d(nx, ny)        5-point Laplacian, Dirichlet: diag 4, neighbours -1              (config 1)
  G2 convdiff3d(n, c)         7-point convection-diffusion: diag 6, -x/-y/-z = -1-c, +x/+y/+z = -1+c
                              (c = 0.5: configs 2, 3; c = 0: 3D Poisson, config 5)
  G4 powerlaw(n, seed)        row length l_i = floor(49152 / sqrt(m_i)), m_i uniform in [1, 2^24]
                              (pdf ~ l^-3, min 12, mean ~24); one hash-placed column per stratum of
                              [0,n)\\{i}; off-diagonals uniform(-1,1)/l_i, diagonal 2 (strictly
                              diagonally dominant)                                              (config 4)
  xstar(n)                    x*_i = (splitmix64(0xB200, i) >> 40) / 2^24 in [0,1); rhs b = A x*

Natural row order row = (k*ny + j)*nx + i; columns ascending inside a row (the order the reference's
std::map based TripletMatrix -> CSRMatrix conversion produces, H:1606-1641).
"""
import numpy as np

try:
    from .oracle_lib import CSR
except ImportError:  # imported as a top-level module (bench.py, scripts)
    from oracle_lib import CSR

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, idx):
    """splitmix64 of stream position idx (vectorised, uint64 wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.asarray(idx, np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def xstar(n, seed=0xB200):
    return ((splitmix64(seed, np.arange(n, dtype=np.uint64)) >> np.uint64(40)).astype(np.float32) / np.float32(1 << 24)).astype(np.float32)


def _stencil(nx, ny, nz, offsets_vals):
    """offsets_vals: list of (di, dj, dk, value) in ascending column order."""
    n = nx * ny * nz
    idx = np.arange(n, dtype=np.int64)
    i = idx % nx
    j = (idx // nx) % ny
    k = idx // (nx * ny)
    cols, vals, valid = [], [], []
    for di, dj, dk, v in offsets_vals:
        ok = (i + di >= 0) & (i + di < nx) & (j + dj >= 0) & (j + dj < ny) & (k + dk >= 0) & (k + dk < nz)
        cols.append(idx + di + dj * nx + dk * nx * ny)
        vals.append(np.full(n, v, np.float32))
        valid.append(ok)
    valid = np.stack(valid, 1)
    cols = np.stack(cols, 1)
    vals = np.stack(vals, 1)
    counts = valid.sum(1)
    start = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=start[1:])
    return CSR(n, n, start.astype(np.int32), cols[valid].astype(np.int32), vals[valid])


def poisson2d(nx, ny):
    return _stencil(nx, ny, 1, [(0, -1, 0, -1.0), (-1, 0, 0, -1.0), (0, 0, 0, 4.0), (1, 0, 0, -1.0), (0, 1, 0, -1.0)])


def convdiff3d(n, c=0.5, ny=None, nz=None):
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    lo, hi = np.float32(-1.0 - c), np.float32(-1.0 + c)
    return _stencil(n, ny, nz, [(0, 0, -1, lo), (0, -1, 0, lo), (-1, 0, 0, lo), (0, 0, 0, 6.0),
                                (1, 0, 0, hi), (0, 1, 0, hi), (0, 0, 1, hi)])


def poisson3d(n, ny=None, nz=None):
    return convdiff3d(n, 0.0, ny, nz)


POWERLAW_A = 49152  # l = floor(A / sqrt(m))


def powerlaw_row_lengths(n, seed=0x5EED, cap=131072):
    m = (splitmix64(seed, np.arange(n, dtype=np.uint64)) >> np.uint64(40)).astype(np.int64) + 1
    q = (POWERLAW_A * POWERLAW_A) // m
    l = np.floor(np.sqrt(q.astype(np.float64))).astype(np.int64)
    l = np.where(l * l > q, l - 1, l)
    l = np.where((l + 1) * (l + 1) <= q, l + 1, l)          # exact integer sqrt
    return np.minimum(np.minimum(l, cap), n).astype(np.int64)


def powerlaw(n, seed=0x5EED, cap=131072):
    l = powerlaw_row_lengths(n, seed, cap)
    start = np.zeros(n + 1, np.int64)
    np.cumsum(l, out=start[1:])
    nnz = int(start[n])
    d = l - 1                                               # off-diagonals per row
    rows = np.repeat(np.arange(n, dtype=np.int64), d)
    off_start = np.zeros(n + 1, np.int64)
    np.cumsum(d, out=off_start[1:])
    k = np.arange(int(off_start[n]), dtype=np.int64) - np.repeat(off_start[:-1], d)
    dd = np.repeat(d, d)
    S = n - 1
    lo = (k * S) // dd
    hi = ((k + 1) * S) // dd
    key = (rows.astype(np.uint64) << np.uint64(20)) | k.astype(np.uint64)
    h1 = splitmix64(np.uint64(seed) ^ np.uint64(0xA5A5A5A5), key)
    slot = lo + (h1 % (hi - lo).astype(np.uint64)).astype(np.int64)
    col = np.where(slot < rows, slot, slot + 1)
    h2 = splitmix64(np.uint64(seed) ^ np.uint64(0x5A5A5A5A), key)
    u = (h2 >> np.uint64(40)).astype(np.float32) / np.float32(1 << 23) - np.float32(1.0)
    val = (u / np.repeat(l, d).astype(np.float32)).astype(np.float32)
    below = (slot < rows)
    pos = np.repeat(start[:-1], d) + k + np.where(below, 0, 1)
    nbelow = np.zeros(n, np.int64)
    np.add.at(nbelow, rows, below.astype(np.int64))
    positions = np.empty(nnz, np.int32)
    values = np.empty(nnz, np.float32)
    positions[pos] = col
    values[pos] = val
    dpos = start[:-1] + nbelow
    positions[dpos] = np.arange(n, dtype=np.int32)
    values[dpos] = 2.0
    return CSR(n, n, start.astype(np.int32), positions, values)


def symmetrize(m):
    """(A + A^T)/2 with the union pattern (scipy; small test sizes only)."""
    import scipy.sparse as sp
    a = sp.csr_matrix((m.values, m.positions, m.start), shape=(m.rows, m.cols))
    s = ((a + a.T) * np.float32(0.5)).tocsr()
    s.sort_indices()
    return CSR(m.rows, m.cols, s.indptr.astype(np.int32), s.indices.astype(np.int32), s.data.astype(np.float32))


def random_csr(rows, cols, density_rows, rng, empty_row_fraction=0.0, max_len=None):
    """Ragged random matrix for edge-case tests: row lengths from density_rows (callable or int)."""
    lens = np.array([density_rows(r) if callable(density_rows) else density_rows for r in range(rows)], np.int64)
    if max_len is not None:
        lens = np.minimum(lens, max_len)
    lens = np.minimum(lens, cols)
    if empty_row_fraction > 0:
        lens[rng.random(rows) < empty_row_fraction] = 0
    start = np.zeros(rows + 1, np.int64)
    np.cumsum(lens, out=start[1:])
    positions = np.empty(int(start[rows]), np.int32)
    for r in range(rows):
        if lens[r]:
            positions[start[r]:start[r + 1]] = np.sort(rng.choice(cols, int(lens[r]), replace=False))
    values = rng.uniform(-1, 1, int(start[rows])).astype(np.float32)
    return CSR(rows, cols, start.astype(np.int32), positions, values)
