#!/usr/bin/env python
"""Read a per-patch trace of the line schedule's forward sweep (SMM_B200_SGS_LINES=1 SMM_B200_SGS_TRACE=<file> python tools/sgs_bench.py N):
    python tools/sgs_lines_trace.py <file> ny nz      (3D: patches of 8 x 4 lines)"""
import sys
import numpy as np
path, ny, nz = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
A, B = 8, 4
NJ, NK = -(-ny // A), -(-nz // B)
t = np.fromfile(path, dtype=np.uint64).reshape(-1, 4).astype(np.int64)
ids = np.arange(NJ * NK)
T = A * (ids % NJ) + B * (ids // NJ)
by_rank = np.argsort(T, kind="stable")
rank_of = np.empty_like(by_rank); rank_of[by_rank] = np.arange(len(by_rank))
t0 = t[:, 0].min()
claim, start, end, sm = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, t[:, 3]
print(f"{len(t)} patches, sweep {end.max():.1f} us; run time of a patch (start -> end): median {np.median(end - start):.1f} us, "
      f"p10 {np.percentile(end - start, 10):.1f}, p90 {np.percentile(end - start, 90):.1f}")
def lag(dj, dk):
    out = []
    for K in range(NK):
        for J in range(NJ):
            if J - dj < 0 or K - dk < 0: continue
            a, b = rank_of[K * NJ + J], rank_of[(K - dk) * NJ + (J - dj)]
            out.append(start[a] - start[b])
    return np.array(out)
for name, l in (("j", lag(1, 0)), ("k", lag(0, 1))):
    print(f"start of a patch minus start of its {name}-predecessor: median {np.median(l):.2f} us, p10 {np.percentile(l, 10):.2f}, p90 {np.percentile(l, 90):.2f}")
# the critical path: walk back from the patch that ends last through the predecessor that started later
J, K = NJ - 1, NK - 1
path_ = []
while True:
    a = rank_of[K * NJ + J]
    path_.append((J, K, start[a], end[a], claim[a]))
    c = []
    if J > 0: c.append((start[rank_of[K * NJ + J - 1]], J - 1, K))
    if K > 0: c.append((start[rank_of[(K - 1) * NJ + J]], J, K - 1))
    if not c: break
    _, J, K = max(c)
print("critical path (last patch back to the first), start / end / claim in us:")
for e in path_[:: max(1, len(path_) // 24)]:
    print(f"  patch ({e[0]:3d},{e[1]:3d})  start {e[2]:8.1f}  end {e[3]:8.1f}  claimed {e[4]:8.1f}")
late = (claim > start - 0.5).sum()
print(f"patches claimed less than 0.5 us before they could start (no CTA was free earlier): {late}")
