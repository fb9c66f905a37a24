#!/usr/bin/env python
"""Read a SMM_B200_SGS_TRACE dump (forward sweep, tiles in level order): where does a tile's time go?"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4).astype(np.int64)
t0 = t[:, 0].min()
claim, ready, done, sm = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3]
print(f"tiles {len(t)}  sweep {done.max()/1e3:.1f} us")
print(f"wait  (claim->ready): mean {np.mean(ready-claim):.0f} ns  median {np.median(ready-claim):.0f}  p99 {np.percentile(ready-claim,99):.0f}")
print(f"solve (ready->done) : mean {np.mean(done-ready):.0f} ns  median {np.median(done-ready):.0f}  p99 {np.percentile(done-ready,99):.0f}")
# walk the critical path backwards is not possible without the graph; print a few evenly spaced tiles instead
for i in np.linspace(0, len(t) - 1, 24).astype(int):
    print(f"tile {i:8d} sm {sm[i]:3d} claim {claim[i]/1e3:9.2f} ready {ready[i]/1e3:9.2f} done {done[i]/1e3:9.2f} us")

# chain schedule (tile index = chain rank * chain_len + position): python tools/sgs_trace.py dump chain_len TJ TK
if len(sys.argv) > 4:
    clen, TJ, TK = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    nch = len(t) // clen
    d = done.reshape(nch, clen); r = ready.reshape(nch, clen); c = claim.reshape(nch, clen)
    period = np.diff(d, axis=1)
    print(f"chains {nch} x {clen} tiles: period along a chain mean {period.mean():.0f} ns median {np.median(period):.0f}; "
          f"solve {np.mean(d - r):.0f} ns; wait {np.mean(r - c):.0f} ns; claim->claim gap {np.mean(c[:, 1:] - d[:, :-1]):.0f} ns")
    # chain ranks are ordered by (J + K, id): recover (J, K) of every rank
    ids = sorted(range(TJ * TK), key=lambda a: ((a % TJ) + (a // TJ), a))
    rank = {a: i for i, a in enumerate(ids)}
    lagJ, lagK = [], []
    for K in range(TK):
        for J in range(TJ):
            a = K * TJ + J
            if J + 1 < TJ: lagJ.append(np.mean(d[rank[a + 1]] - d[rank[a]]))
            if K + 1 < TK: lagK.append(np.mean(d[rank[a + TJ]] - d[rank[a]]))
    print(f"lag to the +J neighbour chain: mean {np.mean(lagJ):.0f} ns median {np.median(lagJ):.0f}; +K: mean {np.mean(lagK):.0f} median {np.median(lagK):.0f}")
    first, last = rank[0], rank[TJ * TK - 1]
    print(f"first chain: start {c[first,0]/1e3:.1f} us end {d[first,-1]/1e3:.1f} us; last chain: start of solve {r[last,0]/1e3:.1f} us end {d[last,-1]/1e3:.1f} us")
    for a in (0, TJ * TK // 2 + TJ // 2, TJ * TK - 1):
        k = rank[a]
        print(f"chain (J={a % TJ}, K={a // TJ}) rank {k}: claim0 {c[k,0]/1e3:.1f} ready0 {r[k,0]/1e3:.1f} done0 {d[k,0]/1e3:.1f} ... done_last {d[k,-1]/1e3:.1f} us; mean period {np.mean(np.diff(d[k])):.0f} ns")
