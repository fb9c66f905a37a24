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
