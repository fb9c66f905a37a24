"""The rounding model behind csrc/dots.cu: sum_squares_serial_kernel, checked on the CPU.

The kernel reproduces the reference's LEFT-TO-RIGHT float sum of squares (BiCGStab's ||r||^2, H:2262-2267) in parallel:
inside one binade of the running sum s = m ulp, adding p = (k + f) ulp is m -> m + k, rounded up when f > 1/2 or when
f = 1/2 and m + k is odd; a term is a map of m that depends on m's parity only, maps compose associatively, and the first
"thread" at which m would reach 2^24 replays its terms with real float additions.  This file restates term_inc /
inc_then / the window loop in Python integers (same formulas as the CUDA source) and compares with numpy's float32
arithmetic on adversarial inputs; the GPU test `test_serial_sum_of_squares_bit_exact` checks the kernel itself."""
import struct

import numpy as np
import pytest

CAP = 1 << 26


def bits(f):
    return struct.unpack("<I", struct.pack("<f", float(f)))[0]


def from_bits(b):
    return np.float32(struct.unpack("<f", struct.pack("<I", b))[0])


def term_inc(pbits, e_eff):                                  # dots.cu: term_inc (branch-free form)
    ep, mant = (pbits >> 23) & 0xFF, pbits & 0x7FFFFF
    P = (mant | 0x800000) if ep else mant
    sh_raw = e_eff - (ep if ep else 1)
    sh = min(max(sh_raw, 0), 31)
    k, rem, half = P >> sh, P & ((1 << sh) - 1), (1 << sh) >> 1
    up = 1 if rem > half else 0
    tie = 1 if (rem == half and sh != 0) else 0
    if ep == 255 or (sh_raw < 0 and P != 0):
        return (CAP, CAP)
    return (k + up + (tie & k), k + up + (tie & (~k & 1)))


def inc_then(a, b):                                          # dots.cu: inc_then
    return (min(a[0] + (b[1] if a[0] & 1 else b[0]), CAP), min(a[1] + (b[0] if a[1] & 1 else b[1]), CAP))


def model_sum(r, threads=8, ept=4):
    p = (r * r).astype(np.float32)
    pb = [bits(x) for x in p]
    n, sb, base = len(r), 0, 0
    while base < n:
        done = 0
        open_ended = False
        while done < threads:
            se = (sb >> 23) & 0xFF
            if se == 255:
                open_ended = True
                break
            e_eff = se if se else 1
            m = ((sb & 0x7FFFFF) | 0x800000) if se else sb
            maps = []
            for t in range(threads):
                mp = (0, 0)
                if t >= done:
                    for j in range(ept):
                        i = base + t * ept + j
                        mp = inc_then(mp, term_inc(pb[i] if i < n else 0, e_eff))
                maps.append(mp)
            before, first, excl, own = (0, 0), None, 0, 0
            for t in range(threads):
                excl = before[1] if m & 1 else before[0]
                own = maps[t][1] if (m + excl) & 1 else maps[t][0]
                if excl >= CAP or own >= CAP or m + excl + own >= (1 << 24):
                    first = t
                    break
                before = inc_then(before, maps[t])

            def make(m2):
                return ((e_eff << 23) | (m2 & 0x7FFFFF)) if m2 >= 0x800000 else m2
            if first is None:
                sb, done = make(m + excl + own), threads
            else:
                cur = from_bits(make(m + excl))
                for j in range(ept):
                    i = base + first * ept + j
                    cur = np.float32(cur + (p[i] if i < n else np.float32(0)))
                sb, done = bits(cur), first + 1
        base += done * ept if open_ended else threads * ept
        if open_ended:
            break
    if ((sb >> 23) & 0xFF) == 255 and (sb & 0x7FFFFF) == 0 and base < n and np.any(np.isnan(r[base:])):
        sb = 0x7FFFFFFF
    return from_bits(sb)


def serial_sum(r):
    p = (r * r).astype(np.float32)
    s = np.float32(0)
    for x in p:
        s = np.float32(s + x)
    return s


@pytest.mark.parametrize("kind", ["normal", "wide_range", "ties", "denormal", "overflow", "nan_inf"])
def test_rounding_model_reproduces_the_serial_sum(kind):
    rng = np.random.default_rng(len(kind))
    with np.errstate(all="ignore"):
        for trial in range(150):
            n = int(rng.integers(1, 200))
            if kind == "normal":
                r = rng.standard_normal(n).astype(np.float32)
            elif kind == "wide_range":
                r = (rng.standard_normal(n) * np.exp2(rng.integers(-70, 60, n))).astype(np.float32)
            elif kind == "ties":
                r = np.exp2(rng.integers(-20, 5, n) / 2.0).astype(np.float32)
                r[rng.random(n) < 0.5] = np.float32(2.0) ** int(rng.integers(-12, 3))
            elif kind == "denormal":
                r = (rng.standard_normal(n) * 1e-20).astype(np.float32)
                r[rng.random(n) < 0.2] = 0
            elif kind == "overflow":
                r = (rng.standard_normal(n) * 1e19).astype(np.float32)
            else:
                r = rng.standard_normal(n).astype(np.float32)
                if rng.random() < 0.5:
                    r[rng.integers(0, n)] = np.nan
                if rng.random() < 0.5:
                    r[rng.integers(0, n)] = np.inf
            got, want = model_sum(r), serial_sum(r)
            assert bits(got) == bits(want) or (np.isnan(got) and np.isnan(want)), (kind, trial, got, want)
