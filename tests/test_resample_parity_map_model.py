"""CPU model of the exact-cumsum algorithm used by csrc/resample.cu (parity maps + raw elements).

The CUDA kernels cannot run here, but their arithmetic can be checked: this file restates
elem_map / compose / apply_map / the classification rule in Python integers and verifies that the
scheme reproduces ``np.cumsum`` (strictly sequential fp64 adds, resampling.py:142) BIT FOR BIT,
including exact ties, zeros, subnormals and binade crossings.  It documents the algorithm the GPU
tests then check end to end."""
import struct

import numpy as np
import pytest


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def from_bits(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def efield(x):
    return (bits(x) >> 52) & 0x7FF


def sint(x):
    b = bits(x)
    m = b & ((1 << 52) - 1)
    return m | (1 << 52) if (b >> 52) & 0x7FF else m


def rebuild(e, si):
    return from_bits((e << 52) | (si & ((1 << 52) - 1)))


def limit_of(e):
    return (1 << 53) if e else (1 << 52)


def elem_map(w, e):
    """Parity map of adding w in binade e, computed as csrc/resample.cu does: two IEEE adds on the
    binade base (even mantissa) and base+1ulp (odd mantissa); bit patterns of positive doubles
    are linear in units of the ulp inside a binade."""
    base = e << 52
    d0 = bits(from_bits(base) + w) - base
    d1 = bits(from_bits(base + 1) + w) - (base + 1)
    assert d1 - d0 in (-1, 0, 1)
    return (d0, d1)


def elem_map_reference(w, e):
    """The same map from first principles (integer rounding of w / ulp)."""
    b = bits(w)
    ew = (b >> 52) & 0x7FF
    mw = (b & ((1 << 52) - 1)) | ((1 << 52) if ew else 0)
    sh = max(ew, 1) - max(e, 1)
    if sh >= 0:
        a = mw << sh
        return (a, a)
    r = -sh
    if r >= 64:
        return (0, 0)
    a = mw >> r
    rem = mw & ((1 << r) - 1)
    half = 1 << (r - 1)
    if rem > half:
        return (a + 1, a + 1)
    if rem < half:
        return (a, a)
    return (a + (a & 1), a + 1 - (a & 1))


def compose(f, g):
    return (f[0] + (g[1] if (f[0] & 1) else g[0]), f[1] + (g[1] if ((f[1] + 1) & 1) else g[0]))


def apply_map(S, m, e):
    """State advanced on its BIT PATTERN (csrc/resample.cu:apply_map)."""
    sb = bits(S)
    assert (sb >> 52) == e, "start state left the assumed binade"
    sb += m[1] if (sb & 1) else m[0]
    assert (sb >> 52) == e, "end state left the assumed binade"
    return from_bits(sb)


def exact_cumsum_model(w, chunk=64):
    """Approximate prefix (pairwise, different rounding order than cumsum) -> classification ->
    parity maps composed per chunk in a tree-ish order -> sequential chain over chunks."""
    n = len(w)
    eps = (n + 4096.0) * 2.0 ** -52
    approx = np.zeros(n + 1)
    # a deliberately different summation order: chunk sums first, then within chunks
    for c0 in range(0, n, chunk):
        seg = w[c0:c0 + chunk]
        approx[c0 + 1:c0 + 1 + len(seg)] = approx[c0] + np.cumsum(seg[::-1])[::-1][0] * 0 + np.cumsum(seg)
    out = np.zeros(n)
    S = 0.0
    nraw = 0
    j = 0
    while j < n:
        # gather a maximal run of clean elements (composed right-to-left to vary association)
        before, after = approx[j], approx[j + 1]
        lo, hi = before * (1.0 - eps), after * (1.0 + eps)
        e = efield(lo)
        if efield(hi) != e:
            S = S + w[j]            # raw element: a true fp64 add
            out[j] = S
            nraw += 1
            j += 1
            continue
        run = []
        k = j
        while k < n:
            lo, hi = approx[k] * (1.0 - eps), approx[k + 1] * (1.0 + eps)
            if efield(lo) != e or efield(hi) != e:
                break
            run.append(elem_map(w[k], e))
            k += 1
        # inclusive prefix maps by a Hillis-Steele style doubling (parallel association order)
        inc = list(run)
        d = 1
        while d < len(inc):
            nxt = list(inc)
            for q in range(d, len(inc)):
                nxt[q] = compose(inc[q - d], inc[q])
            inc = nxt
            d *= 2
        for q, m in enumerate(inc):
            out[j + q] = apply_map(S, m, e)
        S = out[k - 1]
        j = k
    return out, nraw


CASES = {
    "random": lambda rng, n: rng.random(n),
    "heavy": lambda rng, n: rng.random(n) ** 4,
    "zeros": lambda rng, n: np.where(rng.random(n) < 0.3, 0.0, rng.random(n)),
    "leading_zeros": lambda rng, n: np.concatenate([np.zeros(n // 3), rng.random(n - n // 3)]),
    "dyadic_ties": lambda rng, n: rng.integers(0, 8, n) * 2.0 ** -55 + rng.integers(0, 3, n) * 2.0 ** -20,
    "tiny_then_big": lambda rng, n: np.concatenate([rng.random(n // 2) * 1e-300, rng.random(n - n // 2)]),
    "subnormal": lambda rng, n: rng.integers(0, 1 << 20, n).astype(np.float64) * 5e-324,
    "huge_range": lambda rng, n: 10.0 ** rng.uniform(-30, 0, n),
    "powers_of_two": lambda rng, n: 2.0 ** rng.integers(-60, -5, n).astype(np.float64),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("n", [1, 5, 257, 3000])
def test_parity_maps_reproduce_sequential_cumsum(name, n):
    rng = np.random.default_rng(hash(name) % 1000 + n)
    w = np.asarray(CASES[name](rng, n), dtype=np.float64)
    if name not in ("subnormal", "dyadic_ties", "tiny_then_big", "powers_of_two") and w.sum() > 0:
        w = w / w.sum()
    want = np.cumsum(w)
    got, nraw = exact_cumsum_model(w)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (name, n, nraw)


def test_tie_rule_matches_ieee():
    # S even / odd mantissa with an exact half-ulp addend: round-half-even depends on the parity
    for S in (1.0, 1.0 + 2.0 ** -52, 0.75, 0.75 + 2.0 ** -53):
        e = efield(S)
        q = 2.0 ** (max(e, 1) - 1075)
        for mult in (0.5, 1.5, 2.5, 1.0, 0.25, 0.75):
            w = mult * q
            m = elem_map(w, e)
            assert apply_map(S, m, e) == S + w, (S, mult)


def test_dadd_trick_equals_integer_rounding():
    rng = np.random.default_rng(0)
    for e in (0, 1, 2, 500, 1000, 1022, 1023, 1030):
        q = 2.0 ** (max(e, 1) - 1075)
        top = 2.0 ** (max(e, 1) - 1023)
        ws = list(rng.random(200) * top * 0.9) + [k * q * 0.5 for k in range(0, 40)] + [0.0, q / 2, q / 4, 3 * q / 4, q]
        for w in ws:
            m = elem_map(w, e)
            if m[0] < (1 << 52) and m[1] < (1 << 52):
                assert m == elem_map_reference(w, e), (e, w)
