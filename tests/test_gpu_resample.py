"""GPU parity: systematic / stratified resampling must equal the reference bit for bit
(index arrays are integers: north_star's bar is exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_sys(w, u):
    import torch
    from filterpy_b200.monte_carlo import ResamplePlan
    wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)).cuda()
    plan = ResamplePlan(len(w))
    idx = plan.systematic(wd, u).cpu().numpy()
    return idx, plan.info(), float(plan.cumsum_last.item())


def run_str(w, U):
    import torch
    from filterpy_b200.monte_carlo import ResamplePlan
    wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)).cuda()
    Ud = torch.from_numpy(np.ascontiguousarray(U, dtype=np.float64)).cuda()
    plan = ResamplePlan(len(w))
    idx = plan.stratified(wd, Ud).cpu().numpy()
    return idx, plan.info(), float(plan.cumsum_last.item())


def test_golden_vectors_from_reference(golden):
    """Every (weights, u) pair the reference was run on (tests/golden/make_golden.py)."""
    g = golden("resample")
    idx, info, _ = run_sys(np.array([.1, .2, .3, .4]), 0.5)
    assert idx.tolist() == [1, 2, 3, 3]
    for (i, N, ok, ok_s, seed) in g["meta"]:
        w, u, U = g["w%d" % i], float(g["u%d" % i]), g["U%d" % i]
        idx, info, cl = run_sys(w, u)
        assert cl == np.cumsum(w)[-1]
        if ok:
            assert info[0] == 0 and idx.dtype == np.int32
            assert np.array_equal(idx, g["sys%d" % i]), ("systematic", i, N, info)
        else:
            assert info[0] > 0
        idx, info, _ = run_str(w, U)
        if ok_s:
            assert info[0] == 0
            assert np.array_equal(idx, g["str%d" % i]), ("stratified", i, N, info)
        else:
            assert info[0] > 0


@pytest.mark.parametrize("kind", ["heavy", "uniform", "random", "zeros", "degenerate", "dyadic"])
@pytest.mark.parametrize("N", [2047, 2048, 2049, 100003, 1 << 20])
def test_systematic_vs_oracle(kind, N):
    from filterpy_b200.common import workloads as wl
    from oracle import resample as ors
    w = wl.resample_weights(N, kind, seed=N % 1000)
    for u in (0.0, 0.37454011884736254, 0.9999999999):
        idx, info, cl = run_sys(w, u)
        assert info[1] == 0, "sequential fallback was taken: %s" % info
        assert cl == np.cumsum(w)[-1]
        try:
            want = ors.systematic_resample_c(w, u)
        except IndexError:
            assert info[0] > 0
            continue
        assert info[0] == 0, info
        assert np.array_equal(idx, want), (kind, N, u, info, np.flatnonzero(idx != want)[:5])


@pytest.mark.parametrize("kind", ["heavy", "uniform", "zeros", "degenerate"])
@pytest.mark.parametrize("N", [1000, 4097, 300007])
def test_stratified_vs_oracle(kind, N):
    from filterpy_b200.common import workloads as wl
    from oracle import resample as ors
    w = wl.resample_weights(N, kind, seed=7)
    U = np.random.default_rng(N).random(N)
    idx, info, _ = run_str(w, U)
    assert info[1] == 0
    try:
        want = ors.stratified_resample_c(w, U)
    except IndexError:
        assert info[0] > 0
        return
    assert np.array_equal(idx, want), (kind, N, info)


def test_adversarial_rounding_cases():
    """Ties, subnormals, huge dynamic range, leading zeros: the cumulative sum itself must be the
    sequential one (cumsum_last is exact) and the indexes must match."""
    from oracle import resample as ors
    rng = np.random.default_rng(5)
    n = 50000
    cases = {
        "ties": rng.integers(0, 8, n) * 2.0 ** -55 + rng.integers(0, 3, n) * 2.0 ** -20,
        "leading_zeros": np.concatenate([np.zeros(n // 3), rng.random(n - n // 3)]),
        "huge_range": 10.0 ** rng.uniform(-30, 0, n),
        "pow2": 2.0 ** rng.integers(-60, -5, n).astype(np.float64),
        "tiny_then_big": np.concatenate([rng.random(n // 2) * 1e-300, rng.random(n - n // 2)]),
    }
    for name, w in cases.items():
        w = w / w.sum()
        idx, info, cl = run_sys(w, 0.123)
        assert cl == np.cumsum(w)[-1], name
        try:
            want = ors.systematic_resample_c(w, 0.123)
        except IndexError:
            assert info[0] > 0, name
            continue
        assert np.array_equal(idx, want), (name, info)


def test_invalid_weights_use_literal_fallback():
    """Negative weights: the reference's merge loop is still well defined; the engine switches to
    its literal single-thread transcription (info[1] == 1) and must agree."""
    from oracle import resample as ors
    rng = np.random.default_rng(1)
    w = rng.random(3000); w[100] = -0.2; w /= w.sum()
    idx, info, _ = run_sys(w, 0.4)
    assert info[1] == 1
    want = ors.resample_loop(w, ors.positions_systematic(len(w), 0.4))
    assert np.array_equal(idx, want)


def test_unnormalised_overflow_flag():
    w = np.full(1000, 0.0005)        # sums to 0.5: the reference raises IndexError
    idx, info, _ = run_sys(w, 0.3)
    assert info[0] == 500
    from filterpy_b200.monte_carlo import systematic_resample
    with pytest.raises(IndexError):
        systematic_resample(w)


def test_public_function_reproduces_reference_rng_stream(golden):
    """np.random.seed(s); systematic_resample(w) must equal the reference called the same way."""
    from filterpy_b200.monte_carlo import systematic_resample, stratified_resample
    g = golden("resample")
    for (i, N, ok, ok_s, seed) in g["meta"]:
        if not (ok and ok_s) or N < 64:
            continue
        w = g["w%d" % i]
        np.random.seed(int(seed))
        got = systematic_resample(w)
        assert got.dtype == np.int32 and np.array_equal(got, g["sys%d" % i])
        np.random.seed(int(seed))
        assert np.array_equal(stratified_resample(w), g["str%d" % i])


def test_full_size_64M_bit_exact_and_properties():
    """BASELINE config 5 size (2^26 particles): bit-exact against the C oracle, plus the
    size-independent properties: sorted output, counts = floor/ceil(N w) for systematic."""
    import torch
    from filterpy_b200.common import workloads as wl
    from filterpy_b200.monte_carlo import ResamplePlan
    from oracle import resample as ors
    N = 1 << 26
    w = wl.resample_weights(N, "heavy", seed=97)
    np.random.seed(7); u = np.random.random()
    wd = torch.from_numpy(w).cuda()
    plan = ResamplePlan(N)
    idx_d = plan.systematic(wd, u)
    info = plan.info()
    assert info[0] == 0 and info[1] == 0, info
    assert bool((idx_d[1:] >= idx_d[:-1]).all())
    counts = torch.bincount(idx_d.long(), minlength=N).double()
    assert float((counts - wd * N).abs().max()) <= 1.0 + 1e-6
    idx = idx_d.cpu().numpy()
    want = ors.systematic_resample_c(w, u)
    assert np.array_equal(idx, want)
