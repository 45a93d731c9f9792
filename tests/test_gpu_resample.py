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
@pytest.mark.parametrize("N", [2047, 2049, 4095, 4096, 4097, 100003, 1 << 20])
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


@pytest.mark.parametrize("kind", ["heavy", "uniform", "degenerate", "zeros"])
@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_resample_equals_whole_array(kind, shards):
    """bke_resample_shard: the shards of one particle set, processed the way the ranks of a multi-GPU
    job do (approximate carry from the shard sums, exact carry handed from shard to shard), produce
    exactly the single-array result — here all shards run on one GPU, in rank order."""
    import ctypes
    import torch
    from filterpy_b200 import _lib
    from filterpy_b200.common import workloads as wl
    from filterpy_b200.distributed import shard_bounds
    from oracle import resample as ors
    N, u = 300007, 0.6180339887
    w = wl.resample_weights(N, kind, seed=11)
    want = ors.systematic_resample_c(w, u)
    lib = _lib.load()
    wd = torch.from_numpy(w).cuda()
    b = shard_bounds(N, shards)
    out = np.full(N, -1, dtype=np.int32)
    st = torch.cuda.current_stream().cuda_stream
    sums = []
    state = []
    for r in range(shards):
        n_loc = int(b[r + 1] - b[r])
        ws_bytes = int(lib.bke_resample_workspace_bytes(n_loc))
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
        ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
        sl = wd[int(b[r]):int(b[r + 1])]
        s = torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib.check(lib.bke_weights_sum(n_loc, sl.data_ptr(), s.data_ptr(), ws_ptr, ws_bytes, st))
        sums.append(s)
        state.append((n_loc, ws, ws_ptr, ws_bytes, sl))
    carry_exact = None
    covered = 0
    for r in range(shards):
        n_loc, ws, ws_ptr, ws_bytes, sl = state[r]
        cap = N
        idx = torch.full((cap,), -7, dtype=torch.int32, device="cuda")
        info = torch.zeros(8, dtype=torch.int32, device="cuda")
        rng_t = torch.zeros(2, dtype=torch.int64, device="cuda")
        carry_out = torch.zeros(1, dtype=torch.float64, device="cuda")
        capx = torch.stack(sums[:r]).sum().reshape(1) if r else torch.zeros(1, dtype=torch.float64, device="cuda")
        a = _lib.ResampleShardArgs()
        a.n_local, a.n_global, a.j_offset, a.capacity = n_loc, N, int(b[r]), cap
        a.weights, a.u = sl.data_ptr(), u
        a.carry_approx = capx.data_ptr()
        a.carry_exact = None if carry_exact is None else carry_exact.data_ptr()
        a.indexes, a.out_range, a.carry_out = idx.data_ptr(), rng_t.data_ptr(), carry_out.data_ptr()
        a.workspace, a.workspace_bytes, a.info = ws_ptr, ws_bytes, info.data_ptr()
        a.is_last = 1 if r == shards - 1 else 0
        a.phase = 1
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))
        a.phase = 2
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))
        a.phase = 4
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))
        lo, hi = [int(v) for v in rng_t.cpu().numpy()]
        inf = info.cpu().numpy()
        assert inf[1] == 0 and inf[6] == 0, inf
        assert lo == covered, (r, lo, covered)
        out[lo:hi] = idx[:hi - lo].cpu().numpy()
        covered = hi
        carry_exact = carry_out
        assert float(carry_out.item()) == np.cumsum(w[:int(b[r + 1])])[-1]
    assert covered == N
    assert np.array_equal(out, want)


@pytest.mark.parametrize("kind", ["heavy", "uniform", "zeros", "random"])
@pytest.mark.parametrize("shards", [2, 3, 8])
def test_shard_composites_give_the_exact_carry(kind, shards):
    """The multi-GPU exchange without a serial hand-over: every shard's COMPOSITE (formed after
    phase 1, from the approximate carry alone) applied in order to 0 must give exactly the running
    sum np.cumsum has before the next shard — and the shards emitted with those carries must equal
    the single-array result.  All shards run on one GPU here; tests/…::test_nccl_* runs the ranks."""
    import ctypes
    import torch
    from filterpy_b200 import _lib
    from filterpy_b200.common import workloads as wl
    from filterpy_b200.distributed import shard_bounds
    from oracle import resample as ors
    N, u = 300007, 0.6180339887
    w = wl.resample_weights(N, kind, seed=13)
    want = ors.systematic_resample_c(w, u)
    csum = np.cumsum(w)
    lib = _lib.load()
    wd = torch.from_numpy(w).cuda()
    b = shard_bounds(N, shards)
    st = torch.cuda.current_stream().cuda_stream
    cbytes = int(lib.bke_resample_composite_bytes())
    allc = torch.zeros(shards * cbytes, dtype=torch.uint8, device="cuda")
    sums, args, keep = [], [], []
    for r in range(shards):
        n_loc = int(b[r + 1] - b[r])
        ws_bytes = int(lib.bke_resample_workspace_bytes(n_loc))
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
        ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
        sl = wd[int(b[r]):int(b[r + 1])]
        s = torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib.check(lib.bke_weights_sum(n_loc, sl.data_ptr(), s.data_ptr(), ws_ptr, ws_bytes, st))
        sums.append(s)
        keep.append((ws, sl))
        args.append((n_loc, ws_ptr, ws_bytes, sl))
    outs = []
    for r in range(shards):                       # phase 1 + composite: independent of every other shard's chain
        n_loc, ws_ptr, ws_bytes, sl = args[r]
        idx = torch.full((N,), -7, dtype=torch.int32, device="cuda")
        info = torch.zeros(8, dtype=torch.int32, device="cuda")
        rng_t = torch.zeros(2, dtype=torch.int64, device="cuda")
        carry_out = torch.zeros(1, dtype=torch.float64, device="cuda")
        capx = torch.stack(sums[:r]).sum().reshape(1) if r else torch.zeros(1, dtype=torch.float64, device="cuda")
        a = _lib.ResampleShardArgs()
        a.n_local, a.n_global, a.j_offset, a.capacity = n_loc, N, int(b[r]), N
        a.weights, a.u = sl.data_ptr(), u
        a.carry_approx = capx.data_ptr()
        a.indexes, a.out_range, a.carry_out = idx.data_ptr(), rng_t.data_ptr(), carry_out.data_ptr()
        a.workspace, a.workspace_bytes, a.info = ws_ptr, ws_bytes, info.data_ptr()
        a.is_last = 1 if r == shards - 1 else 0
        a.phase = 1
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))
        _lib.check(lib.bke_resample_shard_compose(ctypes.byref(a), allc.data_ptr() + r * cbytes, st))
        outs.append((a, idx, info, rng_t, carry_out, capx))
    out = np.full(N, -1, dtype=np.int32)
    covered = 0
    for r in range(shards):
        a, idx, info, rng_t, carry_out, capx = outs[r]
        carry = torch.zeros(1, dtype=torch.float64, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.bke_resample_compose_carry(r, allc.data_ptr(), carry.data_ptr(), status.data_ptr(), st))
        assert int(status.item()) == 0, (kind, shards, r)
        assert float(carry.item()) == (csum[int(b[r]) - 1] if r else 0.0), (kind, shards, r)
        a.carry_exact = carry.data_ptr()
        a.phase = 2
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))
        a.phase = 4
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))
        lo, hi = [int(v) for v in rng_t.cpu().numpy()]
        inf = info.cpu().numpy()
        assert inf[1] == 0 and inf[6] == 0, inf
        assert lo == covered
        out[lo:hi] = idx[:hi - lo].cpu().numpy()
        covered = hi
    assert covered == N
    assert np.array_equal(out, want)


def test_nccl_sharded_resample_two_ranks():
    """The NCCL path of filterpy_b200.distributed (one process per GPU, torchrun): bit-equal to the C
    oracle for both exchange methods.  Needs two visible GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for method in ("plan", "compose", "relay"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
               "--master-addr", "127.0.0.1", "--master-port", "29517",
               os.path.join(root, "scripts", "dist_resample_check.py"), "22", method]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "bit-exact: True" in r.stdout, r.stdout[-2000:]
        assert "exact on every rank: True" in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("kind", ["heavy", "zeros"])
def test_sharded_plan_single_rank_equals_oracle(kind):
    """ShardedResamplePlan with one rank (no process group): the staged C-ABI sequence (sum ->
    composite -> exact carry -> emit) must reproduce the single-array result."""
    import torch
    from filterpy_b200.common import workloads as wl
    from filterpy_b200.distributed import ShardedResamplePlan
    from oracle import resample as ors
    N, u = 500000, 0.271828
    w = wl.resample_weights(N, kind, seed=5)
    plan = ShardedResamplePlan([N])
    for _ in range(2):                                    # twice: the plan's buffers are reused
        idx, rng_t = plan.resample(torch.from_numpy(w).cuda(), u)
    lo, hi = [int(v) for v in rng_t.cpu().numpy()]
    assert (lo, hi) == (0, N) and int(plan.status.item()) == 0 and plan.info.cpu().numpy()[1] == 0
    assert np.array_equal(idx[:N].cpu().numpy(), ors.systematic_resample_c(w, u))
