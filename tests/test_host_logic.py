"""Host-side logic that needs no GPU: shape rules, weights, workload generators, sharding."""
import os
import subprocess
import sys

import numpy as np
import pytest

from filterpy_b200.common.helpers import reshape_z
from filterpy_b200.common import workloads as wl
from filterpy_b200.kalman.sigma_points import MerweScaledSigmaPoints
from filterpy_b200 import distributed as bd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reshape_z_contract():
    # filterpy/common/helpers.py:324-342 — accepted and rejected shapes (test_kf.py:529-654)
    assert reshape_z(3.0, 1, 2).shape == (1, 1)
    assert reshape_z([1., 2.], 2, 2).shape == (2, 1)
    assert reshape_z([[1., 2.]], 2, 1).shape == (2,)
    assert reshape_z([[1.], [2.]], 2, 1).shape == (2,)
    assert reshape_z(3.0, 1, 0) == 3.0
    for bad in ([1., 2., 3.], [[1., 2.], [3., 4.], [5., 6.]]):
        with pytest.raises(ValueError):
            reshape_z(bad, 2, 2)


def test_merwe_weights_golden(golden):
    g = golden("ukf_sigma")
    p = MerweScaledSigmaPoints(6, float(g["alpha"]), float(g["beta"]), float(g["kappa"]))
    np.testing.assert_allclose(p.Wm, g["Wm"], rtol=1e-15)
    np.testing.assert_allclose(p.Wc, g["Wc"], rtol=1e-15)
    assert p.num_sigmas() == 13
    p4 = MerweScaledSigmaPoints(4, .5, 2, 0)
    np.testing.assert_allclose(p4.Wm, g["Wm4"]); np.testing.assert_allclose(p4.Wc, g["Wc4"])
    assert abs(p4.Wm.sum() - 1) < 1e-12       # test_ukf.py:102-109
    with pytest.raises(NotImplementedError):
        MerweScaledSigmaPoints(4, .5, 2, 0, sqrt_method=np.linalg.cholesky)


def test_workloads_are_seeded_and_well_posed():
    a = wl.kf_bank_cv2d(100, seed=5, steps=2); b = wl.kf_bank_cv2d(100, seed=5, steps=2)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["F"].shape == (100, 4, 4) and a["H"].shape == (100, 2, 4) and a["zs"].shape == (2, 100, 2)
    c = wl.kf_bank_ca3d(10, steps=1)
    assert c["F"].shape == (10, 9, 9) and c["H"][0, 1, 3] == 1 and c["H"][0].sum() == 3
    # Q blocks are symmetric PSD
    assert np.allclose(c["Q"], np.swapaxes(c["Q"], 1, 2)) and np.linalg.eigvalsh(c["Q"]).min() > -1e-12
    for kind in ["heavy", "uniform", "random", "zeros", "degenerate"]:
        w = wl.resample_weights(1000, kind)
        assert abs(w.sum() - 1) < 1e-12 and w.min() >= 0
    d = wl.resample_weights(1000, "dyadic")
    assert np.all(d * 2.0 ** 52 == np.floor(d * 2.0 ** 52))


def test_shard_bounds():
    b = bd.shard_bounds(10, 4)
    assert b.tolist() == [0, 3, 6, 8, 10]
    assert bd.shard_bounds(1 << 20, 8)[-1] == 1 << 20
    x = np.arange(10)
    parts = [bd.shard_of(x, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), x)


_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from filterpy_b200 import distributed as bd
from filterpy_b200.common import workloads as wl
from oracle import kf as okf, resample as ors
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# (1) bank sharding: every rank steps its slice with the oracle; the gathered result equals the full bank
N = 37
w = wl.kf_bank_cv2d(N, seed=3, steps=1)
sl = {k: bd.shard_of(v, rank, world, axis=1 if k == "zs" else 0) for k, v in w.items()}
o = okf.kf_step_bank(sl["x"], sl["P"], sl["zs"][0], sl["F"], sl["H"], sl["Q"], sl["R"])
parts = [None] * world
dist.all_gather_object(parts, o["x"])
full = okf.kf_step_bank(w["x"], w["P"], w["zs"][0], w["F"], w["H"], w["Q"], w["R"])["x"]
assert np.array_equal(np.concatenate(parts), full)
# (2) particle-weight sum all-reduce + exclusive prefix of shard sums
wts = wl.resample_weights(1001, "heavy", seed=1)
mine = bd.shard_of(wts, rank, world)
s = torch.tensor([mine.sum()], dtype=torch.float64)
tot = bd.all_reduce_sum(s.clone())
assert abs(float(tot) - wts.sum()) < 1e-12
pre = bd.exclusive_prefix(s)
b = bd.shard_bounds(len(wts), world)
assert abs(float(pre) - wts[:b[rank]].sum()) < 1e-12
# (3) sharded resampling semantics: rank r owns outputs [cnt(c_start), cnt(c_end)); concatenation == reference
u = 0.3
idx = ors.systematic_resample_vec(wts, u)
c = np.cumsum(wts)
pos = ors.positions_systematic(len(wts), u)
lo = 0 if rank == 0 else int(np.searchsorted(pos, c[b[rank] - 1], side="left"))
hi = int(np.searchsorted(pos, c[b[rank + 1] - 1], side="left"))
local = idx[lo:hi]
assert local.size == 0 or (local.min() >= b[rank] and local.max() < b[rank + 1])
cnts = [None] * world
dist.all_gather_object(cnts, (lo, hi))
assert cnts[0][0] == 0 and all(cnts[i][1] == cnts[i + 1][0] for i in range(world - 1)) and cnts[-1][1] == len(wts)
# (4) re-sharding after the resample: local gather of the rank's own rows + contiguous slice exchange == particles[idx]
particles = np.random.default_rng(5).normal(size=(len(wts), 3))
mine_p = bd.shard_of(particles, rank, world)
rows = torch.from_numpy(mine_p[local - b[rank]])                 # the local gather (bke_gather_rows on the GPU)
sends, recvs = bd.exchange_plan(cnts, b, rank)
assert sum(h - l for _, l, h in sends) == hi - lo and sum(h - l for _, l, h in recvs) == b[rank + 1] - b[rank]
out = torch.empty(int(b[rank + 1] - b[rank]), 3, dtype=torch.float64)
bd.exchange_rows(rows, out, sends, recvs)
assert np.array_equal(out.numpy(), particles[idx][b[rank]:b[rank + 1]])
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank %d ok\n" % rank)
sys.stdout.flush()
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rank%d.ok" % rank), "w").write("ok")
'''


def test_gloo_world_size_2(tmp_path):
    """The N>1 host path on CPU: 2 ranks over gloo (sharding, weight-sum all-reduce, shard prefix,
    output-range ownership of a sharded resample)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists(), r.stdout[-2000:]


def test_exchange_plan_covers_every_output_exactly_once():
    """After a sharded resample rank r owns the output positions [o_r, o_r+1); the re-sharding plan
    must move every position to the rank that owns it under the even split, exactly once."""
    from filterpy_b200 import distributed as bd
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for n in (world, 17, 1000, 4097):
            b = bd.shard_bounds(n, world)
            cuts = np.sort(rng.integers(0, n + 1, size=world - 1)) if world > 1 else np.zeros(0, int)
            edges = np.concatenate([[0], cuts, [n]])
            out_ranges = [(int(edges[r]), int(edges[r + 1])) for r in range(world)]     # includes empty ranks
            got = np.zeros(n, int)
            for r in range(world):
                sends, recvs = bd.exchange_plan(out_ranges, b, r)
                for dst, lo, hi in sends:
                    g_lo, g_hi = out_ranges[r][0] + lo, out_ranges[r][0] + hi
                    assert b[dst] <= g_lo and g_hi <= b[dst + 1]
                    got[g_lo:g_hi] += 1
                assert sum(h - l for _, l, h in recvs) == b[r + 1] - b[r]
                for src, lo, hi in recvs:
                    g_lo, g_hi = b[r] + lo, b[r] + hi
                    assert out_ranges[src][0] <= g_lo and g_hi <= out_ranges[src][1]
            assert (got == 1).all()


# ------------------------------------------------------------------ user-supplied UKF models (NVRTC half, no GPU)
def test_user_ukf_model_text_compiles_for_sm100a():
    """The program text bke_ukf_model_compile builds around DeviceFx / DeviceHx sources compiles with NVRTC for
    sm_100a (both element types, user fx + user hx and user fx + built-in hx); a broken source comes back
    with the compiler's message."""
    from filterpy_b200 import _lib
    from filterpy_b200.common import workloads as wl
    lib = _lib.load()
    inc = _lib.kernel_include_dirs().encode()
    src = (wl.CT_FX_SOURCE + wl.OFFSET_RB_HX_SOURCE).encode()
    for dt in (_lib.BKE_F32, _lib.BKE_F64):
        assert lib.bke_debug_ukf_model_cubin_bytes(4, 2, dt, _lib.BKE_FX_USER, _lib.BKE_HX_USER, src, inc) > 10000, lib.bke_last_error()
    assert lib.bke_debug_ukf_model_cubin_bytes(4, 2, _lib.BKE_F64, _lib.BKE_FX_USER, _lib.BKE_HX_LINEAR, wl.CT_FX_SOURCE.encode(), inc) > 10000
    bad = b"__device__ void fx(const real *x, real *out, real dt, const real *args) { out[0] = undefined_symbol; }"
    assert lib.bke_debug_ukf_model_cubin_bytes(4, 2, _lib.BKE_F64, _lib.BKE_FX_USER, _lib.BKE_HX_LINEAR, bad, inc) == 0
    msg = lib.bke_last_error().decode()
    assert "undefined_symbol" in msg and "user_model.cu" in msg
    # neither function user-supplied / unsupported built-in partner: refused before NVRTC runs
    assert lib.bke_debug_ukf_model_cubin_bytes(4, 2, _lib.BKE_F64, _lib.BKE_FX_LINEAR, _lib.BKE_HX_LINEAR, src, inc) == 0
    assert lib.bke_debug_ukf_model_cubin_bytes(4, 2, _lib.BKE_F64, _lib.BKE_FX_USER, _lib.BKE_HX_RANGE_BEARING, src, inc) == 0


def test_device_model_argument_packing():
    import torch
    from filterpy_b200.kalman import DeviceFx
    m = DeviceFx("", arg_names=("a", "b"), a=1.5)
    with pytest.raises(TypeError):
        m.pack({}, 4, torch.float64, "cpu")                      # b has no value yet
    t, stride = m.pack({"b": 2.0}, 4, torch.float64, "cpu")
    assert stride == 0 and t.tolist() == [1.5, 2.0]
    t, stride = m.pack({"a": np.arange(4.0)}, 4, torch.float32, "cpu")
    assert stride == 2 and t.shape == (4, 2) and t[:, 0].tolist() == [0, 1, 2, 3] and t[:, 1].tolist() == [2.0] * 4
    with pytest.raises(TypeError):
        m.pack({"c": 1.0}, 4, torch.float64, "cpu")
    with pytest.raises(ValueError):
        m.pack({"a": np.arange(3.0)}, 4, torch.float64, "cpu")
    with pytest.raises(TypeError):
        DeviceFx("", arg_names=("a",), z=1)
