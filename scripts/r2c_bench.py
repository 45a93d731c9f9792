"""Third part of round 2: timings of the rows added last (residual_resample, the dim_x = 16 row-block
instances).  python scripts/r2c_bench.py  -> JSON lines."""
import json
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.monte_carlo import residual_resample_with_uniforms
from filterpy_b200.common import workloads as wl

PEAK = 6571.6


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))


def kf16(n, m, N, dtype, shared):
    rng = np.random.default_rng(n * 100 + m)

    def spd(k, cnt, scale):
        a = rng.normal(size=(cnt, k, k)).astype(np.float32)
        return (scale * (a @ np.swapaxes(a, -1, -2) / k + np.eye(k, dtype=np.float32))).astype(dtype)
    cnt = 1 if shared else N
    F = (np.eye(n) + 0.1 * rng.normal(size=(cnt, n, n))).astype(dtype)
    H = rng.normal(size=(cnt, m, n)).astype(dtype)
    Q, R, P0 = spd(n, cnt, 0.05), spd(m, cnt, 0.5), spd(n, N, 2.0)
    kf = KalmanFilter(n, m, n_filters=N, dtype=dtype, diagnostics=False)
    kf.x, kf.P = rng.normal(size=(N, n)).astype(dtype), P0
    kf.F, kf.H, kf.Q, kf.R = (F[0], H[0], Q[0], R[0]) if shared else (F, H, Q, R)
    z = torch.from_numpy(rng.normal(size=(N, m)).astype(dtype)).cuda()
    s = 4 if dtype == np.float32 else 8
    bpu = ((2 * n + 2 * n * n + m) if shared else (2 * n + 4 * n * n + m + m * n + m * m)) * s

    def step():
        kf.predict(); kf.update(z)
    ms = timeit(step, reps=10)
    gbs = N * bpu / (ms * 1e-3) / 1e9
    print(json.dumps({"case": "kf %d/%d %s %s N=%d (row-block instance)" % (n, m, np.dtype(dtype).name, "shared" if shared else "per-filter", N),
                      "ms": round(ms, 4), "filter_steps_per_s": N / (ms * 1e-3), "bytes_per_unit": bpu, "GBps": round(gbs, 1),
                      "frac_of_measured_hbm": round(gbs / PEAK, 4)}), flush=True)


def residual(N, kind):
    w = wl.resample_weights(N, kind, seed=5)
    wd = torch.from_numpy(w).cuda()
    rng = np.random.default_rng(1)
    U = rng.random(N)
    info = {}

    def run():
        _, i = residual_resample_with_uniforms(wd, lambda m: U[:m])
        info.update(i)
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"case": "residual_resample N=%d %s (wall clock incl. the host's k read-back, key upload and sweep flags)" % (N, kind),
                      "ms": round(ms, 3), "particles_per_s": N / (ms * 1e-3), "k": info["k"], "sweeps": info["sweeps"]}), flush=True)


if __name__ == "__main__":
    for (n, m) in [(16, 4), (16, 2)]:
        for dtype in (np.float32, np.float64):
            for shared in (False, True):
                kf16(n, m, 1 << 18, dtype, shared)
    for N in (1 << 16, 1 << 20, 1 << 24):
        residual(N, "heavy")
    residual(1 << 20, "uniform")
