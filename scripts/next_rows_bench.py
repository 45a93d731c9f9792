"""Timings for the SURVEY §8f rows (one GPU): exact cumsum, multinomial resample, particle gather,
RTS smoother.  CUDA events, median of 10 after 3 warm-ups; bytes are the algorithmic ones."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from filterpy_b200.monte_carlo import ResamplePlan, gather_particles   # noqa: E402
from filterpy_b200.kalman import KalmanFilter, IMMEstimator              # noqa: E402
from filterpy_b200.common import workloads as wl                         # noqa: E402

PEAK = 6571.6


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(case, ms, units, bytes_per_unit):
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    print(json.dumps({"case": case, "ms": round(ms, 4), "units_per_s": units / (ms * 1e-3),
                      "bytes_per_unit": bytes_per_unit, "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 4)}))


def main():
    n = 1 << 26
    w = torch.rand(n, dtype=torch.float64, device="cuda") ** 4
    w /= w.sum()
    plan = ResamplePlan(n)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    report("exact cumsum 2^26 fp64", timeit(lambda: plan.cumsum(w, out=out)), n, 16)
    U = torch.rand(n, dtype=torch.float64, device="cuda")
    idx64 = torch.empty(n, dtype=torch.int64, device="cuda")
    report("multinomial_resample 2^26 (bracket table)", timeit(lambda: plan.multinomial(w, U, out=idx64, scratch=out)), n, 24)
    report("multinomial_resample 2^26 (plain bisection)", timeit(lambda: plan.multinomial(w, U, out=idx64, scratch=out, lut=False), reps=3, warm=1), n, 24)
    assert plan.info()[1] == 0
    idx = plan.systematic(w, 0.37).clone()
    for d in [4, 16]:
        parts = torch.randn(n // 4, d, dtype=torch.float32, device="cuda")
        ii = idx[: n // 4] // 4
        dst = torch.empty_like(parts)
        report("gather_particles 2^24 x %d fp32 (systematic indexes)" % d,
               timeit(lambda: gather_particles(parts, ii, out=dst, check=False)), n // 4, 4 + 8 * d)
    del w, U, idx64, out, plan
    torch.cuda.empty_cache()
    for dtype, s in [(np.float32, 4), (np.float64, 8)]:
        N, T = 1 << 18, 32
        wk = wl.kf_bank_cv2d(4096, seed=5)
        reps = N // 4096
        kf = KalmanFilter(4, 2, n_filters=N, dtype=dtype, diagnostics=False)
        kf.x = np.tile(wk["x"], (reps, 1)); kf.P = np.tile(wk["P"], (reps, 1, 1))
        kf.F = np.tile(wk["F"], (reps, 1, 1)); kf.Q = np.tile(wk["Q"], (reps, 1, 1))
        kf.H = np.tile(wk["H"], (reps, 1, 1)); kf.R = np.tile(wk["R"], (reps, 1, 1))
        zs = torch.randn(T, N, 2, device="cuda").to(kf._dtype)
        means, covs, _, _ = kf.batch_filter(zs)
        report("rts_smoother 4/2 %s, 2^18 filters x 32 epochs" % np.dtype(dtype).name,
               timeit(lambda: kf.rts_smoother(means, covs)), N * T, (2 * 4 + 4 * 16) * s)
    # IMM: 2^20 tracks x 3 constant-velocity models (different process noise), fp32
    N, M = 1 << 20, 3
    wk = wl.kf_bank_cv2d(4096, seed=6)
    reps = N // 4096
    fs = []
    for j in range(M):
        f = KalmanFilter(4, 2, n_filters=N, dtype=np.float32)
        f.x = np.tile(wk["x"], (reps, 1)); f.P = np.tile(wk["P"], (reps, 1, 1))
        f.F = np.tile(wk["F"], (reps, 1, 1)); f.H = wk["H"][0]; f.R = wk["R"][0]
        f.Q = np.tile(wk["Q"], (reps, 1, 1)) * (10.0 ** j)
        fs.append(f)
    imm = IMMEstimator(fs, [0.5, 0.3, 0.2], np.array([[.9, .05, .05], [.1, .8, .1], [.05, .15, .8]]))
    z = torch.randn(N, 2, device="cuda")

    def imm_step():
        imm.predict(); imm.update(z)
    # per track and step: the mix reads and writes M (n + n^2), the two combined estimates read M (n + n^2)
    # and write n + n^2 each; the model filters' own traffic is on top (reported by the KF rows)
    e = 4 + 16
    report("IMM predict+update, 2^20 tracks x 3 models 4/2 fp32 (mixing traffic only)", timeit(imm_step), N,
           (2 * M * e + 2 * (M * e + e)) * 4)
    g3 = imm.capture(lambda: [imm_step() for _ in range(3)])
    report("IMM predict+update as a CUDA graph of 3 steps (per step)", timeit(g3.replay) / 3, N,
           (2 * M * e + 2 * (M * e + e)) * 4)
    ms_mix = timeit(lambda: imm._compute_state_estimate())
    report("IMM combined estimate alone", ms_mix, N, (M * e + e) * 4)


if __name__ == "__main__":
    main()
