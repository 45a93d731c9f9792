"""Secondary configurations (BASELINE configs 2-shared, 3, 4 and batch_filter): time on one GPU and
report achieved GB/s against the algorithmic bytes of DESIGN.md.  python scripts/variants_bench.py"""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.kalman import (KalmanFilter, UnscentedKalmanFilter, MerweScaledSigmaPoints, ConstVelFx,
                                  RangeAzElHx, LinearFx, LinearHx)
from filterpy_b200.common import workloads as wl

PEAK = 6571.6


def timeit(fn, reps=20, warm=3):
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


def report(name, ms, units, bytes_per_unit):
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    print(json.dumps({"case": name, "ms": round(ms, 4), "units_per_s": units / (ms * 1e-3), "bytes_per_unit": bytes_per_unit,
                      "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 4)}), flush=True)


def kf_case(name, n, m, N, dtype, gen, shared=False, diagnostics=False):
    w = gen(N, steps=1)
    kf = KalmanFilter(n, m, n_filters=N, dtype=dtype, diagnostics=diagnostics)
    kf.x = w["x"]; kf.P = w["P"]
    for k in "FHQR":
        setattr(kf, k, w[k][0] if shared else w[k])
    z = torch.from_numpy(w["zs"][0].astype(dtype)).cuda()
    s = 4 if dtype == np.float32 else 8
    bpu = ((2 * n + 2 * n * n + m) if shared else (2 * n + 4 * n * n + m + m * n + m * m)) * s

    def step():
        kf.predict(); kf.update(z)
    report(name, timeit(step), N, bpu)


def main():
    kf_case("C2 kf 4/2 f32 per-filter models 2^20", 4, 2, 1 << 20, np.float32, lambda N, steps: wl.kf_bank_cv2d(N, steps=steps))
    kf_case("C2 kf 4/2 f32 shared models 2^20", 4, 2, 1 << 20, np.float32, lambda N, steps: wl.kf_bank_cv2d(N, steps=steps), shared=True)
    kf_case("C2 kf 4/2 f32 per-filter + diagnostics", 4, 2, 1 << 20, np.float32, lambda N, steps: wl.kf_bank_cv2d(N, steps=steps), diagnostics=True)
    kf_case("kf 4/2 f64 (register tile, direct loads) 2^20", 4, 2, 1 << 20, np.float64, lambda N, steps: wl.kf_bank_cv2d(N, steps=steps))
    kf_case("C3 kf 9/3 f64 (row-block kernel) 1.25M", 9, 3, 1250000, np.float64, lambda N, steps: wl.kf_bank_ca3d(N, steps=steps))
    kf_case("C3 kf 9/3 f64 shared models 1.25M", 9, 3, 1250000, np.float64, lambda N, steps: wl.kf_bank_ca3d(N, steps=steps), shared=True)
    kf_case("C3 kf 9/3 f64 + all diagnostics outputs 1.25M", 9, 3, 1250000, np.float64, lambda N, steps: wl.kf_bank_ca3d(N, steps=steps), diagnostics=True)
    kf_case("kf 9/3 f32 (row-block kernel, 8 filters per warp) 1.25M", 9, 3, 1250000, np.float32, lambda N, steps: wl.kf_bank_ca3d(N, steps=steps))
    # 6/3 (3-D constant velocity with a linear position sensor): fp32 -> direct kernel, fp64 -> row-block kernel
    def cv3d(N, steps):
        u = wl.ukf_bank_cv3d(min(N, 1 << 14), steps=steps, linear_hx=True)
        r = max(1, N // u["x"].shape[0])
        tile = lambda a, lead: np.tile(a, (r,) + (1,) * (a.ndim - 1)) if lead else a     # noqa: E731
        return {"x": tile(u["x"], 1), "P": tile(u["P"], 1), "F": np.tile(u["F"], (N, 1, 1)), "H": np.tile(u["H"], (N, 1, 1)),
                "Q": tile(u["Q"], 1), "R": tile(u["R"], 1), "zs": np.tile(u["zs"], (1, r, 1))}
    kf_case("kf 6/3 f32 per-filter models 2^19", 6, 3, 1 << 19, np.float32, cv3d)
    kf_case("kf 6/3 f64 per-filter models 2^19", 6, 3, 1 << 19, np.float64, cv3d)
    # batch_filter: T epochs inside one kernel
    N, T = 1 << 18, 32
    w = wl.kf_bank_cv2d(N, steps=T, dtype=np.float32)
    kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=False)
    for k in "xPFHQR":
        setattr(kf, k, w[k])
    zs = torch.from_numpy(w["zs"]).cuda()
    report("batch_filter 4/2 f32 2^18 x 32 epochs (per filter-step)", timeit(lambda: kf.batch_filter(zs), reps=5), N * T, (2 + 2 * 4 + 2 * 16) * 4)
    w = wl.kf_bank_cv2d(N, steps=T, dtype=np.float64)
    kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float64, diagnostics=False)
    for k in "xPFHQR":
        setattr(kf, k, w[k])
    zs = torch.from_numpy(w["zs"]).cuda()
    report("batch_filter 4/2 f64 2^18 x 32 epochs (per filter-step)", timeit(lambda: kf.batch_filter(zs), reps=5), N * T, (2 + 2 * 4 + 2 * 16) * 8)
    # UKF (config C4)
    for dtype in (np.float64, np.float32):
        N = 1 << 18
        uw = wl.ukf_bank_cv3d(N, steps=1)
        u = UnscentedKalmanFilter(6, 3, 0.1, RangeAzElHx(), ConstVelFx(), MerweScaledSigmaPoints(6, .5, 2., 0.),
                                  n_filters=N, dtype=dtype, diagnostics=False)
        u.x = uw["x"]; u.P = uw["P"]; u.Q = uw["Q"]; u.R = uw["R"]
        z = torch.from_numpy(uw["zs"][0].astype(dtype)).cuda()
        s = 4 if dtype == np.float32 else 8

        def ustep():
            u.predict(); u.update(z)
        report("C4 ukf 6/3 %s cv + range/az/el 2^18" % ("f64" if dtype == np.float64 else "f32"), timeit(ustep), N,
               (2 * 6 + 3 * 36 + 3 + 9) * s)
        ul = UnscentedKalmanFilter(6, 3, 0.1, LinearHx(uw["H"]), LinearFx(uw["F"]), MerweScaledSigmaPoints(6, .5, 2., 0.),
                                   n_filters=N, dtype=dtype, diagnostics=False)
        ul.x = uw["x"]; ul.P = uw["P"]; ul.Q = uw["Q"]; ul.R = uw["R"]

        def lstep():
            ul.predict(); ul.update(z)
        report("ukf 6/3 %s linear fx/hx 2^18" % ("f64" if dtype == np.float64 else "f32"), timeit(lstep), N,
               (2 * 6 + 3 * 36 + 3 + 9) * s)


if __name__ == "__main__":
    main()
