"""First-contact check of the single-pass resample kernel (csrc/resample_fused.cu) on a GPU:
parity against the oracle over ragged sizes / weight kinds / modes, then timings at 2^26.
python scripts/rs_fused_check.py [quick]"""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.common import workloads as wl
from filterpy_b200.monte_carlo import ResamplePlan
from oracle import resample as ors

fails = 0


def report(name, ok, extra=""):
    global fails
    if not ok:
        fails += 1
    print("%s %-60s %s" % ("ok  " if ok else "FAIL", name, extra), flush=True)


def sys_case(name, w, u, offset=0):
    n = len(w)
    buf = torch.zeros(n + 2, dtype=torch.float64, device="cuda")
    wd = buf[offset:offset + n]
    wd.copy_(torch.from_numpy(w))
    plan = ResamplePlan(n)
    plan.indexes.fill_(-7)
    idx = plan.systematic(wd, u).cpu().numpy()
    info = plan.info()
    try:
        want = ors.resample_vec(w, ors.positions_systematic(n, u))
        ok = info[0] == 0 and np.array_equal(idx, want)
        bad = np.flatnonzero(idx != want)[:4] if not ok else []
    except IndexError:
        ok = info[0] > 0
        bad = []
    cl = float(plan.cumsum_last.item())
    ok = ok and cl == np.cumsum(w)[-1] and info[1] == 0
    report(name, ok, "info=%s bad=%s" % (info.tolist(), list(bad)))


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    print(torch.cuda.get_device_name(0), flush=True)
    rng = np.random.default_rng(0)
    for nw, st in (("8", "3"), ("4", "3"), ("8", "2"), ("8", "5")):
        os.environ["BKE_RS_WARPS"] = nw
        os.environ["BKE_RS_STAGES"] = st
        nw = nw + "/" + st
        for n in (1, 2, 15, 16, 17, 31, 33, 255, 2047, 4095, 4096, 4097, 8191, 8193, 65537, 300007):
            for kind in ("heavy", "uniform", "zeros", "degenerate", "dyadic"):
                w = wl.resample_weights(n, kind, seed=n % 997)
                sys_case("nw%s sys n=%d %s" % (nw, n, kind), w, 0.37454011884736254)
        w = wl.resample_weights(100003, "heavy", seed=3)
        sys_case("nw%s sys unaligned pointer" % nw, w, 0.25, offset=1)
        sys_case("nw%s sys u=0" % nw, w, 0.0)
        sys_case("nw%s sys u~1" % nw, w, 0.9999999999)
        # skewed: tiles whose output count exceeds the window, long runs
        n = 1 << 18
        w = rng.random(n); w[: n // 2] *= 1e-3; w[1000] = 50.0; w /= w.sum()
        sys_case("nw%s sys skewed (general expansion, long run)" % nw, w, 0.61)
        w = rng.integers(0, 8, n) * 2.0 ** -55 + rng.integers(0, 3, n) * 2.0 ** -20; w /= w.sum()
        sys_case("nw%s sys ties" % nw, w, 0.123)
        w = np.concatenate([np.zeros(n // 3), rng.random(n - n // 3)]); w /= w.sum()
        sys_case("nw%s sys leading zeros" % nw, w, 0.123)
        w = 10.0 ** rng.uniform(-30, 0, n); w /= w.sum()
        sys_case("nw%s sys huge range" % nw, w, 0.123)
        # unnormalised sums beyond 1 (several binades, ties at 2^-52 multiples)
        w = np.floor(rng.random(n) * 2.0 ** 40) * 2.0 ** -50
        sys_case("nw%s sys unnormalised dyadic (overflow expected)" % nw, w * 1e-3, 0.5)
        # stratified
        for n2 in (1000, 4097, 300007):
            w = wl.resample_weights(n2, "heavy", seed=7)
            U = np.random.default_rng(n2).random(n2)
            plan = ResamplePlan(n2)
            idx = plan.stratified(torch.from_numpy(w).cuda(), torch.from_numpy(U).cuda()).cpu().numpy()
            want = ors.resample_vec(w, ors.positions_stratified(n2, U))
            report("nw%s stratified n=%d" % (nw, n2), np.array_equal(idx, want) and plan.info()[1] == 0, str(plan.info().tolist()))
        # cumsum
        for n2 in (17, 5000, 300007):
            w = wl.resample_weights(n2, "zeros", seed=11)
            plan = ResamplePlan(n2)
            c = plan.cumsum(torch.from_numpy(w).cuda()).cpu().numpy()
            report("nw%s cumsum n=%d" % (nw, n2), np.array_equal(c, np.cumsum(w)), str(plan.info().tolist()))
        # fused normalisation
        for n2 in (1000, 65537, 1 << 20):
            w = rng.random(n2) ** 4 * 37.5
            plan = ResamplePlan(n2)
            wn = torch.empty(n2, dtype=torch.float64, device="cuda")
            idx, S = plan.normalized(torch.from_numpy(w).cuda(), u=0.4242, weights_out=wn)
            S = float(S.item())
            wref = w / S
            want = ors.resample_vec(wref, ors.positions_systematic(n2, 0.4242))
            ok = np.array_equal(idx.cpu().numpy(), want) and np.array_equal(wn.cpu().numpy(), wref) and abs(S / w.sum() - 1) < 1e-12
            report("nw%s normalised n=%d" % (nw, n2), ok, str(plan.info().tolist()))
        # invalid weights -> literal fallback
        w = rng.random(3000); w[100] = -0.2; w /= w.sum()
        plan = ResamplePlan(3000)
        idx = plan.systematic(torch.from_numpy(w).cuda(), 0.4).cpu().numpy()
        want = ors.resample_loop(w, ors.positions_systematic(3000, 0.4))
        report("nw%s negative weight -> fallback" % nw, plan.info()[1] == 1 and np.array_equal(idx, want), str(plan.info().tolist()))
    if quick:
        print("FAILS", fails)
        return
    # ---- timings
    N = 1 << 26
    for kind in ("heavy", "uniform"):
        w = wl.resample_weights(N, kind, seed=97)
        wd = torch.from_numpy(w).cuda()
        want = None
        for impl, nw, ctas in (("old", "8", "0"), ("new", "8", "2"), ("new", "8", "1"), ("new", "4", "4"), ("new", "4", "3"), ("new", "4", "5")):
            os.environ["BKE_RS_IMPL"] = impl; os.environ["BKE_RS_WARPS"] = nw; os.environ["BKE_RS_CTAS"] = ctas
            plan = ResamplePlan(N)
            for _ in range(3):
                plan.systematic(wd, 0.0763)
            torch.cuda.synchronize()
            reps = 10
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
            ev[0].record()
            for i in range(reps):
                plan.systematic(wd, 0.0763)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
            idx = plan.indexes.cpu().numpy()
            if want is None:
                t0 = time.time()
                want = ors.systematic_resample_c(w, 0.0763)
                print("oracle C: %.1f s" % (time.time() - t0), flush=True)
            ok = np.array_equal(idx, want)
            print("TIME %s impl=%s nw=%s ctas=%s min=%.3f ms med=%.3f ms  GB/s(12B)=%.0f exact=%s info=%s" % (
                kind, impl, nw, ctas, min(ms), sorted(ms)[len(ms) // 2], 12.0 * N / (min(ms) * 1e-3) / 1e9, ok,
                plan.info().tolist()), flush=True)
            del plan
    os.environ["BKE_RS_IMPL"] = "new"; os.environ["BKE_RS_WARPS"] = "8"; os.environ["BKE_RS_CTAS"] = "0"
    # stratified at the C5 size
    w = wl.resample_weights(N, "heavy", seed=97)
    wd = torch.from_numpy(w).cuda()
    U = np.random.default_rng(5).random(N)
    Ud = torch.from_numpy(U).cuda()
    plan = ResamplePlan(N)
    for _ in range(2):
        plan.stratified(wd, Ud)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.stratified(wd, Ud); e1.record(); torch.cuda.synchronize()
    want = ors.stratified_resample_c(w, U)
    print("TIME stratified 2^26 %.3f ms exact=%s info=%s" % (e0.elapsed_time(e1), np.array_equal(plan.indexes.cpu().numpy(), want), plan.info().tolist()))
    print("FAILS", fails)


if __name__ == "__main__":
    main()
