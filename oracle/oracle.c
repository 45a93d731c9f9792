/* oracle.c — plain-C restatement of the filterpy hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (filterpy_b200/) never links it.
 *
 * Follows rlabbe/filterpy 1.4.5 @ 3b51149:
 *   oracle_systematic_resample / oracle_stratified_resample
 *        filterpy/monte_carlo/resampling.py:117-150 / :80-114
 *        (np.cumsum = strictly sequential fp64 adds, :142; two-pointer merge :143-149;
 *         positions = (u + i) / N, :139 and (U[i] + i) / N, :103)
 *   oracle_multinomial_resample
 *        filterpy/monte_carlo/resampling.py:173-176 (np.cumsum, cumulative_sum[-1] = 1., np.searchsorted side='left')
 *   oracle_kf_step_f64
 *        filterpy/kalman/kalman_filter.py:471-478 (predict) and :533-556 (update, Joseph form)
 *        S^-1 by Gauss-Jordan with partial pivoting (the reference calls np.linalg.inv ->
 *        LAPACK getrf/getri; same pivoting strategy, equal to rounding level).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no libgomp in the image, so callers thread over shards with Python threads — ctypes drops the GIL).  -ffp-contract=off keeps
 * gcc from fusing a*b+c, so every add/mul rounds as the NumPy path does.
 * Parity pinning: tests/test_oracle_golden.py checks these against tests/golden/.npz vectors
 * generated from the unmodified reference (tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---------------------------------------------------------------- resampling */
static int merge_positions(const double *cs, int64_t N, const double *pos, int32_t *idx)
{
    int64_t i = 0, j = 0;
    while (i < N) {
        if (j >= N) return 1;                 /* reference: IndexError, resampling.py:145 */
        if (pos[i] < cs[j]) { idx[i] = (int32_t)j; i++; }
        else j++;
    }
    return 0;
}

void oracle_cumsum_f64(const double *w, int64_t N, double *cs)
{
    double s = 0.0;
    for (int64_t j = 0; j < N; j++) { s = (j == 0) ? w[0] : s + w[j]; cs[j] = s; }
}

int oracle_systematic_resample(const double *w, int64_t N, double u, int32_t *idx)
{
    if (N <= 0) return 0;
    double *cs = (double *)malloc(sizeof(double) * (size_t)N);
    double *pos = (double *)malloc(sizeof(double) * (size_t)N);
    oracle_cumsum_f64(w, N, cs);
    for (int64_t i = 0; i < N; i++) pos[i] = (u + (double)i) / (double)N;
    int rc = merge_positions(cs, N, pos, idx);
    free(cs); free(pos);
    return rc;
}

int oracle_stratified_resample(const double *w, int64_t N, const double *U, int32_t *idx)
{
    if (N <= 0) return 0;
    double *cs = (double *)malloc(sizeof(double) * (size_t)N);
    double *pos = (double *)malloc(sizeof(double) * (size_t)N);
    oracle_cumsum_f64(w, N, cs);
    for (int64_t i = 0; i < N; i++) pos[i] = (U[i] + (double)i) / (double)N;
    int rc = merge_positions(cs, N, pos, idx);
    free(cs); free(pos);
    return rc;
}

void oracle_multinomial_resample(const double *w, int64_t N, const double *U, int64_t nk, double *cs, int64_t *idx)
{
    if (N <= 0) return;
    oracle_cumsum_f64(w, N, cs);
    cs[N - 1] = 1.0;                           /* resampling.py:174 */
    for (int64_t q = 0; q < nk; q++) {         /* np.searchsorted(cs, U[q]), side='left': #{j : cs[j] < key} for sorted cs */
        int64_t lo = 0, hi = N;
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (cs[mid] < U[q]) lo = mid + 1; else hi = mid;
        }
        idx[q] = lo;
    }
}

/* ---------------------------------------------------------------- linear KF, fp64 */
#define MAXD 32

static void matmul(const double *A, const double *B, double *C, int r, int k, int c, int tb)
{   /* C[r,c] = A[r,k] * (tb ? B[c,k]' : B[k,c]) ; accumulation order k ascending, like a
       naive dot (BLAS may block differently; differences are at rounding level). */
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) {
            double s = 0.0;
            for (int q = 0; q < k; q++) s += A[i * k + q] * (tb ? B[j * k + q] : B[q * c + j]);
            C[i * c + j] = s;
        }
}

static int inv_gj(const double *S, double *SI, int m)
{
    double a[MAXD * MAXD];
    memcpy(a, S, sizeof(double) * m * m);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) SI[i * m + j] = (i == j);
    for (int c = 0; c < m; c++) {
        int p = c; double best = fabs(a[c * m + c]);
        for (int r = c + 1; r < m; r++) if (fabs(a[r * m + c]) > best) { best = fabs(a[r * m + c]); p = r; }
        if (best == 0.0) return 1;
        if (p != c) for (int j = 0; j < m; j++) {
            double t = a[c * m + j]; a[c * m + j] = a[p * m + j]; a[p * m + j] = t;
            t = SI[c * m + j]; SI[c * m + j] = SI[p * m + j]; SI[p * m + j] = t;
        }
        double d = 1.0 / a[c * m + c];
        for (int j = 0; j < m; j++) { a[c * m + j] *= d; SI[c * m + j] *= d; }
        for (int r = 0; r < m; r++) if (r != c) {
            double f = a[r * m + c];
            if (f != 0.0) for (int j = 0; j < m; j++) { a[r * m + j] -= f * a[c * m + j]; SI[r * m + j] -= f * SI[c * m + j]; }
        }
    }
    return 0;
}

/* One predict+update of N independent filters.  Strides are in elements per filter (0 = shared).
 * valid may be NULL.  x/P are updated in place.  Returns the number of singular-S filters. */
int oracle_kf_step_f64(int64_t N, int n, int m, double *x, double *P,
                       const double *F, int64_t sF, const double *H, int64_t sH,
                       const double *Q, int64_t sQ, const double *R, int64_t sR,
                       const double *z, const uint8_t *valid, double alpha_sq, int nthreads)
{
    int bad = 0;
    if (n > MAXD || m > MAXD) return -1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) reduction(+:bad) schedule(static)
#endif
    for (int64_t f = 0; f < N; f++) {
        const double *Ff = F + f * sF, *Hf = H + f * sH, *Qf = Q + f * sQ, *Rf = R + f * sR;
        double *xf = x + f * n, *Pf = P + f * (int64_t)n * n;
        double xp[MAXD], FP[MAXD * MAXD], Pp[MAXD * MAXD];
        /* predict: kalman_filter.py:475,478 */
        matmul(Ff, xf, xp, n, n, 1, 0);
        matmul(Ff, Pf, FP, n, n, n, 0);
        matmul(FP, Ff, Pp, n, n, n, 1);
        for (int e = 0; e < n * n; e++) Pp[e] = alpha_sq * Pp[e] + Qf[e];
        if (valid && !valid[f]) {           /* kalman_filter.py:515-520 */
            memcpy(xf, xp, sizeof(double) * n); memcpy(Pf, Pp, sizeof(double) * n * n);
            continue;
        }
        /* update: kalman_filter.py:533-556 */
        double y[MAXD], PHT[MAXD * MAXD], S[MAXD * MAXD], SI[MAXD * MAXD], K[MAXD * MAXD];
        double IKH[MAXD * MAXD], T[MAXD * MAXD], KR[MAXD * MAXD], A1[MAXD * MAXD], A2[MAXD * MAXD];
        matmul(Hf, xp, y, m, n, 1, 0);
        for (int a = 0; a < m; a++) y[a] = z[f * m + a] - y[a];
        matmul(Pp, Hf, PHT, n, n, m, 1);
        matmul(Hf, PHT, S, m, n, m, 0);
        for (int e = 0; e < m * m; e++) S[e] += Rf[e];
        if (inv_gj(S, SI, m)) { bad++; continue; }
        matmul(PHT, SI, K, n, m, m, 0);
        double Ky[MAXD];
        matmul(K, y, Ky, n, m, 1, 0);
        for (int i = 0; i < n; i++) xf[i] = xp[i] + Ky[i];
        matmul(K, Hf, IKH, n, m, n, 0);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) IKH[i * n + j] = (i == j) - IKH[i * n + j];
        matmul(IKH, Pp, T, n, n, n, 0);
        matmul(T, IKH, A1, n, n, n, 1);
        matmul(K, Rf, KR, n, m, m, 0);
        matmul(KR, K, A2, n, m, n, 1);
        for (int e = 0; e < n * n; e++) Pf[e] = A1[e] + A2[e];
    }
    return bad;
}
