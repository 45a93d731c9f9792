// mix.cu — bank-level model mixing for multiple-model estimators: a bank of N tracks, each
// followed by the same M filters (one KalmanFilter bank per model).  The reference loops over a
// Python list of filter objects for ONE track (filterpy/kalman/IMM.py:160-249,
// filterpy/kalman/mmae.py:140-210); here every step is one launch over all tracks.
//
//   bke_mm_probabilities   IMM.py:178-184 + :239-247  (mu = cbar * L; normalise; cbar = mu M;
//                                                      omega[i,j] = M[i,j] mu[i] / cbar[j])
//                          mmae.py:180-184            (p *= L; p /= sum(p))
//                          with L = max(exp(log_likelihood), DBL_MIN)  (kalman_filter.py:1213-1223)
//   bke_mm_mix             IMM.py:201-213   mixed initial conditions x0_i, P0_i for every model i
//   bke_mm_estimate        IMM.py:228-237   x = sum mu_j x_j, P = sum mu_j ((x_j - x)(x_j - x)' + P_j)
//                          mmae.py:186-201  (with the reference's element-wise zip, see BKE_MM_MMAE)
//
// Work split: one thread per (track, element of [x | P]) so that consecutive threads touch
// consecutive addresses of every model's AoS arrays; the few per-track scalars (mu, omega) are
// re-read by the n + n^2 threads of a track from L1.  All HBM-bound: per track the mix reads and
// writes M (n + n^2) scalars.
#include <float.h>
#include "bke_internal.cuh"

namespace bke {
namespace {

template <typename T>
struct MixP {
    int64_t N;
    int n, M;
    unsigned flags;
    const T *x[BKE_MM_MAX_MODELS], *P[BKE_MM_MAX_MODELS];
    T *xo[BKE_MM_MAX_MODELS], *Po[BKE_MM_MAX_MODELS];
    const T *ll[BKE_MM_MAX_MODELS];
    const double *w;             // omega[N,M,M] (mix) or mu[N,M] (estimate)
    int64_t sw;                  // per-track stride of w (0 = shared)
    double *mu, *cbar, *omega;   // probabilities kernel
    const double *trans;         // M[M,M]
};

// mixed initial conditions (IMM.py:201-213): for every target model i
//   x0_i = sum_j omega[j,i] x_j ;  P0_i = sum_j omega[j,i] ((x_j - x0_i)(x_j - x0_i)' + P_j)
template <typename T>
__global__ void __launch_bounds__(256) k_mm_mix(MixP<T> p)
{
    const int n = p.n, M = p.M, E = n + n * n;
    const int64_t total = p.N * E;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = g / E;
        const int e = (int)(g - t * E);
        const double *om = p.w + t * p.sw;
        if (e < n) {
            double xj[BKE_MM_MAX_MODELS];
            for (int j = 0; j < M; j++) xj[j] = (double)p.x[j][t * n + e];
            for (int i = 0; i < M; i++) {
                double s = 0.0;
                for (int j = 0; j < M; j++) s += xj[j] * om[j * M + i];
                p.xo[i][t * n + e] = (T)s;
            }
        } else {
            const int rc = e - n, r = rc / n, c = rc - r * n;
            double xr[BKE_MM_MAX_MODELS], xc[BKE_MM_MAX_MODELS], Pj[BKE_MM_MAX_MODELS];
            for (int j = 0; j < M; j++) {
                xr[j] = (double)p.x[j][t * n + r];
                xc[j] = (double)p.x[j][t * n + c];
                Pj[j] = (double)p.P[j][t * n * n + rc];
            }
            for (int i = 0; i < M; i++) {
                double mr = 0.0, mc = 0.0;
                for (int j = 0; j < M; j++) { mr += xr[j] * om[j * M + i]; mc += xc[j] * om[j * M + i]; }
                if (sizeof(T) == 4) { mr = (double)(T)mr; mc = (double)(T)mc; }     // the mixed mean as it is stored
                double s = 0.0;
                for (int j = 0; j < M; j++) s += om[j * M + i] * ((xr[j] - mr) * (xc[j] - mc) + Pj[j]);
                p.Po[i][t * n * n + rc] = (T)s;
            }
        }
    }
}

// combined estimate (IMM.py:228-237; mmae.py:186-201 with BKE_MM_MMAE)
template <typename T>
__global__ void __launch_bounds__(256) k_mm_estimate(MixP<T> p)
{
    const int n = p.n, M = p.M, E = n + n * n;
    const bool mmae = p.flags & BKE_MM_MMAE;
    const int64_t total = p.N * E;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = g / E;
        const int e = (int)(g - t * E);
        const double *mu = p.w + t * p.sw;
        if (e < n) {
            double s = 0.0;
            for (int j = 0; j < M; j++) s += (double)p.x[j][t * n + e] * mu[j];
            p.xo[0][t * n + e] = (T)s;
        } else {
            const int rc = e - n, r = rc / n, c = rc - r * n;
            double s = 0.0;
            if (!mmae) {
                double mr = 0.0, mc = 0.0;
                for (int j = 0; j < M; j++) { mr += (double)p.x[j][t * n + r] * mu[j]; mc += (double)p.x[j][t * n + c] * mu[j]; }
                if (sizeof(T) == 4) { mr = (double)(T)mr; mc = (double)(T)mc; }
                for (int j = 0; j < M; j++)
                    s += mu[j] * (((double)p.x[j][t * n + r] - mr) * ((double)p.x[j][t * n + c] - mc) + (double)p.P[j][t * n * n + rc]);
            } else {
                // mmae.py:197-199 zips the COMPONENTS of the mixed x with the filters: term j uses
                // y = f_j.x - x[j] (a scalar), and only min(dim_x, M) terms exist
                const int terms = M < n ? M : n;
                for (int j = 0; j < terms; j++) {
                    double mj = 0.0;
                    for (int q = 0; q < M; q++) mj += (double)p.x[q][t * n + j] * mu[q];
                    if (sizeof(T) == 4) mj = (double)(T)mj;
                    s += mu[j] * (((double)p.x[j][t * n + r] - mj) * ((double)p.x[j][t * n + c] - mj) + (double)p.P[j][t * n * n + rc]);
                }
            }
            p.Po[0][t * n * n + rc] = (T)s;
        }
    }
}

// mode probabilities, one thread per track
template <typename T>
__global__ void __launch_bounds__(256) k_mm_probabilities(MixP<T> p)
{
    const int M = p.M;
    const bool mmae = p.flags & BKE_MM_MMAE, from_mu = p.flags & BKE_MM_FROM_MU;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.N; t += (int64_t)gridDim.x * blockDim.x) {
        double mu[BKE_MM_MAX_MODELS];
        if (from_mu) {
            for (int j = 0; j < M; j++) mu[j] = p.mu[t * M + j];
        } else {
            double sum = 0.0;
            for (int j = 0; j < M; j++) {
                double L = exp((double)p.ll[j][t]);
                if (L == 0.0) L = DBL_MIN;                       // kalman_filter.py:1221-1222
                const double prior = mmae ? p.mu[t * M + j] : p.cbar[t * M + j];
                mu[j] = prior * L;
                sum += mu[j];
            }
            for (int j = 0; j < M; j++) { mu[j] /= sum; p.mu[t * M + j] = mu[j]; }
        }
        if (mmae) continue;
        double cb[BKE_MM_MAX_MODELS];
        for (int j = 0; j < M; j++) {
            double s = 0.0;
            for (int i = 0; i < M; i++) s += mu[i] * p.trans[i * M + j];
            cb[j] = s;
            p.cbar[t * M + j] = s;
        }
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) p.omega[(t * M + i) * M + j] = (p.trans[i * M + j] * mu[i]) / cb[j];
    }
}

unsigned grid_for(int64_t work)
{
    int64_t b = (work + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

template <typename T>
int launch_t(const bke_mm_args &a, int op, cudaStream_t s)
{
    MixP<T> p;
    p.N = a.n_tracks; p.n = a.dim_x; p.M = a.n_models; p.flags = a.flags;
    for (int j = 0; j < BKE_MM_MAX_MODELS; j++) {
        p.x[j] = (const T *)a.x[j]; p.P[j] = (const T *)a.P[j];
        p.xo[j] = (T *)a.x_out[j]; p.Po[j] = (T *)a.P_out[j];
        p.ll[j] = (const T *)a.log_likelihood[j];
    }
    p.mu = a.mu; p.cbar = a.cbar; p.omega = a.omega; p.trans = a.trans;
    const int64_t E = a.dim_x + (int64_t)a.dim_x * a.dim_x;
    if (op == 0) {
        k_mm_probabilities<T><<<grid_for(p.N), 256, 0, s>>>(p);
    } else if (op == 1) {
        p.w = a.omega; p.sw = a.weights_stride;
        k_mm_mix<T><<<grid_for(p.N * E), 256, 0, s>>>(p);
    } else {
        p.w = a.mu; p.sw = a.weights_stride;
        k_mm_estimate<T><<<grid_for(p.N * E), 256, 0, s>>>(p);
    }
    return check_cuda(cudaGetLastError(), "mm launch");
}

int validate(const bke_mm_args *args, int op)
{
    if (!args) { set_error("args is NULL"); return BKE_ERR_BAD_ARG; }
    const bke_mm_args &a = *args;
    if (a.n_tracks < 0) { set_error("n_tracks < 0"); return BKE_ERR_BAD_ARG; }
    if (a.n_models < 1 || a.n_models > BKE_MM_MAX_MODELS) { set_error("n_models must be in [1, %d]", BKE_MM_MAX_MODELS); return BKE_ERR_UNSUPPORTED; }
    if (a.dtype != BKE_F32 && a.dtype != BKE_F64) { set_error("bad dtype"); return BKE_ERR_BAD_ARG; }
    if (op != 0 && (a.dim_x < 1 || a.dim_x > 64)) { set_error("dim_x must be in [1, 64]"); return BKE_ERR_BAD_ARG; }
    if (a.n_tracks == 0) return BKE_OK;
    const int M = a.n_models;
    if (op == 0) {
        if (!(a.flags & BKE_MM_FROM_MU))
            for (int j = 0; j < M; j++) if (!a.log_likelihood[j]) { set_error("log_likelihood[%d] is NULL", j); return BKE_ERR_BAD_ARG; }
        if (!a.mu) { set_error("mu is NULL"); return BKE_ERR_BAD_ARG; }
        if (!(a.flags & BKE_MM_MMAE) && (!a.cbar || !a.omega || !a.trans)) { set_error("cbar, omega and trans are required"); return BKE_ERR_BAD_ARG; }
    } else {
        for (int j = 0; j < M; j++) if (!a.x[j] || !a.P[j]) { set_error("x[%d] / P[%d] is NULL", j, j); return BKE_ERR_BAD_ARG; }
        const int outs = op == 1 ? M : 1;
        for (int j = 0; j < outs; j++) {
            if (!a.x_out[j] || !a.P_out[j]) { set_error("x_out[%d] / P_out[%d] is NULL", j, j); return BKE_ERR_BAD_ARG; }
            for (int q = 0; q < M; q++)
                if (a.x_out[j] == a.x[q] || a.P_out[j] == a.P[q]) { set_error("outputs must not alias the inputs"); return BKE_ERR_BAD_ARG; }
        }
        if (op == 1 && !a.omega) { set_error("omega is NULL"); return BKE_ERR_BAD_ARG; }
        if (op == 2 && !a.mu) { set_error("mu is NULL"); return BKE_ERR_BAD_ARG; }
        if (a.weights_stride < 0) { set_error("negative stride"); return BKE_ERR_BAD_ARG; }
    }
    if (bke_device_count() <= 0) { set_error("no CUDA device"); return BKE_ERR_CUDA; }
    return -1;      // go
}

int run(const bke_mm_args *args, int op, void *stream)
{
    const int v = validate(args, op);
    if (v >= 0) return v;
    return args->dtype == BKE_F32 ? launch_t<float>(*args, op, (cudaStream_t)stream)
                                  : launch_t<double>(*args, op, (cudaStream_t)stream);
}

}  // namespace
}  // namespace bke

extern "C" {
int bke_mm_probabilities(const bke_mm_args *args, void *stream) { return bke::run(args, 0, stream); }
int bke_mm_mix(const bke_mm_args *args, void *stream) { return bke::run(args, 1, stream); }
int bke_mm_estimate(const bke_mm_args *args, void *stream) { return bke::run(args, 2, stream); }
}
