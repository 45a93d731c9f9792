// kf_regtile.cuh — one filter's predict/update held entirely in registers (thread-per-filter
// tile used by the specialised kernels in kf_fast.cu).  Fully unrolled for compile-time (N, M);
// every array index is a constant so nothing touches local memory.
//
// Arithmetic (filterpy/kalman/kalman_filter.py, reference @ 3b51149):
//   predict :471-478   x = F x ;  P = alpha_sq * (F P) F' + Q
//   update  :533-556   y = z - H x ; PHT = P H' ; S = H PHT + R ; SI = S^-1 ; K = PHT SI ;
//                      x = x + K y ; P = (I-KH) P (I-KH)' + (K R) K'
#pragma once
#include "bke_internal.cuh"

namespace bke {

template <typename T, int N, int M>
struct KfRegs {
    T x[N];
    T P[N][N];
};

template <typename T, int N>
__device__ __forceinline__ void reg_predict(T (&x)[N], T (&P)[N][N], const T (&F)[N][N], const T (&Q)[N][N], T alpha_sq)
{
    T xn[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        T s = F[i][0] * x[0];
#pragma unroll
        for (int k = 1; k < N; k++) s += F[i][k] * x[k];
        xn[i] = s;
    }
    T FP[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            T s = F[i][0] * P[0][j];
#pragma unroll
            for (int k = 1; k < N; k++) s += F[i][k] * P[k][j];
            FP[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            T s = FP[i][0] * F[j][0];
#pragma unroll
            for (int k = 1; k < N; k++) s += FP[i][k] * F[j][k];
            P[i][j] = alpha_sq * s + Q[i][j];
        }
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = xn[i];
}

// In-register inverse of a small M x M matrix: closed form for M <= 2, Gauss-Jordan with partial
// pivoting for M >= 3 (an indefinite but non-singular S — numerical drift of P, a user R that is not
// PD — inverts exactly where np.linalg.inv does).  Returns false on a zero pivot = singular S.
// logdet = log |det S|.
template <typename T, int M>
__device__ __forceinline__ bool reg_inverse(const T (&S)[M][M], T (&SI)[M][M], T &logdet)
{
    if constexpr (M == 1) {
        logdet = log(fabs(S[0][0]));
        SI[0][0] = T(1) / S[0][0];
        return S[0][0] != T(0);
    } else if constexpr (M == 2) {
        T det = S[0][0] * S[1][1] - S[0][1] * S[1][0];
        T r = T(1) / det;
        SI[0][0] = S[1][1] * r; SI[0][1] = -S[0][1] * r;
        SI[1][0] = -S[1][0] * r; SI[1][1] = S[0][0] * r;
        logdet = log(fabs(det));
        return det != T(0);
    } else {
        T A[M][M];
        bool ok = true;
        T ld = T(0);
#pragma unroll
        for (int i = 0; i < M; i++)
#pragma unroll
            for (int j = 0; j < M; j++) { A[i][j] = S[i][j]; SI[i][j] = (i == j) ? T(1) : T(0); }
#pragma unroll
        for (int c = 0; c < M; c++) {
            // partial pivoting, as np.linalg.inv (LU, kalman_filter.py:541): the largest |entry| of
            // column c at or below the diagonal is bubbled into row c with predicated row swaps of
            // [A | SI] (row operations on both sides leave SI = S^-1; |det| is unchanged)
            T best = fabs(A[c][c]);
#pragma unroll
            for (int r = c + 1; r < M; r++) {
                const T cand = fabs(A[r][c]);
                const bool sw = cand > best;
                best = sw ? cand : best;
#pragma unroll
                for (int j = 0; j < M; j++) {
                    const T a0 = A[c][j], a1 = A[r][j], s0 = SI[c][j], s1 = SI[r][j];
                    A[c][j] = sw ? a1 : a0; A[r][j] = sw ? a0 : a1;
                    SI[c][j] = sw ? s1 : s0; SI[r][j] = sw ? s0 : s1;
                }
            }
            T piv = A[c][c];
            ok = ok && (piv != T(0));
            ld += log(fabs(piv));
            T d = T(1) / piv;
#pragma unroll
            for (int j = 0; j < M; j++) { A[c][j] *= d; SI[c][j] *= d; }
#pragma unroll
            for (int r = 0; r < M; r++) {
                if (r != c) {
                    T f = A[r][c];
#pragma unroll
                    for (int j = 0; j < M; j++) { A[r][j] -= f * A[c][j]; SI[r][j] -= f * SI[c][j]; }
                }
            }
        }
        logdet = ld;
        return ok;
    }
}

template <typename T, int N, int M>
struct KfUpdateOut {
    T y[M];
    T K[N][M];
    T S[M][M];
    T SI[M][M];
    T logdet;
    bool ok;
};

template <typename T, int N, int M>
__device__ __forceinline__ void reg_update(T (&x)[N], T (&P)[N][N], const T (&H)[M][N], const T (&R)[M][M],
                                           const T (&z)[M], KfUpdateOut<T, N, M> &o)
{
#pragma unroll
    for (int a = 0; a < M; a++) {
        T s = H[a][0] * x[0];
#pragma unroll
        for (int k = 1; k < N; k++) s += H[a][k] * x[k];
        o.y[a] = z[a] - s;
    }
    T PHT[N][M];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int a = 0; a < M; a++) {
            T s = P[i][0] * H[a][0];
#pragma unroll
            for (int k = 1; k < N; k++) s += P[i][k] * H[a][k];
            PHT[i][a] = s;
        }
#pragma unroll
    for (int a = 0; a < M; a++)
#pragma unroll
        for (int b = 0; b < M; b++) {
            T s = H[a][0] * PHT[0][b];
#pragma unroll
            for (int k = 1; k < N; k++) s += H[a][k] * PHT[k][b];
            o.S[a][b] = s + R[a][b];
        }
    o.ok = reg_inverse<T, M>(o.S, o.SI, o.logdet);
    if (!o.ok) return;      // np.linalg.inv would raise; state stays at the prior
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int a = 0; a < M; a++) {
            T s = PHT[i][0] * o.SI[0][a];
#pragma unroll
            for (int b = 1; b < M; b++) s += PHT[i][b] * o.SI[b][a];
            o.K[i][a] = s;
        }
#pragma unroll
    for (int i = 0; i < N; i++) {
        T s = x[i];
#pragma unroll
        for (int a = 0; a < M; a++) s += o.K[i][a] * o.y[a];
        x[i] = s;
    }
    T IKH[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            T s = (i == j) ? T(1) : T(0);
#pragma unroll
            for (int a = 0; a < M; a++) s -= o.K[i][a] * H[a][j];
            IKH[i][j] = s;
        }
    T T1[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            T s = IKH[i][0] * P[0][j];
#pragma unroll
            for (int k = 1; k < N; k++) s += IKH[i][k] * P[k][j];
            T1[i][j] = s;
        }
    T KR[N][M];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int b = 0; b < M; b++) {
            T s = o.K[i][0] * R[0][b];
#pragma unroll
            for (int a = 1; a < M; a++) s += o.K[i][a] * R[a][b];
            KR[i][b] = s;
        }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            T s = T1[i][0] * IKH[j][0];
#pragma unroll
            for (int k = 1; k < N; k++) s += T1[i][k] * IKH[j][k];
#pragma unroll
            for (int a = 0; a < M; a++) s += KR[i][a] * o.K[j][a];
            P[i][j] = s;
        }
}

}  // namespace bke
