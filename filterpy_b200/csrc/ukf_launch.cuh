// ukf_launch.cuh — host helpers shared by the pre-built (ukf.cu) and the run-time compiled (ukf_rtc.cu)
// instances of the UKF kernel: parameter block and dynamic shared-memory size.
#pragma once
#include "ukf_kernel.cuh"

namespace bke {

template <typename T>
inline void ukf_fill_params(const bke_ukf_args &a, int N, ukfk::UkfP<T> &p)
{
    const double lambda_ = a.alpha * a.alpha * (N + a.kappa) - N;         // sigma_points.py:167
    const double c = .5 / (N + lambda_);
    p.N = a.n_filters; p.flags = a.flags; p.dt = (T)a.dt;
    p.scale = (T)(lambda_ + N);
    p.wm0 = (T)(lambda_ / (N + lambda_));
    p.wc0 = (T)(lambda_ / (N + lambda_) + (1 - a.alpha * a.alpha + a.beta));
    p.wi = (T)c;
    p.x = (const T *)a.x; p.P = (const T *)a.P; p.Q = (const T *)a.Q; p.R = (const T *)a.R;
    p.F = (const T *)a.F; p.H = (const T *)a.H; p.z = (const T *)a.z;
    p.sQ = a.Q_stride; p.sR = a.R_stride; p.sF = a.F_stride; p.sH = a.H_stride;
    p.valid = a.z_valid;
    p.x_out = (T *)a.x_out; p.P_out = (T *)a.P_out; p.x_prior = (T *)a.x_prior; p.P_prior = (T *)a.P_prior;
    p.K = (T *)a.K; p.y = (T *)a.y; p.S = (T *)a.S; p.SI = (T *)a.SI; p.ll = (T *)a.log_likelihood;
    p.status = a.status;
    p.fx_args = nullptr; p.hx_args = nullptr; p.s_fx_args = 0; p.s_hx_args = 0;
}

// the slab (measurement-space sigma points + parked prior, or one P / Q tile) and the staged F / H
template <typename T>
inline size_t ukf_smem_bytes(int N, int M, bool fx_linear, bool F_shared, bool hx_linear, bool H_shared)
{
    const int PADP = (N * N) | 1;
    const int zpark = (2 * N + 1) * M + N * (N + 1) / 2;
    size_t smem = sizeof(T) * (size_t)(zpark > PADP ? zpark : PADP) * ukfk::UB;
    if (fx_linear) smem += sizeof(T) * (F_shared ? N * N : N * N * ukfk::UB);
    if (hx_linear) smem += sizeof(T) * (H_shared ? M * N : M * N * ukfk::UB);
    return smem;
}

}  // namespace bke
