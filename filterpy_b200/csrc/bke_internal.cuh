// Internal helpers shared by the sm_100a kernels of the engine (not part of the C-ABI).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#endif
#include "../../include/bke.h"

namespace bke {

#ifndef __CUDACC_RTC__
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

// Number of SMs of the current device (cached).
int sm_count();
#endif

constexpr unsigned FULL = 0xffffffffu;

template <typename T> struct DType;
template <> struct DType<float> { static constexpr int id = BKE_F32; };
template <> struct DType<double> { static constexpr int id = BKE_F64; };

// log(2*pi)
constexpr double LOG_2PI = 1.8378770664093454835606594728112;

#ifndef __CUDACC_RTC__
// ---- launchers implemented in the .cu files ---------------------------------------------
int launch_kf_generic(const bke_kf_args &a, cudaStream_t s);
// returns BKE_ERR_UNSUPPORTED when the specialised kernel does not cover the call
int launch_kf_fast(const bke_kf_args &a, cudaStream_t s);
int launch_kf_rowblock(const bke_kf_args &a, cudaStream_t s);
int launch_kf_direct(const bke_kf_args &a, cudaStream_t s);
// tcgen05 covariance propagation for shared-model fp32 banks with dim_x = 16 / 32 (kf_tc.cu); a fused step runs
// its update through launch_kf_any afterwards
int launch_kf_tc(const bke_kf_args &a, cudaStream_t s);
// the dispatch order of bke_kf_step: tensor-core predict (dim_x 16 / 32, shared models) -> TMA-staged 4/2 fp32 ->
// register tile with direct loads -> row-block -> catch-all
int launch_kf_any(const bke_kf_args &a, cudaStream_t s);
int launch_kf_batch(const bke_kf_batch_args &a, cudaStream_t s);
int launch_ukf(const bke_ukf_args &a, cudaStream_t s);
#endif

}  // namespace bke
