// ukf_rts.cu — host side of UnscentedKalmanFilter.rts_smoother for a bank (kernel: ukf_rts_kernel.cuh).
#include "ukf_rts_kernel.cuh"
#include "ukf_rts_launch.cuh"

namespace bke {
namespace {

template <typename T>
int launch_t(const bke_ukf_rts_args &a, cudaStream_t s)
{
    UrP<T> p;
    ukf_rts_fill_params<T>(a, p);
    ukf_rts_kernel<T, false><<<(unsigned)((p.N + 63) / 64), 64, 0, s>>>(p);
    return check_cuda(cudaGetLastError(), "ukf rts launch");
}

}  // namespace
}  // namespace bke

using namespace bke;

extern "C" int bke_ukf_rts_smoother(const bke_ukf_rts_args *args, void *stream)
{
    if (!args) { set_error("args is NULL"); return BKE_ERR_BAD_ARG; }
    const bke_ukf_rts_args &a = *args;
    if (a.n_filters < 0 || a.n_steps < 0) { set_error("negative sizes"); return BKE_ERR_BAD_ARG; }
    if (a.dim_x < 1 || a.dim_x > UR_MAXN) { set_error("bke_ukf_rts_smoother: dim_x must be in [1, %d]", UR_MAXN); return BKE_ERR_UNSUPPORTED; }
    if (a.dtype != BKE_F32 && a.dtype != BKE_F64) { set_error("bad dtype"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model != BKE_FX_LINEAR && a.fx_model != BKE_FX_CONST_VEL) { set_error("unknown fx_model %d", a.fx_model); return BKE_ERR_UNSUPPORTED; }
    if (a.fx_model == BKE_FX_CONST_VEL && (a.dim_x & 1)) { set_error("BKE_FX_CONST_VEL needs an even dim_x"); return BKE_ERR_BAD_ARG; }
    if (a.n_filters == 0 || a.n_steps == 0) return BKE_OK;
    if (!a.Xs || !a.Ps || !a.Q || !a.x_out || !a.P_out) { set_error("NULL argument"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model == BKE_FX_LINEAR && !a.F) { set_error("BKE_FX_LINEAR needs F"); return BKE_ERR_BAD_ARG; }
    if (a.Q_stride < 0 || a.F_stride < 0) { set_error("negative stride"); return BKE_ERR_BAD_ARG; }
    if (bke_device_count() <= 0) { set_error("no CUDA device"); return BKE_ERR_CUDA; }
    return a.dtype == BKE_F32 ? launch_t<float>(a, (cudaStream_t)stream) : launch_t<double>(a, (cudaStream_t)stream);
}
