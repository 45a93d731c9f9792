// api.cu — the extern "C" boundary declared in include/bke.h: argument validation, error
// text, dispatch to the sm_100a kernels.  No torch types, no allocation, no host sync.
#include <stdarg.h>
#include <string.h>
#include "bke_internal.cuh"

namespace bke {

static thread_local char g_err[8192] = "";     // room for an NVRTC log (bke_ukf_model_compile)

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return BKE_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return BKE_ERR_CUDA;
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device available (%s); the engine has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return BKE_ERR_CUDA;
    }
    return BKE_OK;
}

static int validate_kf(const bke_kf_args *a, bool need_z)
{
    if (!a) { set_error("args is NULL"); return BKE_ERR_BAD_ARG; }
    if (a->n_filters < 0) { set_error("n_filters < 0"); return BKE_ERR_BAD_ARG; }
    if (a->dim_x < 1) { set_error("dim_x must be 1 or greater"); return BKE_ERR_BAD_ARG; }   // kalman_filter.py:388
    if (a->dim_z < 1) { set_error("dim_z must be 1 or greater"); return BKE_ERR_BAD_ARG; }   // :390
    if (a->dim_u < 0) { set_error("dim_u must be 0 or greater"); return BKE_ERR_BAD_ARG; }   // :392
    if (a->dtype != BKE_F32 && a->dtype != BKE_F64) { set_error("dtype must be BKE_F32 or BKE_F64"); return BKE_ERR_BAD_ARG; }
    if (!(a->flags & (BKE_DO_PREDICT | BKE_DO_UPDATE))) { set_error("flags selects neither predict nor update"); return BKE_ERR_BAD_ARG; }
    if (!a->x || !a->P || !a->x_out || !a->P_out) { set_error("x, P, x_out, P_out must be non-NULL"); return BKE_ERR_BAD_ARG; }
    if (a->flags & BKE_DO_PREDICT) {
        if (!a->F || !a->Q) { set_error("predict needs F and Q"); return BKE_ERR_BAD_ARG; }
        if ((a->B == nullptr) != (a->u == nullptr) && a->dim_u > 0 && a->B && !a->u) { /* u=None: no control, kalman_filter.py:472 */ }
    }
    if (a->flags & BKE_DO_UPDATE) {
        if (!a->H || !a->R) { set_error("update needs H and R"); return BKE_ERR_BAD_ARG; }
        if (need_z && !a->z) { set_error("update needs z"); return BKE_ERR_BAD_ARG; }
    }
    const int64_t n = a->dim_x, m = a->dim_z;
    auto bad_stride = [](int64_t s, int64_t full) { return s != 0 && s != full; };
    if (bad_stride(a->F_stride, n * n) || bad_stride(a->Q_stride, n * n) || bad_stride(a->H_stride, m * n) ||
        bad_stride(a->R_stride, m * m)) {
        set_error("model strides must be 0 (shared) or the dense per-filter size");
        return BKE_ERR_BAD_ARG;
    }
    return BKE_OK;
}

}  // namespace bke

using namespace bke;

namespace bke {
int launch_kf_any(const bke_kf_args &a, cudaStream_t s)
{
    int rc = launch_kf_tc(a, s);
    if (rc == BKE_ERR_UNSUPPORTED) rc = launch_kf_fast(a, s);
    if (rc == BKE_ERR_UNSUPPORTED) rc = launch_kf_direct(a, s);
    if (rc == BKE_ERR_UNSUPPORTED) rc = launch_kf_rowblock(a, s);
    if (rc == BKE_ERR_UNSUPPORTED) rc = launch_kf_generic(a, s);
    return rc;
}
}  // namespace bke

extern "C" {

int bke_abi_version(void) { return BKE_ABI_VERSION; }

const char *bke_last_error(void) { return g_err; }

int bke_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int bke_kf_step(const bke_kf_args *args, void *stream)
{
    int rc = validate_kf(args, true);
    if (rc) return rc;
    if ((rc = require_device())) return rc;
    if (args->n_filters == 0) return BKE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    return launch_kf_any(*args, s);
}

int bke_kf_batch_filter(const bke_kf_batch_args *args, void *stream)
{
    if (!args) { set_error("args is NULL"); return BKE_ERR_BAD_ARG; }
    bke_kf_args st = args->step;
    st.flags |= BKE_DO_PREDICT | BKE_DO_UPDATE;
    int rc = validate_kf(&st, false);
    if (rc) return rc;
    if (args->n_steps < 0) { set_error("n_steps < 0"); return BKE_ERR_BAD_ARG; }
    if (args->n_steps > 0 && !args->zs) { set_error("zs is NULL"); return BKE_ERR_BAD_ARG; }
    if ((rc = require_device())) return rc;
    if (st.n_filters == 0) return BKE_OK;
    bke_kf_batch_args a = *args;
    a.step = st;
    return launch_kf_batch(a, (cudaStream_t)stream);
}

int bke_ukf_step(const bke_ukf_args *args, void *stream)
{
    if (!args) { set_error("args is NULL"); return BKE_ERR_BAD_ARG; }
    const bke_ukf_args &a = *args;
    if (a.n_filters < 0 || a.dim_x < 1 || a.dim_z < 1) { set_error("bad dimensions"); return BKE_ERR_BAD_ARG; }
    if (a.dtype != BKE_F32 && a.dtype != BKE_F64) { set_error("dtype must be BKE_F32 or BKE_F64"); return BKE_ERR_BAD_ARG; }
    if (!(a.flags & (BKE_DO_PREDICT | BKE_DO_UPDATE))) { set_error("flags selects neither predict nor update"); return BKE_ERR_BAD_ARG; }
    if (!a.x || !a.P || !a.x_out || !a.P_out) { set_error("x, P, x_out, P_out must be non-NULL"); return BKE_ERR_BAD_ARG; }
    if ((a.flags & BKE_DO_PREDICT) && !a.Q) { set_error("predict needs Q"); return BKE_ERR_BAD_ARG; }
    if ((a.flags & BKE_DO_UPDATE) && (!a.R || !a.z)) { set_error("update needs R and z"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model == BKE_FX_LINEAR && (a.flags & BKE_DO_PREDICT) && !a.F) { set_error("BKE_FX_LINEAR needs F"); return BKE_ERR_BAD_ARG; }
    if (a.hx_model == BKE_HX_LINEAR && (a.flags & BKE_DO_UPDATE) && !a.H) { set_error("BKE_HX_LINEAR needs H"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model == BKE_FX_CONST_VEL && (a.dim_x & 1)) { set_error("BKE_FX_CONST_VEL needs an even dim_x"); return BKE_ERR_BAD_ARG; }
    if (a.hx_model == BKE_HX_RANGE_AZ_EL && !(a.dim_x == 6 && a.dim_z == 3)) { set_error("BKE_HX_RANGE_AZ_EL needs dim_x=6, dim_z=3"); return BKE_ERR_BAD_ARG; }
    if (a.hx_model == BKE_HX_RANGE_BEARING && !(a.dim_x == 4 && a.dim_z == 2)) { set_error("BKE_HX_RANGE_BEARING needs dim_x=4, dim_z=2"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model < 0 || a.fx_model > BKE_FX_CONST_VEL || a.hx_model < 0 || a.hx_model > BKE_HX_RANGE_BEARING) {
        set_error("unknown fx/hx model id"); return BKE_ERR_BAD_ARG;
    }
    double lam_n = a.alpha * a.alpha * (a.dim_x + a.kappa);
    if (!(lam_n != 0.0)) { set_error("alpha^2 (n + kappa) must be non-zero"); return BKE_ERR_BAD_ARG; }
    int rc = require_device();
    if (rc) return rc;
    if (a.n_filters == 0) return BKE_OK;
    return launch_ukf(a, (cudaStream_t)stream);
}

}  // extern "C"
