// ukf_rtc.cu — UKF instances around USER-SUPPLIED process / measurement functions.
//
// The reference takes fx(x, dt, **args) and hx(x, **args) as Python callables (filterpy/kalman/UKF.py:
// 284-288, called at :521-522 and :463-464).  A device cannot call back into Python, so the drop-in
// takes them as CUDA C++ source text instead: bke_ukf_model_compile() wraps the text around the very
// kernel the pre-built instances use (ukf_kernel.cuh, model ids BKE_FX_USER / BKE_HX_USER), compiles it
// for sm_100a with NVRTC (libnvrtc is dlopen'ed: libbke.so does not link it), loads the cubin with
// cudaLibraryLoadData and launches the resulting cudaKernel_t like any other kernel.  No CPU path.
//
// Program text handed to NVRTC (the user's part between the markers):
//     typedef double real;                       // or float
//     #define BKE_DIM_X 6 / BKE_DIM_Z 3
//     #include "ukf_kernel.cuh"
//     /* user */ __device__ void fx(const real *x, real *out, real dt, const real *args) { ... }
//     /* user */ __device__ void hx(const real *x, real *z, const real *args) { ... }
//     template <> bke_user_fx<real> -> ::fx, bke_user_hx<real> -> ::hx
#include <dlfcn.h>
#include <nvrtc.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "ukf_launch.cuh"
#include "ukf_rts_launch.cuh"

struct bke_ukf_model {
    int n, m, dtype, fx_model, hx_model;
    cudaLibrary_t lib;
    cudaKernel_t kern[2];          // [0] plain, [1] with the optional outputs
    cudaKernel_t kern_rts;         // RTS smoother around the user's fx (NULL: fx is built in, or dim_x > UR_MAXN)
    int regs[2];
    std::string log;
};

namespace bke {
namespace {

struct Nvrtc {
    void *h = nullptr;
    decltype(&nvrtcCreateProgram) create;
    decltype(&nvrtcCompileProgram) compile;
    decltype(&nvrtcDestroyProgram) destroy;
    decltype(&nvrtcGetProgramLogSize) log_size;
    decltype(&nvrtcGetProgramLog) log;
    decltype(&nvrtcGetCUBINSize) cubin_size;
    decltype(&nvrtcGetCUBIN) cubin;
    decltype(&nvrtcAddNameExpression) add_name;
    decltype(&nvrtcGetLoweredName) lowered;
    decltype(&nvrtcGetErrorString) errstr;
};

// libnvrtc: BKE_NVRTC_LIB (set by the Python loader to the copy next to torch's CUDA libraries), the
// loader path, then the toolkit directory
Nvrtc *nvrtc()
{
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n.h ? &n : nullptr;
    tried = true;
    std::vector<std::string> cand;
    if (const char *e = getenv("BKE_NVRTC_LIB")) cand.push_back(e);
    cand.push_back("libnvrtc.so.12");
    cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    cand.push_back("libnvrtc.so");
    for (const auto &c : cand) {
        n.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (n.h) break;
    }
    if (!n.h) return nullptr;
#define BKE_SYM(field, name) n.field = (decltype(n.field))dlsym(n.h, #name); if (!n.field) { n.h = nullptr; return nullptr; }
    BKE_SYM(create, nvrtcCreateProgram) BKE_SYM(compile, nvrtcCompileProgram) BKE_SYM(destroy, nvrtcDestroyProgram)
    BKE_SYM(log_size, nvrtcGetProgramLogSize) BKE_SYM(log, nvrtcGetProgramLog) BKE_SYM(cubin_size, nvrtcGetCUBINSize)
    BKE_SYM(cubin, nvrtcGetCUBIN) BKE_SYM(add_name, nvrtcAddNameExpression) BKE_SYM(lowered, nvrtcGetLoweredName)
    BKE_SYM(errstr, nvrtcGetErrorString)
#undef BKE_SYM
    return &n;
}

std::string kernel_name(const bke_ukf_model &m, int occ, bool extras)
{
    char buf[256];
    snprintf(buf, sizeof buf, "bke::ukfk::ukf_kernel<real, %d, %d, %d, %d, %d, %s>", m.n, m.m, m.fx_model, m.hx_model, occ,
             extras ? "true" : "false");
    return buf;
}

template <typename T>
int launch_model(const bke_ukf_args &a, const bke_ukf_model &m, const void *fx_args, int64_t s_fx, const void *hx_args, int64_t s_hx,
                 cudaStream_t s)
{
    ukfk::UkfP<T> p;
    ukf_fill_params<T>(a, m.n, p);
    p.fx_args = (const T *)fx_args; p.s_fx_args = s_fx;
    p.hx_args = (const T *)hx_args; p.s_hx_args = s_hx;
    const size_t smem = ukf_smem_bytes<T>(m.n, m.m, m.fx_model == BKE_FX_LINEAR, a.F_stride == 0, m.hx_model == BKE_HX_LINEAR, a.H_stride == 0);
    const bool ex = a.x_prior || a.P_prior || a.K || a.y || a.S || a.SI || a.log_likelihood;
    const void *kern = (const void *)m.kern[ex ? 1 : 0];
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute")) return BKE_ERR_CUDA;
    const int64_t grid = (p.N + ukfk::UB - 1) / ukfk::UB;
    void *params[] = {&p};
    if (check_cuda(cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(ukfk::UB), params, smem, s), "ukf model launch")) return BKE_ERR_CUDA;
    return BKE_OK;
}

}  // namespace
}  // namespace bke

using namespace bke;

extern "C" {

// NVRTC half of bke_ukf_model_compile (needs no GPU): program text -> sm_100a cubin + the lowered names
// of the two kernel instances
static int compile_cubin(int32_t dim_x, int32_t dim_z, int32_t dtype, int32_t fx_model, int32_t hx_model, const char *source,
                         const char *include_dirs, std::vector<char> &cubin, std::string (&lowered)[3], std::string &log)
{
    if (dim_x < 1 || dim_x > 16 || dim_z < 1 || dim_z > dim_x + 8) { set_error("bke_ukf_model_compile: 1 <= dim_x <= 16, 1 <= dim_z"); return BKE_ERR_BAD_ARG; }
    if (dtype != BKE_F32 && dtype != BKE_F64) { set_error("dtype must be BKE_F32 or BKE_F64"); return BKE_ERR_BAD_ARG; }
    const bool ufx = fx_model == BKE_FX_USER, uhx = hx_model == BKE_HX_USER;
    if (!ufx && !uhx) { set_error("bke_ukf_model_compile: neither fx nor hx is BKE_*_USER (use bke_ukf_step)"); return BKE_ERR_BAD_ARG; }
    if ((!ufx && fx_model != BKE_FX_LINEAR && fx_model != BKE_FX_CONST_VEL) || (!uhx && hx_model != BKE_HX_LINEAR)) {
        set_error("bke_ukf_model_compile: the built-in partner of a user function must be BKE_FX_LINEAR / BKE_FX_CONST_VEL / BKE_HX_LINEAR");
        return BKE_ERR_BAD_ARG;
    }
    if (!ufx && fx_model == BKE_FX_CONST_VEL && (dim_x & 1)) { set_error("BKE_FX_CONST_VEL needs an even dim_x"); return BKE_ERR_BAD_ARG; }
    if (!source || !include_dirs) { set_error("source and include_dirs must be non-NULL"); return BKE_ERR_BAD_ARG; }
    Nvrtc *rt = nvrtc();
    if (!rt) { set_error("libnvrtc.so.12 not found (set BKE_NVRTC_LIB)"); return BKE_ERR_UNSUPPORTED; }

    std::string text;
    text += dtype == BKE_F64 ? "typedef double real;\n" : "typedef float real;\n";
    text += "#define BKE_DIM_X " + std::to_string(dim_x) + "\n#define BKE_DIM_Z " + std::to_string(dim_z) + "\n";
    text += "#include \"ukf_kernel.cuh\"\n#include \"ukf_rts_kernel.cuh\"\n";
    text += "#line 1 \"user_model.cu\"\n";
    text += source;
    text += "\n#line 1 \"bke_glue.cu\"\nnamespace bke { namespace ukfk {\n";
    if (ufx) text += "template <> __device__ __forceinline__ void bke_user_fx<real>(const real *x, real *out, real dt, const real *args) { ::fx(x, out, dt, args); }\n";
    if (uhx) text += "template <> __device__ __forceinline__ void bke_user_hx<real>(const real *x, real *z, const real *args) { ::hx(x, z, args); }\n";
    text += "} }\n";

    // resident CTAs the instance is compiled for: the pre-built kernels' choice (ukf.cu)
    const int occ = dim_x >= 6 ? (dtype == BKE_F64 ? 3 : 5) : 1;
    bke_ukf_model tmp;
    tmp.n = dim_x; tmp.m = dim_z; tmp.fx_model = fx_model; tmp.hx_model = hx_model;
    nvrtcProgram prog;
    nvrtcResult r = rt->create(&prog, text.c_str(), "bke_ukf_user.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) { set_error("nvrtcCreateProgram: %s", rt->errstr(r)); return BKE_ERR_CUDA; }
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};     // (bke.h declares the host C-ABI)
    {
        std::string dirs = include_dirs;
        size_t pos = 0;
        while (pos <= dirs.size()) {
            size_t e = dirs.find(':', pos);
            if (e == std::string::npos) e = dirs.size();
            if (e > pos) opts.push_back("-I" + dirs.substr(pos, e - pos));
            pos = e + 1;
        }
    }
    std::vector<const char *> copts;
    for (auto &o : opts) copts.push_back(o.c_str());
    // the step kernel with / without the optional outputs and, around a user fx, the RTS smoother
    const bool with_rts = ufx && dim_x <= UR_MAXN;
    const int n_names = with_rts ? 3 : 2;
    const std::string names[3] = {kernel_name(tmp, occ, false), kernel_name(tmp, occ, true), "bke::ukf_rts_kernel<real, true>"};
    for (int i = 0; i < n_names; i++) rt->add_name(prog, names[i].c_str());
    r = rt->compile(prog, (int)copts.size(), copts.data());
    size_t lsz = 0;
    rt->log_size(prog, &lsz);
    log.clear();
    if (lsz > 1) { log.resize(lsz); rt->log(prog, &log[0]); }
    if (r != NVRTC_SUCCESS) {
        set_error("NVRTC could not compile the UKF model: %s\n%s", rt->errstr(r), log.c_str());
        rt->destroy(&prog);
        return BKE_ERR_BAD_ARG;
    }
    size_t csz = 0;
    rt->cubin_size(prog, &csz);
    cubin.resize(csz);
    rt->cubin(prog, cubin.data());
    lowered[2].clear();
    for (int i = 0; i < n_names; i++) {
        const char *ln = nullptr;
        if (rt->lowered(prog, names[i].c_str(), &ln) != NVRTC_SUCCESS || !ln) {
            set_error("nvrtcGetLoweredName failed for %s", names[i].c_str());
            rt->destroy(&prog);
            return BKE_ERR_CUDA;
        }
        lowered[i] = ln;
    }
    rt->destroy(&prog);
    return BKE_OK;
}

int bke_ukf_model_compile(int32_t dim_x, int32_t dim_z, int32_t dtype, int32_t fx_model, int32_t hx_model, const char *source,
                          const char *include_dirs, bke_ukf_model **out)
{
    if (!out) { set_error("out is NULL"); return BKE_ERR_BAD_ARG; }
    *out = nullptr;
    std::vector<char> cubin;
    std::string lowered[3], log;
    int rc = compile_cubin(dim_x, dim_z, dtype, fx_model, hx_model, source, include_dirs, cubin, lowered, log);
    if (rc != BKE_OK) return rc;
    bke_ukf_model *m = new bke_ukf_model();
    m->n = dim_x; m->m = dim_z; m->dtype = dtype; m->fx_model = fx_model; m->hx_model = hx_model; m->lib = nullptr; m->log = log;
    m->kern_rts = nullptr;
    if (check_cuda(cudaLibraryLoadData(&m->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0), "cudaLibraryLoadData")) { delete m; return BKE_ERR_CUDA; }
    for (int i = 0; i < 2; i++) {
        if (check_cuda(cudaLibraryGetKernel(&m->kern[i], m->lib, lowered[i].c_str()), "cudaLibraryGetKernel")) {
            cudaLibraryUnload(m->lib); delete m;
            return BKE_ERR_CUDA;
        }
        cudaFuncAttributes fa;
        m->regs[i] = cudaFuncGetAttributes(&fa, (const void *)m->kern[i]) == cudaSuccess ? fa.numRegs : -1;
    }
    if (!lowered[2].empty() && check_cuda(cudaLibraryGetKernel(&m->kern_rts, m->lib, lowered[2].c_str()), "cudaLibraryGetKernel (rts)")) {
        cudaLibraryUnload(m->lib); delete m;
        return BKE_ERR_CUDA;
    }
    cudaGetLastError();
    *out = m;
    return BKE_OK;
}

// the NVRTC half alone (CPU-only check that a model's text compiles for sm_100a): cubin size or 0
size_t bke_debug_ukf_model_cubin_bytes(int32_t dim_x, int32_t dim_z, int32_t dtype, int32_t fx_model, int32_t hx_model, const char *source,
                                       const char *include_dirs)
{
    std::vector<char> cubin;
    std::string lowered[3], log;
    if (compile_cubin(dim_x, dim_z, dtype, fx_model, hx_model, source, include_dirs, cubin, lowered, log) != BKE_OK) return 0;
    return cubin.size();
}

const char *bke_ukf_model_log(const bke_ukf_model *m) { return m ? m->log.c_str() : ""; }

int bke_ukf_model_registers(const bke_ukf_model *m, int32_t extras) { return m ? m->regs[extras ? 1 : 0] : -1; }

void bke_ukf_model_free(bke_ukf_model *m)
{
    if (!m) return;
    if (m->lib) cudaLibraryUnload(m->lib);
    delete m;
}

int bke_ukf_step_model(const bke_ukf_args *args, const bke_ukf_model *model, const void *fx_args, int64_t fx_args_stride,
                       const void *hx_args, int64_t hx_args_stride, void *stream)
{
    if (!args || !model) { set_error("args / model is NULL"); return BKE_ERR_BAD_ARG; }
    const bke_ukf_args &a = *args;
    if (a.dim_x != model->n || a.dim_z != model->m || a.dtype != model->dtype || a.fx_model != model->fx_model || a.hx_model != model->hx_model) {
        set_error("bke_ukf_step_model: args (dim_x=%d dim_z=%d dtype=%d fx=%d hx=%d) do not match the compiled model (%d %d %d %d %d)",
                  a.dim_x, a.dim_z, a.dtype, a.fx_model, a.hx_model, model->n, model->m, model->dtype, model->fx_model, model->hx_model);
        return BKE_ERR_BAD_ARG;
    }
    if (a.n_filters < 0) { set_error("bad dimensions"); return BKE_ERR_BAD_ARG; }
    if (!(a.flags & (BKE_DO_PREDICT | BKE_DO_UPDATE))) { set_error("flags selects neither predict nor update"); return BKE_ERR_BAD_ARG; }
    if (!a.x || !a.P || !a.x_out || !a.P_out) { set_error("x, P, x_out, P_out must be non-NULL"); return BKE_ERR_BAD_ARG; }
    if ((a.flags & BKE_DO_PREDICT) && !a.Q) { set_error("predict needs Q"); return BKE_ERR_BAD_ARG; }
    if ((a.flags & BKE_DO_UPDATE) && (!a.R || !a.z)) { set_error("update needs R and z"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model == BKE_FX_LINEAR && (a.flags & BKE_DO_PREDICT) && !a.F) { set_error("BKE_FX_LINEAR needs F"); return BKE_ERR_BAD_ARG; }
    if (a.hx_model == BKE_HX_LINEAR && (a.flags & BKE_DO_UPDATE) && !a.H) { set_error("BKE_HX_LINEAR needs H"); return BKE_ERR_BAD_ARG; }
    if (a.fx_model == BKE_FX_CONST_VEL && (a.dim_x & 1)) { set_error("BKE_FX_CONST_VEL needs an even dim_x"); return BKE_ERR_BAD_ARG; }
    if (fx_args_stride < 0 || hx_args_stride < 0) { set_error("negative args stride"); return BKE_ERR_BAD_ARG; }
    const double lam_n = a.alpha * a.alpha * (a.dim_x + a.kappa);
    if (!(lam_n != 0.0)) { set_error("alpha^2 (n + kappa) must be non-zero"); return BKE_ERR_BAD_ARG; }
    if (a.n_filters == 0) return BKE_OK;
    return a.dtype == BKE_F32 ? launch_model<float>(a, *model, fx_args, fx_args_stride, hx_args, hx_args_stride, (cudaStream_t)stream)
                              : launch_model<double>(a, *model, fx_args, fx_args_stride, hx_args, hx_args_stride, (cudaStream_t)stream);
}

int bke_ukf_rts_smoother_model(const bke_ukf_rts_args *args, const bke_ukf_model *model, const void *fx_args, int64_t fx_args_stride,
                               void *stream)
{
    if (!args || !model) { set_error("args / model is NULL"); return BKE_ERR_BAD_ARG; }
    const bke_ukf_rts_args &a = *args;
    if (!model->kern_rts) { set_error("bke_ukf_rts_smoother_model: the model has no user fx (use bke_ukf_rts_smoother) or dim_x > %d", UR_MAXN); return BKE_ERR_UNSUPPORTED; }
    if (a.dim_x != model->n || a.dtype != model->dtype || a.fx_model != BKE_FX_USER) { set_error("bke_ukf_rts_smoother_model: args do not match the compiled model"); return BKE_ERR_BAD_ARG; }
    if (a.n_filters < 0 || a.n_steps < 0 || fx_args_stride < 0 || a.Q_stride < 0) { set_error("negative sizes"); return BKE_ERR_BAD_ARG; }
    if (a.n_filters == 0 || a.n_steps == 0) return BKE_OK;
    if (!a.Xs || !a.Ps || !a.Q || !a.x_out || !a.P_out) { set_error("NULL argument"); return BKE_ERR_BAD_ARG; }
    void *params[1];
    UrP<float> pf; UrP<double> pd;
    if (a.dtype == BKE_F32) { ukf_rts_fill_params<float>(a, pf); pf.fx_args = (const float *)fx_args; pf.s_fx_args = fx_args_stride; params[0] = &pf; }
    else { ukf_rts_fill_params<double>(a, pd); pd.fx_args = (const double *)fx_args; pd.s_fx_args = fx_args_stride; params[0] = &pd; }
    if (check_cuda(cudaLaunchKernel((const void *)model->kern_rts, dim3((unsigned)((a.n_filters + 63) / 64)), dim3(64), params, 0, (cudaStream_t)stream),
                   "ukf rts model launch")) return BKE_ERR_CUDA;
    return BKE_OK;
}

}  // extern "C"
