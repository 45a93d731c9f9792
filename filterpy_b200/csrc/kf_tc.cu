// kf_tc.cu — the tensor-core tile of the linear Kalman filter: covariance propagation
//     x <- F x,   P <- alpha^2 F P F' + Q          (filterpy/kalman/kalman_filter.py:471-478)
// for banks with dim_x = 16 or 32, fp32, whose F and Q are SHARED by the bank (stride 0 — one motion
// model for every track, the usual way a bank is built), on tcgen05.mma with the accumulators in TMEM.
//
// Why only here.  A UMMA tile is M = 128 rows deep.  With per-filter models every 16 x 16 product has
// its own left AND right operand, so a 128-row tile could only be filled block-diagonally (8x wasted
// multiplies, and the operands would have to be re-laid-out per filter).  With a shared F the stacked
// rows of 128 / n covariances ARE one 128 x n operand in the layout they have in HBM:
//     D1[(i,r)][c] = sum_k P_i[r][k] F[c][k]        = (P_i F')[r][c]          A = the P rows, B = F (K-major)
//     D2[(i,c)][j] = sum_k (P_i F')[k][c] F[j][k]   = (F P_i F')[c][j]        A = D1's 16 x 16 blocks transposed
// (P symmetric: (P F')' = F P), i.e. both products of the sandwich have the bank on the M side and the
// one shared matrix on the N side.  Below dim_x = 16 nothing fills an MMA fragment (north_star), at
// dim_x >= 16 the CUDA cores fall behind HBM (fp32: 2 x 2 n^3 = 16 k flop per 2 KB of covariance at
// n = 16, 131 k per 8 KB at n = 32) and this kernel takes the predict of the shared-model banks.
//
// Precision.  kind::tf32 reads 10 mantissa bits, north_star asks for 1e-3 relative on P: one TF32 pass
// per product would sit right at that bound, so every product is the usual three-term split
//     a b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi,   a_hi = a with the low 13 mantissa bits cleared, a_lo = a - a_hi
// accumulated in fp32 in TMEM (relative error ~ 2^-20 per term).  The split is made by the CUDA cores
// while the tile is written into the operand layout, so it costs no extra pass.
//
// One CTA = 4 warps = one 128-row tile at a time (8 filters at n = 16, 4 at n = 32), persistent over
// the tiles; several CTAs per SM overlap each other's loads, MMAs and stores.  Operands are written in
// the no-swizzle K-major canonical layout (8-row x 16-byte core matrices; UMMA descriptors with
// LBO = distance between the two 16-byte K chunks of an instruction, SBO = distance between 8-row
// groups), so no tensor map is needed and the hi / lo split happens on the way in.  One elected thread
// issues the tcgen05.mma's and a tcgen05.commit onto an mbarrier; the four warps read their 32 TMEM
// lanes back with tcgen05.ld (32x32b: lane = row of the tile, registers = the row's n columns).
#include <stdlib.h>
#include "bke_internal.cuh"
#include "kf_regtile.cuh"

namespace bke {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NX>
struct Geom {
    static constexpr int FPT = 128 / NX;            // filters per tile
    static constexpr int KC = NX / 4;               // 16-byte chunks along K
    static constexpr int KS = NX / 8;               // tcgen05.mma steps along K (8 tf32 = 32 bytes each)
    static constexpr int A_LBO = 128 * 16;          // bytes between consecutive K chunks of the 128-row operand
    static constexpr int A_BYTES = 128 * NX * 4;
    static constexpr int B_LBO = NX * 16;           // the same for the NX-row operand (F)
    static constexpr int B_BYTES = NX * NX * 4;
    static constexpr int SBO = 128;                 // bytes between 8-row groups (core matrices are contiguous)
    static constexpr int TMEM_COLS = 2 * NX;        // D1 in columns [0, NX), D2 in [NX, 2 NX): 32 or 64 (powers of two)
    // shared memory: A_hi | A_lo | F_hi | F_lo | F (plain, rows padded to NX + 1 words: x' = F x reads row r in
    // thread r) | Q (rows padded to NX + 4 words: conflict-free 16-byte reads) | mbarrier, TMEM slot.  The
    // transposition scratch [128][NX + 1] lies over A_hi | A_lo (free between the two products).
    static constexpr int FP = NX + 1, QP = NX + 4;
    static constexpr int O_AHI = 0, O_ALO = O_AHI + A_BYTES, O_FHI = O_ALO + A_BYTES, O_FLO = O_FHI + B_BYTES;
    static constexpr int O_F = O_FLO + B_BYTES, O_Q = O_F + ((NX * FP * 4 + 15) & ~15), O_T = O_AHI;
    static constexpr int O_BAR = O_Q + NX * QP * 4, O_X = O_BAR + 64;                            // x of the tile's filters
    // fused update (shared H, R; dim_z <= 4): H and R plain, for the CUDA-core part
    static constexpr int O_H = O_X + 128 * 4, O_R = O_H + 4 * NX * 4;
    static constexpr int SMEM = O_R + 64;
    static_assert(128 * FP * 4 <= 2 * A_BYTES, "the scratch fits the two operand buffers");
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = f32 (bits 4-5 = 1), A = B = tf32
    // (bits 7-9 / 10-12 = 2), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NX >> 3) << 17) | ((128u >> 4) << 24);
};

// byte offset of element (row, k) of a K-major operand with `lbo` bytes between K chunks
__device__ __forceinline__ int op_off(int row, int k, int lbo) { return (k >> 2) * lbo + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4; }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48),
// base offset 0, layout type SWIZZLE_NONE (0) in [61,64)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// bounded wait (a descriptor mistake must not hang the box): false after ~2^22 polls
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (int i = 0; i < (1 << 22); i++) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

template <int NX> __device__ __forceinline__ void tmem_ld_row(uint32_t taddr, float (&v)[NX]);
template <> __device__ __forceinline__ void tmem_ld_row<16>(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
template <> __device__ __forceinline__ void tmem_ld_row<32>(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// round to TF32 (10 mantissa bits, round to nearest, ties away): the tensor core reads exactly these bits, the low 13 are 0
__device__ __forceinline__ float tf32_hi(float v)
{
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
// the remainder v - hi is exact in fp32 (|.| <= 2^-11 |v|); rounded to TF32 itself it leaves 2^-22 |v|
__device__ __forceinline__ float tf32_lo(float v, float hi) { return tf32_hi(v - hi); }

struct TcP {
    int64_t N;                       // filters
    float alpha_sq;
    const float *x, *P, *F, *Q;      // F, Q shared by the bank
    float *x_out, *P_out, *x_prior, *P_prior;
    int32_t *status;
    int sticky;
    int *err;                        // device flag: a wait timed out
    // fused update (kf_cov_tc_kernel<NX, M> with M > 0): H [M, NX] and R [M, M] shared by the bank
    const float *H, *R, *z;
    const uint8_t *valid;
    float *K, *y, *S, *SI, *ll;
};

// M = 0: predict only.  M = dim_z in 1 .. 4: the update of kalman_filter.py:533-556 follows in the same launch (H, R shared).
// CTAs per SM the register budget is planned for: the update's dim_z x dim_z pieces need more registers at dim_z >= 3
constexpr int tc_ctas(int nx, int m) { return nx == 16 ? (m >= 4 ? 5 : (m == 3 ? 6 : 8)) : 4; }

template <int NX, int M>
__global__ void __launch_bounds__(128, tc_ctas(NX, M)) kf_cov_tc_kernel(TcP p)
{
    using G = Geom<NX>;
    extern __shared__ __align__(1024) unsigned char smem[];
    float *Fs = reinterpret_cast<float *>(smem + G::O_F), *Qs = reinterpret_cast<float *>(smem + G::O_Q);
    float *xs = reinterpret_cast<float *>(smem + G::O_X);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + G::O_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + G::O_BAR + 16);
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- one-time setup: the shared model in both forms, the barrier, the TMEM columns
    for (int e = tid; e < NX * NX; e += 128) {
        const float f = p.F[e];
        const int n = e / NX, k = e % NX;
        const float hi = tf32_hi(f);
        *reinterpret_cast<float *>(smem + G::O_FHI + op_off(n, k, G::B_LBO)) = hi;
        *reinterpret_cast<float *>(smem + G::O_FLO + op_off(n, k, G::B_LBO)) = tf32_lo(f, hi);
        Fs[n * G::FP + k] = f;
        Qs[n * G::QP + k] = p.Q[e];
    }
    float *Hs = reinterpret_cast<float *>(smem + G::O_H), *Rs = reinterpret_cast<float *>(smem + G::O_R);
    if constexpr (M > 0) {
        for (int e = tid; e < M * NX; e += 128) Hs[e] = p.H[e];
        if (tid < M * M) Rs[tid] = p.R[tid];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)G::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 TMEM lanes

    const uint32_t a_hi = smem_u32(smem + G::O_AHI), a_lo = smem_u32(smem + G::O_ALO);
    const uint32_t f_hi = smem_u32(smem + G::O_FHI), f_lo = smem_u32(smem + G::O_FLO);
    auto issue_product = [&](uint32_t d_col) {
        // D = A_lo F_hi' + A_hi F_lo' + A_hi F_hi'   (small terms first), K = NX in steps of 8
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < 3; term++) {
            const uint32_t a = term == 0 ? a_lo : a_hi, b = term == 1 ? f_lo : f_hi;
#pragma unroll
            for (int ks = 0; ks < G::KS; ks++) {
                mma_tf32(tmem + d_col, smem_desc(a + ks * 2 * G::A_LBO, G::A_LBO, G::SBO),
                         smem_desc(b + ks * 2 * G::B_LBO, G::B_LBO, G::SBO), G::IDESC, acc);
                acc = 1;
            }
        }
        mma_commit(bar);
    };

    const int64_t rows = p.N * NX;
    const int64_t tiles = (rows + 127) / 128;
    const int i_in_tile = tid / NX, r = tid % NX;         // this thread's row of the tile: filter i, matrix row r
    uint32_t phase = 0;
    bool ok = true;
    // this thread's row of the NEXT tile is fetched while the current tile's products run
    float4 pre[G::KC];
    float xpre = 0.f;                 // x[(tile * FPT + i) * NX + r] = x[tile * 128 + tid]: the thread's own component, coalesced
    auto fetch_row = [&](int64_t tile) {
        const int64_t row = tile * 128 + tid;
        const float4 *src = reinterpret_cast<const float4 *>(p.P + row * NX);
#pragma unroll
        for (int kc = 0; kc < G::KC; kc++) pre[kc] = (row < rows) ? src[kc] : make_float4(0.f, 0.f, 0.f, 0.f);
        xpre = (row < rows) ? p.x[row] : 0.f;
    };
    if ((int64_t)blockIdx.x < tiles) fetch_row(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < tiles && ok; tile += gridDim.x) {
        const int64_t row = tile * 128 + tid;
        const bool live = row < rows;
        const int64_t f = tile * G::FPT + i_in_tile;
        // ---- 1. this thread's row of P -> hi / lo parts in the A-operand layout; x' = F x (own component)
        float xr = 0.f;
        {
#pragma unroll
            for (int kc = 0; kc < G::KC; kc++) {
                const float4 v = pre[kc];
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                const int off = kc * G::A_LBO + (tid >> 3) * 128 + (tid & 7) * 16;
                *reinterpret_cast<float4 *>(smem + G::O_AHI + off) = h;
                *reinterpret_cast<float4 *>(smem + G::O_ALO + off) = make_float4(tf32_lo(v.x, h.x), tf32_lo(v.y, h.y), tf32_lo(v.z, h.z), tf32_lo(v.w, h.w));
            }
            xs[tid] = xpre;
            if (tile + gridDim.x < tiles) fetch_row(tile + gridDim.x);
        }
        tc_fence_before();
        fence_proxy_async();          // generic-proxy writes of the operands -> visible to the tensor core (async proxy)
        __syncthreads();
        // ---- 2. D1 = P F'  (rows (i,r), columns c)
        if (warp == 0) {                 // lane 0 issues; its warp waits for it before polling the barrier
            if (tid == 0) { tc_fence_after(); issue_product(0); }
            __syncwarp();
        }
        {   // x' = F x from the staged state while the tensor core works (the barrier above published xs)
            const float *xf = xs + i_in_tile * NX;
#pragma unroll
            for (int k = 0; k < NX; k++) xr += Fs[r * G::FP + k] * xf[k];
        }
        ok = mbar_wait(bar, phase); phase ^= 1;
        if (!ok) break;
        tc_fence_after();
        {
            // ---- 3. transpose every filter's block on the way back: A2[(i,c)][k] = Y_i[k][c].  Rows go to a padded
            // scratch (conflict-free), then thread (i,c) gathers column c and writes ITS operand row as 16-byte chunks
            float y[NX];
            tmem_ld_row<NX>(lane_base + 0, y);
            float *Ts = reinterpret_cast<float *>(smem + G::O_T);          // over A_hi | A_lo: the first product has read them
#pragma unroll
            for (int c = 0; c < NX; c++) Ts[tid * G::FP + c] = y[c];
            __syncthreads();
            const float *col = Ts + (i_in_tile * NX) * G::FP + r;          // Y_i[k][c = r], k = 0 .. NX-1
#pragma unroll
            for (int k = 0; k < NX; k++) y[k] = col[k * G::FP];
            __syncthreads();                                               // every column is in registers: the scratch may go
#pragma unroll
            for (int kc = 0; kc < G::KC; kc++) {
                const float4 v = make_float4(y[kc * 4 + 0], y[kc * 4 + 1], y[kc * 4 + 2], y[kc * 4 + 3]);
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                const int off = kc * G::A_LBO + (tid >> 3) * 128 + (tid & 7) * 16;
                *reinterpret_cast<float4 *>(smem + G::O_AHI + off) = h;
                *reinterpret_cast<float4 *>(smem + G::O_ALO + off) = make_float4(tf32_lo(v.x, h.x), tf32_lo(v.y, h.y), tf32_lo(v.z, h.z), tf32_lo(v.w, h.w));
            }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        // ---- 4. D2 = (F P) F'  (rows (i,c), columns j)
        if (warp == 0) {
            if (tid == 0) { tc_fence_after(); issue_product(NX); }
            __syncwarp();
        }
        ok = mbar_wait(bar, phase); phase ^= 1;
        if (!ok) break;
        tc_fence_after();
        float pp[NX];                                   // this thread's row of the prior covariance P' = alpha^2 F P F' + Q
        {
            float v[NX];
            tmem_ld_row<NX>(lane_base + NX, v);
#pragma unroll
            for (int kc = 0; kc < G::KC; kc++) {
                const float4 q = *reinterpret_cast<const float4 *>(Qs + r * G::QP + kc * 4);
                pp[kc * 4 + 0] = fmaf(p.alpha_sq, v[kc * 4 + 0], q.x); pp[kc * 4 + 1] = fmaf(p.alpha_sq, v[kc * 4 + 1], q.y);
                pp[kc * 4 + 2] = fmaf(p.alpha_sq, v[kc * 4 + 2], q.z); pp[kc * 4 + 3] = fmaf(p.alpha_sq, v[kc * 4 + 3], q.w);
            }
        }
        if (live) {
            if (p.P_prior) {
                float4 *dst2 = reinterpret_cast<float4 *>(p.P_prior + row * NX);
#pragma unroll
                for (int kc = 0; kc < G::KC; kc++) dst2[kc] = make_float4(pp[kc * 4], pp[kc * 4 + 1], pp[kc * 4 + 2], pp[kc * 4 + 3]);
            }
            if (p.x_prior) p.x_prior[row] = xr;
        }
        float xo = xr;                                  // posterior := prior unless the update succeeds
        int st = BKE_STATUS_OK;
        if constexpr (M > 0) {
            // ---- 5. update (kalman_filter.py:533-556) with H, R shared, on the CUDA cores: everything is dim_z-sized per
            // filter and the filter's NX threads sit in one warp.  Row r of P' H' is a thin product (NX x dim_z FMAs per
            // thread, H broadcast from shared memory) — in plain fp32 from the row the thread already holds, so that K, S
            // and the covariance correction are mutually consistent to fp32 rounding
            float pht[M];
#pragma unroll
            for (int a = 0; a < M; a++) {
                float sacc = 0.f;
#pragma unroll
                for (int j = 0; j < NX; j++) sacc = fmaf(pp[j], Hs[a * NX + j], sacc);
                pht[a] = sacc;
            }
            const bool vz = live && (p.valid == nullptr || p.valid[f] != 0);
            // S = H (P' H') + R and H x' : this thread's terms, summed over the filter's NX lanes (xor butterfly)
            float Sm[M][M], hx[M];
#pragma unroll
            for (int a = 0; a < M; a++) {
                const float h = Hs[a * NX + r];
                hx[a] = h * xr;
#pragma unroll
                for (int b = 0; b < M; b++) Sm[a][b] = h * pht[b];
            }
#pragma unroll
            for (int o = NX / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int a = 0; a < M; a++) {
                    hx[a] += __shfl_xor_sync(FULL, hx[a], o);
#pragma unroll
                    for (int b = 0; b < M; b++) Sm[a][b] += __shfl_xor_sync(FULL, Sm[a][b], o);
                }
            }
            float yv[M];
#pragma unroll
            for (int a = 0; a < M; a++) {
                yv[a] = (vz ? p.z[f * M + a] : 0.f) - hx[a];
#pragma unroll
                for (int b = 0; b < M; b++) Sm[a][b] += Rs[a * M + b];
            }
            float SI[M][M], logdet;
            const bool inv_ok = reg_inverse<float, M>(Sm, SI, logdet);
            if (vz && !inv_ok) st = BKE_STATUS_SINGULAR_S;
            const bool upd = vz && inv_ok;
            float Kr[M];                                 // row r of K = P' H' S^-1
#pragma unroll
            for (int a = 0; a < M; a++) {
                float sk = 0.f;
#pragma unroll
                for (int b = 0; b < M; b++) sk += pht[b] * SI[b][a];
                Kr[a] = sk;
            }
            if (upd) {
#pragma unroll
                for (int a = 0; a < M; a++) xo += Kr[a] * yv[a];
            }
            // rows of K and P' H' of the filter's other lanes: through this thread's OWN first two operand chunks (only
            // its warp reads them, the second product is complete) — 16 bytes each, dim_z <= 4
            {
                float4 kq = make_float4(0.f, 0.f, 0.f, 0.f), pq = kq;
                float *kf4 = reinterpret_cast<float *>(&kq), *pf4 = reinterpret_cast<float *>(&pq);
#pragma unroll
                for (int a = 0; a < M; a++) { kf4[a] = Kr[a]; pf4[a] = pht[a]; }
                const int off = (tid >> 3) * 128 + (tid & 7) * 16;
                *reinterpret_cast<float4 *>(smem + G::O_AHI + off) = kq;
                *reinterpret_cast<float4 *>(smem + G::O_AHI + G::A_LBO + off) = pq;
            }
            __syncwarp();
            if (upd) {
                // Joseph form in the reference's order, with thin products (kalman_filter.py:555-556):
                //   T1 = (I - K H) P' = P' - K (P'H')'          row r: P'[r][j] - sum_a K[r][a] (P'H')[j][a]
                //   P  = T1 (I - K H)' + (K R) K'               row r: T1[r][j] + sum_a ((K R)[r][a] - (T1 H')[r][a]) K[j][a]
                // errors of K enter quadratically, as in the reference — not linearly as in P' - K S K'
                const int t0 = tid - r;                   // first thread of this filter
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    const int tj = t0 + j;
                    const float4 pq = *reinterpret_cast<const float4 *>(smem + G::O_AHI + G::A_LBO + (tj >> 3) * 128 + (tj & 7) * 16);
                    const float *pj = reinterpret_cast<const float *>(&pq);
                    float acc = pp[j];
#pragma unroll
                    for (int a = 0; a < M; a++) acc = fmaf(-Kr[a], pj[a], acc);
                    pp[j] = acc;
                }
                float g[M];                               // (K R)[r][a] - (T1 H')[r][a]
#pragma unroll
                for (int a = 0; a < M; a++) {
                    float t1h = 0.f, kr = 0.f;
#pragma unroll
                    for (int j = 0; j < NX; j++) t1h = fmaf(pp[j], Hs[a * NX + j], t1h);
#pragma unroll
                    for (int b = 0; b < M; b++) kr = fmaf(Kr[b], Rs[b * M + a], kr);
                    g[a] = kr - t1h;
                }
#pragma unroll
                for (int j = 0; j < NX; j++) {
                    const int tj = t0 + j;
                    const float4 kq = *reinterpret_cast<const float4 *>(smem + G::O_AHI + (tj >> 3) * 128 + (tj & 7) * 16);
                    const float *kj = reinterpret_cast<const float *>(&kq);
                    float acc = pp[j];
#pragma unroll
                    for (int a = 0; a < M; a++) acc = fmaf(g[a], kj[a], acc);
                    pp[j] = acc;
                }
            }
            __syncwarp();                               // the chunks are rewritten by the next tile's operand rows
            if (live) {
                // optional outputs (kalman_filter.py:533-544 attributes): y always (0 when z is None), S when there is a
                // measurement, K / SI / log-likelihood when S was invertible; otherwise the arrays keep their values
                if (p.y && r == 0) {
#pragma unroll
                    for (int a = 0; a < M; a++) p.y[f * M + a] = vz ? yv[a] : 0.f;
                }
                if (vz && r == 0 && p.S) {
#pragma unroll
                    for (int a = 0; a < M; a++)
#pragma unroll
                        for (int b = 0; b < M; b++) p.S[f * M * M + a * M + b] = Sm[a][b];
                }
                if (upd) {
                    if (p.K) {
#pragma unroll
                        for (int a = 0; a < M; a++) p.K[(f * NX + r) * M + a] = Kr[a];
                    }
                    if (r == 0) {
                        if (p.SI) {
#pragma unroll
                            for (int a = 0; a < M; a++)
#pragma unroll
                                for (int b = 0; b < M; b++) p.SI[f * M * M + a * M + b] = SI[a][b];
                        }
                        if (p.ll) {
                            float q = 0.f;
#pragma unroll
                            for (int a = 0; a < M; a++) {
                                float sq = 0.f;
#pragma unroll
                                for (int b = 0; b < M; b++) sq += SI[a][b] * yv[b];
                                q += yv[a] * sq;
                            }
                            p.ll[f] = -0.5f * (q + logdet + float(M) * float(LOG_2PI));
                        }
                    }
                }
            }
        }
        if (live) {
            float4 *dst = reinterpret_cast<float4 *>(p.P_out + row * NX);
#pragma unroll
            for (int kc = 0; kc < G::KC; kc++) dst[kc] = make_float4(pp[kc * 4], pp[kc * 4 + 1], pp[kc * 4 + 2], pp[kc * 4 + 3]);
            // every thread of the filter has read x (staged before the first barrier of this tile): in place is safe
            p.x_out[row] = xo;
            if (p.status && r == 0 && (st != BKE_STATUS_OK || !p.sticky)) p.status[f] = st;
        }
        // no barrier here: the operand buffers are free (the second product's commit was awaited), the next tile's
        // products are issued only after barriers every warp reaches after these TMEM reads, and xs was read before
        // this tile's second barrier
        tc_fence_before();
    }
    if (!ok && p.err) *p.err = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)G::TMEM_COLS) : "memory");
    }
}

template <int NX, int M>
int launch_t(const TcP &p, cudaStream_t s)
{
    using G = Geom<NX>;
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        if (check_cuda(cudaFuncSetAttribute(kf_cov_tc_kernel<NX, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM), "cudaFuncSetAttribute")) return BKE_ERR_CUDA;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int64_t tiles = (p.N * NX + 127) / 128;
    // CTAs per SM: registers (launch bounds), shared memory (+1 KB the runtime reserves per CTA) and TMEM columns;
    // BKE_KF_TC_CTAS overrides (tuning)
    static const int env_ctas = [] { const char *e = getenv("BKE_KF_TC_CTAS"); return e ? atoi(e) : 0; }();
    int occ = tc_ctas(NX, M);
    if (occ > (227 * 1024) / (G::SMEM + 1024)) occ = (227 * 1024) / (G::SMEM + 1024);
    if (occ > 512 / G::TMEM_COLS) occ = 512 / G::TMEM_COLS;
    if (env_ctas > 0 && env_ctas < occ) occ = env_ctas;
    const int64_t cap = (int64_t)sm_count() * occ;
    kf_cov_tc_kernel<NX, M><<<(unsigned)(tiles < cap ? tiles : cap), 128, G::SMEM, s>>>(p);
    return check_cuda(cudaGetLastError(), "kf_cov_tc_kernel launch");
}

template <int NX>
int launch_m(const TcP &p, int m, cudaStream_t s)
{
    switch (m) {
    case 0: return launch_t<NX, 0>(p, s);
    case 1: return launch_t<NX, 1>(p, s);
    case 2: return launch_t<NX, 2>(p, s);
    case 3: return launch_t<NX, 3>(p, s);
    default: return launch_t<NX, 4>(p, s);
    }
}

}  // namespace tc

static bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// BKE_KF_TC=0 keeps the CUDA-core kernels for these shapes (A/B measurements).  Eligible: fp32, dim_x 16 / 32, F and Q
// shared, no control input.  A fused predict+update whose H and R are shared too and whose dim_z <= 4 runs entirely in
// this kernel (BKE_KF_TC_FUSED=0: predict here, update through the usual order); any other fused step runs its predict
// here and its update through launch_kf_any on the prior left in x_out / P_out.
int launch_kf_tc(const bke_kf_args &a, cudaStream_t s)
{
    static const int mode = [] { const char *e = getenv("BKE_KF_TC"); return e ? atoi(e) : 1; }();
    static const bool fused_on = [] { const char *e = getenv("BKE_KF_TC_FUSED"); return !(e && e[0] == '0'); }();
    if (mode == 0) return BKE_ERR_UNSUPPORTED;
    if (a.dtype != BKE_F32 || !(a.dim_x == 16 || a.dim_x == 32)) return BKE_ERR_UNSUPPORTED;
    if (!(a.flags & BKE_DO_PREDICT) || (a.flags & BKE_UPDATE_FIRST)) return BKE_ERR_UNSUPPORTED;
    if (a.F_stride != 0 || a.Q_stride != 0 || (a.B && a.u)) return BKE_ERR_UNSUPPORTED;
    if (!(al16(a.x) && al16(a.P) && al16(a.x_out) && al16(a.P_out) && al16(a.x_prior) && al16(a.P_prior))) return BKE_ERR_UNSUPPORTED;
    const bool fused = (a.flags & BKE_DO_UPDATE) != 0;
    const bool fused_here = fused && fused_on && a.H_stride == 0 && a.R_stride == 0 && a.dim_z >= 1 && a.dim_z <= 4 && a.z;
    // without the fused kernel the 16/4 and 16/2 fused steps are faster as ONE row-block launch than as predict-here + update-there
    if (fused && !fused_here && mode == 1 && a.dim_x == 16 && (a.dim_z == 4 || a.dim_z == 2)) return BKE_ERR_UNSUPPORTED;
    // x_out may alias x and P_out may alias P (each tile reads its rows before it writes them); nothing else may overlap
    tc::TcP p;
    p.N = a.n_filters; p.alpha_sq = (float)a.alpha_sq;
    p.x = (const float *)a.x; p.P = (const float *)a.P; p.F = (const float *)a.F; p.Q = (const float *)a.Q;
    p.x_out = (float *)a.x_out; p.P_out = (float *)a.P_out; p.x_prior = (float *)a.x_prior; p.P_prior = (float *)a.P_prior;
    p.status = (fused && !fused_here) ? nullptr : a.status;      // the update that follows owns the status of a two-launch step
    p.sticky = (a.flags & BKE_STATUS_STICKY) ? 1 : 0;
    p.err = nullptr;
    p.H = (const float *)a.H; p.R = (const float *)a.R; p.z = (const float *)a.z; p.valid = a.z_valid;
    p.K = (float *)a.K; p.y = (float *)a.y; p.S = (float *)a.S; p.SI = (float *)a.SI; p.ll = (float *)a.log_likelihood;
    const int m_here = fused_here ? a.dim_z : 0;
    int rc = a.dim_x == 16 ? tc::launch_m<16>(p, m_here, s) : tc::launch_m<32>(p, m_here, s);
    if (rc != BKE_OK || !fused || fused_here) return rc;
    // two-launch fused step: the update runs on the prior this launch left in x_out / P_out (stream order)
    bke_kf_args u = a;
    u.flags = (a.flags & ~(uint32_t)BKE_DO_PREDICT);
    u.x = a.x_out; u.P = a.P_out;
    u.x_prior = nullptr; u.P_prior = nullptr;         // written above
    return launch_kf_any(u, s);
}

}  // namespace bke
