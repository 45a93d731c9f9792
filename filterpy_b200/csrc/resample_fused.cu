// resample_fused.cu — systematic / stratified resampling and the exact cumulative sum as ONE
// single-pass kernel: every weight is read from HBM once (8 B in, 4 B out per particle).
//
//   indexes[i] = #{ j : c_j <= pos_i },  c_j = fl(c_{j-1} + w_j)   (filterpy/monte_carlo/resampling.py:141-149)
//   pos_i = fl(fl(u + i) / N) (:139)  or  fl(fl(U_i + i) / N) (:103)
//
// The running sum is reproduced EXACTLY (resample_common.cuh: inside one binade adding w is the
// integer map bits(S) -> bits(S) + d[parity]); what is new here is how the tiles are chained:
// a decoupled look-back in TWO stages, both on one 64-bit status word per tile.
//
//   stage 1 (producer warp)  approximate fp64 tile sum -> approximate prefix `tp`  (which binade
//            the tile lives in, and which adds might leave it)
//   stage 2 (consumer warps) the tile's parity map D computed in that binade -> exact state S_in
//            (bit pattern of the reference's running sum before the tile), then every c_j,
//            the output range of every particle and the index expansion.
//
// CTA = NW consumer warps + 1 producer warp, persistent, tiles handed out by an atomic counter
// (so a tile only ever waits for tiles that are already running).  The producer warp claims the
// next tile, pulls its 16*NW*32 weights into shared memory with ONE 2-D TMA copy (128-byte
// swizzle: thread t reads its 16 consecutive weights with conflict-free LDS.128), sums them,
// publishes / resolves stage 1 and hands the stage over while the consumers still work on the
// previous tile: HBM latency and the first look-back are off the consumers' critical path.
// A tile that is not "clean" (a possible binade crossing, ties) takes slow_tile(): the old
// raw-element / segment walk, started before S_in arrives so that only the short walk is serial.
//
// Expansion: each particle with >= 1 copies stores (local index + 1) at its first output slot of
// a zeroed shared-memory window; a max-scan over the slots fills the runs (no divergent copy
// loop), and every thread leaves with 16-byte stores of 20 consecutive indexes.
//
// Everything is verified with the exact values (tile start / end inside the assumed binade); a
// failed check or a negative / non-finite weight switches to the literal sequential kernel.
#include <cuda.h>
#include <stdlib.h>
#include "resample_common.cuh"
#include "resample_fused.cuh"

namespace bke {
namespace rs {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t f_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void f_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void f_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void f_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool f_mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(f_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void f_mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!f_mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void f_tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(f_smem_u32(dst)), "l"(map), "r"(f_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// named barriers of the consumer warps (the producer warp never joins them)
template <int NT> __device__ __forceinline__ void f_bar()
{
    asm volatile("barrier.cta.sync 1, %0;" ::"n"(NT) : "memory");
}
template <int NT> __device__ __forceinline__ int f_bar_and(int pred)
{
    int out;
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 p, %1, 0;\nbarrier.cta.red.and.pred q, 1, %2, p;\nselp.b32 %0, 1, 0, q;\n}\n"
                 : "=r"(out) : "r"(pred), "n"(NT) : "memory");
    return out;
}
__device__ __forceinline__ u64 f_ld(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void f_st(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ------------------------------------------------------------------ event trace (debugging aid)
// bke_debug_resample_trace(buf): every tile writes the global-timer time of 10 pipeline events
__device__ __forceinline__ void f_trace(const FParams &p, int t, int ev)
{
    if (p.trace) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        p.trace[(size_t)t * 10 + ev] = now;
    }
}

// ------------------------------------------------------------------ status words
// st1[i], st2[i] describe tile i-1; entry 0 is the carry into the call (always INCLUSIVE).
//   st1: bits of an fp64 sum with the two lowest mantissa bits replaced by the flag
//        (1 = this tile's sum, 2 = sum of everything up to and including this tile)
//   st2: bit 63 = INCLUSIVE: bits 0..62 = exact running sum after the tile (a non-negative double)
//        bit 62 = AGGREGATE: bits 0..59 = d0 of the tile's parity map, bits 60..61 = t + 1
constexpr u64 ST1_AGG = 1, ST1_INCL = 2;
constexpr u64 ST2_INCL = 1ull << 63, ST2_AGG = 1ull << 62;
constexpr int SPIN_LIMIT = 1 << 22;

__device__ __forceinline__ u64 st2_pack_agg(i64 d, int t) { return ST2_AGG | ((u64)(t + 1) << 60) | ((u64)d & ((1ull << 60) - 1)); }

// Look-back windows: every round a warp fetches up to 32 * LBK_MAX status words at once (ONE L2 round
// trip), walks them in groups of 32 from the nearest predecessor outwards and stops at the nearest
// INCLUSIVE word.  All tiles of a persistent grid start together, so a tile is typically a few
// hundred tiles ahead of the inclusive frontier: the width of the window, not the number of
// resident CTAs, sets how many round trips a look-back costs.
constexpr int LBK_MAX = 8;

// a blocked round (an unpublished word in front of the nearest inclusive one): back off, and give up
// after SPIN_LIMIT rounds or once any look-back has given up (the result then comes from the fallback)
__device__ __forceinline__ bool f_blocked(FHeader *hdr, int &spins, int sleep_ns)
{
    if (sleep_ns > 0) __nanosleep(sleep_ns);
    spins++;
    if ((spins & 1023) == 0 && *reinterpret_cast<volatile int *>(&hdr->timeout)) return true;
    if (spins >= SPIN_LIMIT) { hdr->timeout = 1; hdr->fallback = 1; return true; }
    return false;
}

// stage 1: approximate sum of everything before tile t (all 32 lanes of the producer warp)
__device__ __forceinline__ double f_lookback_sum(const FParams &p, int t, int lane)
{
    double part = 0.0;                                     // this lane's share; reduced once at the end
    int idx = t - lane;                                    // st1 index of tile t-1-lane (group 0 of the round)
    int spins = 0;
    bool done = false;
    while (!done) {
        u64 v[LBK_MAX];
#pragma unroll
        for (int j = 0; j < LBK_MAX; j++) {
            const int i = idx - 32 * j;
            v[j] = ST1_INCL;                               // beyond the carry: 0.0, inclusive
            if (j < p.lbk && i >= 0) v[j] = f_ld(p.st1 + i);
        }
        bool blocked = false;
#pragma unroll
        for (int j = 0; j < LBK_MAX; j++) {
            if (j < p.lbk && !done && !blocked) {
                const unsigned incl = __ballot_sync(FULL, (v[j] & 3) == ST1_INCL);
                const unsigned empty = __ballot_sync(FULL, (v[j] & 3) == 0);
                const int first = incl ? __ffs(incl) - 1 : 32;
                const unsigned closer = first >= 32 ? FULL : ((1u << first) - 1u);
                if (empty & closer) blocked = true;
                else {
                    if (lane <= first) part += __longlong_as_double((i64)(v[j] & ~3ull));
                    if (incl) done = true; else idx -= 32;
                }
            }
        }
        if (blocked && f_blocked(p.hdr, spins, p.sleep_ns)) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
    return part;
}

// stage 2: exact state before tile t (all 32 lanes of a warp).  The aggregates between the nearest
// inclusive predecessor and this tile are composed in tile order; with tie-free maps (t = 0, the
// normal case) that is a plain int64 sum, kept per lane and reduced once.
__device__ __forceinline__ i64 f_lookback_state(const FParams &p, int t, int lane)
{
    i64 part = 0;                                          // tie-free mode: this lane's share of the sum
    bool general = false;                                  // a map with ties was met: ordered composition from here on
    i64 acc_d = 0; int acc_t = 0;                          // general mode: composite of everything nearer (applied LAST)
    i64 S = 0;
    int idx = t - lane;
    int spins = 0;
    bool done = false;
    while (!done) {
        u64 v[LBK_MAX];
#pragma unroll
        for (int j = 0; j < LBK_MAX; j++) {
            const int i = idx - 32 * j;
            v[j] = ST2_INCL;
            if (j < p.lbk && i >= 0) v[j] = f_ld(p.st2 + i);
        }
        bool blocked = false;
#pragma unroll
        for (int j = 0; j < LBK_MAX; j++) {
            if (j < p.lbk && !done && !blocked) {
                const unsigned incl = __ballot_sync(FULL, (v[j] & ST2_INCL) != 0);
                const unsigned empty = __ballot_sync(FULL, (v[j] & (ST2_INCL | ST2_AGG)) == 0);
                const int first = incl ? __ffs(incl) - 1 : 32;
                const unsigned closer = first >= 32 ? FULL : ((1u << first) - 1u);
                if (empty & closer) { blocked = true; continue; }
                i64 d = 0; int tt = 0;
                if (lane < first) { d = (i64)(v[j] & ((1ull << 60) - 1)); tt = (int)((v[j] >> 60) & 3) - 1; }
                const unsigned ties = __ballot_sync(FULL, tt != 0);
                if (!general && ties == 0) part += d;
                else {
                    if (!general) {
                        general = true;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
                        acc_d = part; acc_t = 0;
                    }
                    // ordered: lane L holds tile (..)-L, i.e. higher lanes are applied first
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const i64 pd = __shfl_down_sync(FULL, d, o);
                        const int pt = __shfl_down_sync(FULL, tt, o);
                        if (lane + o < 32) {
                            const SM r = combine(SM{pd, pt, 0, K_ID}, SM{d, tt, 0, K_ID});
                            d = r.d; tt = r.t;
                        }
                    }
                    d = __shfl_sync(FULL, d, 0); tt = __shfl_sync(FULL, tt, 0);
                    const SM r = combine(SM{d, tt, 0, K_ID}, SM{acc_d, acc_t, 0, K_ID});   // this group lies before everything nearer
                    acc_d = r.d; acc_t = r.t;
                }
                if (incl) { S = (i64)(__shfl_sync(FULL, v[j], first) & ~ST2_INCL); done = true; }
                else idx -= 32;
            }
        }
        if (blocked && f_blocked(p.hdr, spins, p.sleep_ns)) break;
    }
    if (!general) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        acc_d = part; acc_t = 0;
    }
    return S + acc_d + ((S & 1) ? acc_t : 0);
}

// ------------------------------------------------------------------ shared memory
enum { TM_FAST = 0, TM_SLOW = 1, TM_BAD = 2 };                       // what the consumers do with a tile
enum { TK_CLEAN = 0, TK_CROSS = 1, TK_TIES = 2, TK_BAD = 3 };        // what stage 1 found out about it
constexpr int MAX_CROSS = 24;                                        // crossing rows resolved by the chain warp

template <int NW, int STAGES>
struct FSmem {
    static constexpr int NT = NW * 32, TILE = NT * F_IPT, WIN = NT * F_SPT;
    double w[STAGES][TILE];            // TMA destinations (128-byte swizzle): must stay first, 1024-aligned
    i64 ex[STAGES][NT];                // fast tiles: parity-map offset of every consumer thread's first particle
    int win[WIN];                      // output window (all zero between tiles)
    double warp_d[NW];
    SM warp_sm[NW];
    int warp_max[NW];
    unsigned cross[STAGES][NT / 32];   // rows (consumer threads) whose 16 adds leave their binade: true adds there
    uint64_t full_tma[STAGES], landed[STAGES], summed[STAGES], mapped[STAGES], ready[STAGES], empty[STAGES];
    struct Part { double tot; int bad; int tie; } part[STAGES][2];                 // the two loaders' halves
    int lcnt[STAGES];                  // loaders done with the tile (the second one publishes)
    int eg[STAGES];                    // binade the loaders' speculative row sums were formed in
    struct Pre { double tot; int bad; int eg; int tie; int pad; } pre[STAGES];     // loaders -> C1
    struct Mid { int kind; int e0; double tp; double tot; i64 D; } mid[STAGES];     // C1 -> C2
    int last_t;                        // trace: the tile the consumers processed last
    int e_last;                        // binade of the last fast tile the chain warp resolved (the loader's guess)
    struct Info { int t; int mode; int good; int pad; double tp; i64 S_in; i64 base; i64 lo; i64 cnt; } info[STAGES];
    int tile_of[STAGES];               // producer-private: tile loaded / loading in each stage (-1 = end of work)
    i64 bc_S_in, bc_lo, bc_cnt;
    int bc_ok, bc_skip;
    // slow path (tiles with raw elements)
    i64 segstate[RMAX + 1];
    i64 segd[RMAX + 1];
    double wraw[RMAX];
    int segk[RMAX + 1];
    int segt[RMAX + 1];
    int first_raw[NT + 1];
    i64 tstart[NT];
};

// byte offset of weight (row r = owning thread, 16-byte chunk c) inside a swizzled stage
__device__ __forceinline__ uint32_t f_swz(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

template <int NW> __device__ __forceinline__ double f_scan_d(double v, double *total, double *sh, int lane, int wid)
{
    const double inc = warp_incl_scan_d(v, lane);
    if (lane == 31) sh[wid] = inc;
    f_bar<NW * 32>();
    double base = 0.0, tot = 0.0;
#pragma unroll
    for (int i = 0; i < NW; i++) { const double x = sh[i]; if (i < wid) base += x; tot += x; }
    f_bar<NW * 32>();
    *total = tot;
    return base + (inc - v);
}

template <int NW> __device__ __forceinline__ SM f_scan_sm(SM v, SM *total, SM *sh, int lane, int wid)
{
    const SM inc = warp_incl_scan_sm(v, lane);
    if (lane == 31) sh[wid] = inc;
    f_bar<NW * 32>();
    SM base = sm_identity(), tot = sm_identity();
    for (int i = 0; i < NW; i++) { const SM x = sh[i]; if (i < wid) base = combine(base, x); tot = combine(tot, x); }
    f_bar<NW * 32>();
    *total = tot;
    SM prev = shfl_up_sm(inc, 1);
    if (lane == 0) prev = sm_identity();
    return combine(base, prev);
}

template <int MODE>
__device__ __forceinline__ i64 f_count_below(const FParams &p, double c)
{
    const double Ngd = (double)p.ng;
    return (MODE == F_STRAT) ? count_below_str(c, p.U, p.ng, Ngd) : count_below_sys(c, p.u, p.ng, Ngd, p.tau);
}

__device__ __forceinline__ void f_put_index(const FParams &p, i64 out_begin, i64 o, int value)
{
    const i64 rel = o - out_begin;
    if (rel >= 0 && rel < p.cap) p.idx[rel] = value;
    else p.hdr->cap_overflow = 1;
}

// one lane, after the tile's exact end state is known and published: the tile's output range
// [lo, lo + cnt) and, for the last tile, the call's bookkeeping
template <int MODE>
__device__ __forceinline__ void f_finish_tile(const FParams &p, int t, i64 S_in, i64 S_out, int &good, i64 &lo, i64 &cnt)
{
    lo = 0; cnt = 0;
    if (MODE != F_CUMSUM && good) {
        lo = f_count_below<MODE>(p, __longlong_as_double(S_in));
        cnt = f_count_below<MODE>(p, __longlong_as_double(S_out)) - lo;
        if (cnt < 0) { cnt = 0; good = 0; }
    }
    if (!good) { p.hdr->fallback = 1; p.hdr->chain_bad = 1; }
    if (t == p.T - 1) {
        if (p.cumsum_last) *p.cumsum_last = __longlong_as_double(S_out);
        if (MODE != F_CUMSUM) {
            i64 O1 = lo + cnt;
            if (p.is_last && O1 < p.ng) {                  // resampling.py:145 would raise IndexError
                p.hdr->overflow = (int)(p.ng - O1 > 0x7fffffff ? 0x7fffffff : p.ng - O1);
                const int r = atomicAdd(&p.hdr->n_runs, 1);
                if (r < p.max_runs) p.runs[r] = Run{O1, p.ng, (int)(p.ng - 1), 0};
                O1 = p.ng;
            }
            p.hdr->out_end = O1;
            if (p.out_range) { p.out_range[0] = p.hdr->out_begin; p.out_range[1] = O1; }
        }
    }
}

// ------------------------------------------------------------------ slow path: ties, raw elements, dense raw zones
// Leaves the exact c_j (bit patterns) of the tile in the stage buffer, at the positions of the
// weights they belong to; returns 1 if the tile verified.  Everything up to the segment export
// runs BEFORE the exact start state is known; only the walk over <= RMAX segments is serial.
template <int NW, int STAGES, int MODE>
__device__ __noinline__ int f_slow_tile(const FParams &p, FSmem<NW, STAGES> &sm, int s, int t)
{
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned char *sb = reinterpret_cast<unsigned char *>(sm.w[s]);
    double w[F_IPT];
#pragma unroll
    for (int c = 0; c < F_IPT / 2; c++) {
        const double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(tid, c));
        w[2 * c] = v.x; w[2 * c + 1] = v.y;
    }
    double ssum = 0.0;
#pragma unroll
    for (int k = 0; k < F_IPT; k++) ssum += w[k];
    double tot;
    double before = sm.info[s].tp + f_scan_d<NW>(ssum, &tot, sm.warp_d, lane, wid);
    SM inc[F_IPT];
    int ek[F_IPT];
    SM run = sm_identity();
#pragma unroll
    for (int k = 0; k < F_IPT; k++) {
        const double after = before + w[k];
        SM el;
        if (w[k] == 0.0) { el = sm_identity(); ek[k] = K_ID; }          // fl(S + 0) = S in every binade
        else {
            int e;
            if (clean_add(before, after, p.eb, &e)) { el = elem_map(w[k], e); ek[k] = e; }
            else { el = SM{0, 0, 1, K_ID}; ek[k] = -1; }
        }
        run = combine(run, el);
        inc[k] = run;
        before = after;
    }
    SM total;
    const SM excl = f_scan_sm<NW>(run, &total, sm.warp_sm, lane, wid);
#pragma unroll
    for (int k = 0; k < F_IPT; k++) inc[k] = combine(excl, inc[k]);
    const int nraw = total.cnt;
    int bad = 0;
    if (nraw > 0 && nraw <= RMAX) {
        sm.first_raw[tid] = (ek[0] == -1);
        if (tid == 0) sm.first_raw[NT] = 1;                 // the tile end closes the last segment
        for (int q = tid; q <= RMAX; q += NT) { sm.segk[q] = -1; sm.segt[q] = 0; sm.segd[q] = 0; }
        f_bar<NT>();
#pragma unroll
        for (int k = 0; k < F_IPT; k++) {
            const int seg = inc[k].cnt;
            if (ek[k] == -1) {
                sm.wraw[seg - 1] = w[k];                    // the raw element that opens segment `seg`
            } else {
                const bool next_raw = (k + 1 < F_IPT) ? (ek[k + 1] == -1) : (sm.first_raw[tid + 1] != 0);
                if (next_raw) {
                    if (inc[k].k == K_POISON) bad = 1;
                    sm.segk[seg] = inc[k].k == K_ID ? -1 : inc[k].k;    // identity segments are skipped
                    sm.segt[seg] = inc[k].t;
                    sm.segd[seg] = inc[k].d;
                }
            }
        }
    }
    f_bar<NT>();
    if (wid == 0) {
        const i64 S_in = sm.info[s].S_in;                  // chain warp C2 resolved it before handing the tile over
        if (lane == 0) {
            int wbad = 0;
            i64 S_out;
            if (nraw == 0) {
                S_out = total.k >= 0 ? apply_bits(S_in, total.d, total.t, total.k, &wbad) : S_in;
                if (total.k == K_POISON) wbad = 1;
                sm.segstate[0] = S_in;
            } else if (nraw <= RMAX) {
                i64 S = S_in;
                for (int q = 0; q <= nraw; q++) {
                    sm.segstate[q] = S;
                    if (sm.segk[q] >= 0) S = apply_bits(S, sm.segd[q], sm.segt[q], sm.segk[q], &wbad);
                    if (q < nraw) S = __double_as_longlong(__dadd_rn(__longlong_as_double(S), sm.wraw[q]));
                }
                S_out = S;
                atomicAdd(&p.hdr->n_unclean, 1);
            } else {
                // a dense zone of raw elements (tiny weights next to a binade boundary): true adds, one by one
                double acc = __longlong_as_double(S_in);
                for (int th = 0; th < NT; th++) {
                    sm.tstart[th] = __double_as_longlong(acc);
#pragma unroll
                    for (int c = 0; c < F_IPT / 2; c++) {
                        const double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(th, c));
                        acc = __dadd_rn(acc, v.x);
                        acc = __dadd_rn(acc, v.y);
                    }
                }
                S_out = __double_as_longlong(acc);
                atomicAdd(&p.hdr->n_seq, 1);
            }
            f_st(p.st2 + t + 1, ST2_INCL | (u64)S_out);
            atomicAdd(&p.hdr->n_slow, 1);
            int good = !wbad;
            i64 lo, cnt;
            f_finish_tile<MODE>(p, t, S_in, S_out, good, lo, cnt);
            sm.bc_S_in = S_in; sm.bc_lo = lo; sm.bc_cnt = cnt; sm.bc_ok = good;
        }
    }
    f_bar<NT>();
    i64 cb[F_IPT];
    if (nraw > RMAX) {
        double acc = __longlong_as_double(sm.tstart[tid]);
#pragma unroll
        for (int k = 0; k < F_IPT; k++) { acc = __dadd_rn(acc, w[k]); cb[k] = __double_as_longlong(acc); }
    } else {
#pragma unroll
        for (int k = 0; k < F_IPT; k++) {
            const i64 S0 = sm.segstate[inc[k].cnt];
            if (inc[k].k == K_POISON) bad = 1;
            // a raw element: the segment it opens starts at its own result; only zeros so far: unchanged
            cb[k] = (ek[k] == -1 || inc[k].k < 0) ? S0 : apply_bits(S0, inc[k].d, inc[k].t, inc[k].k, &bad);
        }
    }
    if (bad) { p.hdr->fallback = 1; p.hdr->chain_bad = 2; }
#pragma unroll
    for (int c = 0; c < F_IPT / 2; c++)
        *reinterpret_cast<longlong2 *>(sb + f_swz(tid, c)) = make_longlong2(cb[2 * c], cb[2 * c + 1]);
    f_fence_proxy_async();                // generic-proxy writes to a stage the TMA engine will refill
    return f_bar_and<NT>(!bad && sm.bc_ok);
}

// cumsum mode: the exact running sums leave as they are (each thread owns 16 consecutive elements = one 128-byte line)
__device__ __forceinline__ void f_store_cumsum(const FParams &p, i64 j, const i64 (&cb)[F_IPT])
{
    double *o = p.cumsum_out + j;
    if (j + F_IPT <= p.n && (reinterpret_cast<uintptr_t>(o) & 15) == 0 && !(p.last_one && j + F_IPT == p.n)) {
#pragma unroll
        for (int k = 0; k < F_IPT; k += 2)
            *reinterpret_cast<double2 *>(o + k) = make_double2(__longlong_as_double(cb[k]), __longlong_as_double(cb[k + 1]));
    } else {
#pragma unroll
        for (int k = 0; k < F_IPT; k++)
            if (j + k < p.n) o[k] = (p.last_one && j + k == p.n - 1) ? 1.0 : __longlong_as_double(cb[k]);
    }
}

// read this thread's F_SPT window slots (and clear them), running maximum, block max-scan:
// m[i] = marker (local particle index + 1) of the particle that owns slot tid*F_SPT + i
template <int NW, int STAGES>
__device__ __forceinline__ void f_window_scan(FSmem<NW, STAGES> &sm, int tid, int lane, int wid, int (&m)[F_SPT])
{
    constexpr int NT = NW * 32;
    int4 *wv = reinterpret_cast<int4 *>(sm.win) + tid * (F_SPT / 4);
#pragma unroll
    for (int i = 0; i < F_SPT / 4; i++) {
        const int4 q = wv[i];
        m[4 * i] = q.x; m[4 * i + 1] = q.y; m[4 * i + 2] = q.z; m[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int i = 0; i < F_SPT / 4; i++) wv[i] = make_int4(0, 0, 0, 0);
#pragma unroll
    for (int i = 1; i < F_SPT; i++) m[i] = max(m[i], m[i - 1]);
    int incm = m[F_SPT - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incm, o); if (lane >= o) incm = max(incm, y); }
    int basem = __shfl_up_sync(FULL, incm, 1);
    if (lane == 0) basem = 0;
    if (lane == 31) sm.warp_max[wid] = incm;
    f_bar<NT>();
#pragma unroll
    for (int i = 0; i < NW; i++) { const int x = sm.warp_max[i]; if (i < wid) basem = max(basem, x); }
#pragma unroll
    for (int i = 0; i < F_SPT; i++) m[i] = max(m[i], basem);
}

// Row sums of the tie-free parity map in binade e for rows >= r0 (lane L owns rows L, L + 32, ...:
// rs[i] belongs to row 32 i + L; 0 for rows < r0); returns true if any of those rows holds an exact tie.
template <int NW, int STAGES>
__device__ __forceinline__ bool f_row_sums(FSmem<NW, STAGES> &sm, int s, int e, int r0, int lane, i64 (&rs)[NW])
{
    constexpr int RPL = NW;                                // rows per lane (NT / 32)
    const unsigned char *sb = reinterpret_cast<const unsigned char *>(sm.w[s]);
    const i64 base = (i64)e << 52;
    const double B0 = __longlong_as_double(base), B1 = __longlong_as_double(base + 1);
    unsigned tie = 0;
#pragma unroll
    for (int i = 0; i < RPL; i++) {
        const int r = i * 32 + lane;
        i64 racc = 0;
        if (r >= r0) {
#pragma unroll
            for (int c = 0; c < F_IPT / 2; c++) {
                const double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(r, c));
                const i64 x0 = __double_as_longlong(__dadd_rn(B0, v.x)), x1 = __double_as_longlong(__dadd_rn(B1, v.x));
                const i64 y0 = __double_as_longlong(__dadd_rn(B0, v.y)), y1 = __double_as_longlong(__dadd_rn(B1, v.y));
                tie |= (((unsigned)x0 + 1u) ^ (unsigned)x1) | (((unsigned)y0 + 1u) ^ (unsigned)y1);   // d1 != d0: an exact tie
                racc += (x0 - base) + (y0 - base);
            }
        }
        rs[i] = racc;
    }
    return __any_sync(FULL, tie != 0) != 0;
}

// A tile that may leave its binade, resolved from the EXACT state before it (chain warp C2):
// rows whose adds all stay in the binade of their start state are integer maps; the row that
// leaves it is walked with true adds; the rows after it are maps of the next binade, and so on.
// On success sm.ex[s][r] = exact state before row r, sm.cross[s] marks the walked rows, *S_out =
// state after the tile.  false: ties or more than MAX_CROSS crossings (the consumers' general path).
template <int NW, int STAGES>
__device__ __noinline__ bool f_resolve_exact(FSmem<NW, STAGES> &sm, int s, i64 S_in, i64 *S_out, int lane)
{
    constexpr int NT = NW * 32, RPL = NW;
    const unsigned char *sb = reinterpret_cast<const unsigned char *>(sm.w[s]);
    int r0 = 0;
    i64 S = S_in;
    for (int round = 0; r0 < NT; round++) {
        if (round > MAX_CROSS) return false;
        const int e = (int)(S >> 52);
        i64 rsum[RPL];
        if (f_row_sums<NW, STAGES>(sm, s, e, r0, lane, rsum)) return false;
        i64 carry = S;                                     // state before row 32 i (rows < r0 contribute 0)
        int found = -1;
        i64 S_row = 0;
#pragma unroll
        for (int i = 0; i < RPL; i++) {
            if (found >= 0) continue;
            const int r = i * 32 + lane;
            const i64 rs = rsum[i];
            const i64 inc = warp_incl_scan_i64(rs, lane);
            const i64 end_state = carry + inc, start_state = end_state - rs;
            const bool leaves = r >= r0 && (int)(end_state >> 52) != e;
            const unsigned m = __ballot_sync(FULL, leaves);
            const int fc = m ? __ffs(m) - 1 : 32;
            if (r >= r0 && lane <= fc) sm.ex[s][r] = start_state;      // rows up to and including the leaving one start here
            if (m) { found = i * 32 + fc; S_row = __shfl_sync(FULL, start_state, fc); }
            carry = __shfl_sync(FULL, end_state, 31);
        }
        __syncwarp();
        if (found < 0) { *S_out = carry; return true; }
        // the row that leaves the binade: true adds, one by one
        double acc = __longlong_as_double(S_row);
#pragma unroll
        for (int c = 0; c < F_IPT / 2; c++) {
            const double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(found, c));
            acc = __dadd_rn(acc, v.x);
            acc = __dadd_rn(acc, v.y);
        }
        if (lane == 0) sm.cross[s][found >> 5] |= 1u << (found & 31);
        __syncwarp();
        S = __double_as_longlong(acc);
        r0 = found + 1;
    }
    *S_out = S;
    return true;
}

// ------------------------------------------------------------------ the kernel
// resident CTAs per SM that the shared-memory footprint of a variant allows (227 KB per SM)
constexpr int f_ctas(int nw, int stages) { return nw == 8 ? (stages <= 2 ? 2 : 1) : (stages <= 2 ? 4 : (stages == 3 ? 3 : 2)); }

template <int NW, int STAGES, int MODE>
__global__ void __launch_bounds__(NW * 32 + 128, f_ctas(NW, STAGES))
k_fused(const __grid_constant__ CUtensorMap wmap, const FParams p)
{
    constexpr int NT = NW * 32, TILE = NT * F_IPT, WIN = NT * F_SPT, RPL = NT / 32;
    extern __shared__ __align__(1024) unsigned char f_smem_raw[];
    FSmem<NW, STAGES> &sm = *reinterpret_cast<FSmem<NW, STAGES> *>(f_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            f_mbar_init(&sm.full_tma[s], 1); f_mbar_init(&sm.landed[s], 1); f_mbar_init(&sm.summed[s], 1); f_mbar_init(&sm.mapped[s], 1);
            f_mbar_init(&sm.ready[s], 1); f_mbar_init(&sm.empty[s], NW);
            sm.lcnt[s] = 0;
        }
        sm.last_t = -1;
        sm.e_last = 1022;              // binade [0.5, 1): where a normalised running sum spends most of its life
        f_fence_mbar_init();
    }
    for (int q = tid; q < WIN; q += NT + 128) sm.win[q] = 0;
    __syncthreads();

    const double divisor = p.div ? *p.div : 1.0;

    if (wid >= NW && wid <= NW + 1) {
        // ============================================================ loader warps L0, L1
        // L0 claims tiles and starts their TMA loads.  Both loaders then sum one half of the tile's
        // rows each; whoever finishes second publishes the stage-1 AGGREGATE.  The loaders never wait
        // for another CTA, so every tile's sum is published as soon as its data has arrived — a
        // look-back only ever waits for loads, not for somebody else's look-back.  They also form the
        // row sums of the parity map in the binade the chain saw last (right for all but a handful of tiles).
        const int h = wid - NW;
        int q_issue = 0, q_proc = 0, exhausted = 0;
        long long pf[3] = {0, 0, 0}, tk = clock64();
        auto lap = [&](int i) { const long long now = clock64(); pf[i] += now - tk; tk = now; };
        // claim tiles (in order) and start their loads while stages are free; `block`: wait for the first
        auto issue = [&](bool block) {
            while (!exhausted && q_issue < q_proc + STAGES) {
                const int s = q_issue % STAGES, use = q_issue / STAGES;
                if (use > 0) {
                    if (block) f_mbar_wait(&sm.empty[s], (use - 1) & 1);
                    else {
                        int is_free = 0;
                        if (lane == 0) is_free = f_mbar_try(&sm.empty[s], (use - 1) & 1) ? 1 : 0;
                        if (!__shfl_sync(FULL, is_free, 0)) break;
                    }
                }
                int t = 0;
                if (lane == 0) t = atomicAdd(&p.hdr->tile_counter, 1);
                t = __shfl_sync(FULL, t, 0);
                if (t >= p.T) { exhausted = 1; t = -1; }
                if (lane == 0) {
                    sm.tile_of[s] = t;
                    if (t >= 0) f_trace(p, t, 0);
                    if (t >= 0 && p.use_tma) {
                        f_mbar_expect_tx(&sm.full_tma[s], TILE * 8);
                        f_tma_load_2d(sm.w[s], &wmap, 0, t * NT, &sm.full_tma[s]);
                    }
                }
                q_issue++;
                block = false;
            }
            __syncwarp();
        };
        for (;; q_proc++) {
            const int s = q_proc % STAGES, use = q_proc / STAGES;
            unsigned char *sb = reinterpret_cast<unsigned char *>(sm.w[s]);
            if (h == 0) {
                issue(q_proc == q_issue);
                lap(0);
                const int t0 = sm.tile_of[s];
                if (t0 >= 0) {
                    if (p.use_tma) {
                        f_mbar_wait(&sm.full_tma[s], use & 1);
                        // rows beyond n/16 arrive zero-filled; the last n % 16 weights are fetched by hand
                        const i64 R = p.n >> 4;
                        const int rem = (int)(p.n & 15);
                        if (rem && (R / NT) == t0 && lane < rem) {
                            const int r = (int)(R - (i64)t0 * NT);
                            *reinterpret_cast<double *>(sb + f_swz(r, lane >> 1) + (lane & 1) * 8) = p.w[R * 16 + lane];
                        }
                    } else {
                        const i64 j0 = (i64)t0 * TILE;
                        for (int i = lane; i < TILE; i += 32) {
                            const i64 j = j0 + i;
                            const double v = (j < p.n) ? p.w[j] : 0.0;
                            *reinterpret_cast<double *>(sb + f_swz(i >> 4, (i >> 1) & 7) + (i & 1) * 8) = v;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    sm.eg[s] = *reinterpret_cast<volatile int *>(&sm.e_last);
                    if (t0 >= 0) f_trace(p, t0, 1);
                    f_mbar_arrive(&sm.landed[s]);
                }
                __syncwarp();
                lap(1);
            } else {
                f_mbar_wait(&sm.landed[s], use & 1);
            }
            const int t = sm.tile_of[s];
            if (t < 0) {
                if (h == 0 && lane == 0) f_mbar_arrive(&sm.summed[s]);
                break;
            }
            if (h == 1 && p.use_tma) f_mbar_wait(&sm.full_tma[s], use & 1);     // already complete: acquires the TMA writes directly
            // tile sum (only the approximate prefix comes from it), validation, optional normalisation,
            // speculative row sums.  Lane L owns rows L, L + 32, ... of this loader's half: consecutive
            // lanes read consecutive swizzled rows (conflict-free).
            const int eg = sm.eg[s];
            const i64 gbase = (i64)eg << 52;
            const double G0 = __longlong_as_double(gbase), G1 = __longlong_as_double(gbase + 1);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            unsigned mx = 0, tie = 0;
#pragma unroll 2
            for (int i = h * (RPL / 2); i < (h + 1) * (RPL / 2); i++) {
                const int r = i * 32 + lane;
                i64 racc = 0;
#pragma unroll
                for (int c = 0; c < F_IPT / 2; c++) {
                    double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(r, c));
                    if (p.div) {
                        // fused normalisation: w / S, the same IEEE division NumPy's `w / w.sum()` performs
                        v.x = __ddiv_rn(v.x, divisor); v.y = __ddiv_rn(v.y, divisor);
                        *reinterpret_cast<double2 *>(sb + f_swz(r, c)) = v;
                    }
                    if (c & 1) { a2 += v.x; a3 += v.y; } else { a0 += v.x; a1 += v.y; }
                    mx = max(mx, max((unsigned)__double2hiint(v.x), (unsigned)__double2hiint(v.y)));
                    const i64 x0 = __double_as_longlong(__dadd_rn(G0, v.x)), x1 = __double_as_longlong(__dadd_rn(G1, v.x));
                    const i64 y0 = __double_as_longlong(__dadd_rn(G0, v.y)), y1 = __double_as_longlong(__dadd_rn(G1, v.y));
                    tie |= (((unsigned)x0 + 1u) ^ (unsigned)x1) | (((unsigned)y0 + 1u) ^ (unsigned)y1);   // d1 != d0: an exact tie
                    racc += (x0 - gbase) + (y0 - gbase);
                }
                sm.ex[s][r] = racc;
            }
            if (p.div) f_fence_proxy_async();              // generic-proxy writes to a stage the TMA engine will refill
            int bad = 0;
            if (mx >= 0x7FF00000u) {                       // negative, inf or nan somewhere (or a harmless -0.0)
                for (int i = h * (RPL / 2); i < (h + 1) * (RPL / 2); i++)
                    for (int c = 0; c < F_IPT / 2; c++) {
                        const double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(i * 32 + lane, c));
                        if (!(v.x >= 0.0) || !(v.y >= 0.0) || isinf(v.x) || isinf(v.y)) bad = 1;
                    }
            }
            bad = __any_sync(FULL, bad);
            const int any_tie = __any_sync(FULL, tie != 0);
            double tot = (a0 + a1) + (a2 + a3);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            __syncwarp();
            if (lane == 0) {
                sm.part[s][h].tot = tot; sm.part[s][h].bad = bad; sm.part[s][h].tie = any_tie;
                __threadfence_block();
                if (atomicAdd(&sm.lcnt[s], 1) == 1) {      // the second loader to finish publishes the tile's sum
                    __threadfence_block();
                    sm.lcnt[s] = 0;
                    double tt = sm.part[s][0].tot + sm.part[s][1].tot;
                    const int tb = sm.part[s][0].bad | sm.part[s][1].bad;
                    if (tb) { tt = 0.0; p.hdr->fallback = 1; }
                    f_st(p.st1 + t + 1, ((u64)__double_as_longlong(tt) & ~3ull) | ST1_AGG);
                    sm.pre[s].tot = tt; sm.pre[s].bad = tb; sm.pre[s].eg = eg; sm.pre[s].tie = sm.part[s][0].tie | sm.part[s][1].tie;
                    f_trace(p, t, 2);
                    f_mbar_arrive(&sm.summed[s]);
                }
            }
            if (h == 0) lap(2);
        }
        if (h == 0 && p.prof && lane == 0)
            for (int i = 0; i < 3; i++) atomicAdd(reinterpret_cast<unsigned long long *>(&p.hdr->prof[i]), (unsigned long long)pf[i]);
        return;
    }
    if (wid == NW + 2) {
        // ============================================================ chain warp C1 (stage 1)
        // approximate prefix -> which binade the tile lives in -> for a tile deep inside one binade the
        // parity map D, published as the stage-2 AGGREGATE right away (it never waits for stage 2)
        long long pf[3] = {0, 0, 0}, tk = clock64();
        auto lap = [&](int i) { const long long now = clock64(); pf[i] += now - tk; tk = now; };
        for (int q = 0;; q++) {
            const int s = q % STAGES, use = q / STAGES;
            f_mbar_wait(&sm.summed[s], use & 1);
            lap(0);
            const int t = sm.tile_of[s];
            if (t < 0) {
                if (lane == 0) f_mbar_arrive(&sm.mapped[s]);
                break;
            }
            unsigned char *sb = reinterpret_cast<unsigned char *>(sm.w[s]);
            if (lane == 0) f_trace(p, t, 3);
            const double tot = sm.pre[s].tot;
            const int bad = sm.pre[s].bad, eg = sm.pre[s].eg;
            int any_tie = sm.pre[s].tie;
            const double tp = f_lookback_sum(p, t, lane);
            if (lane == 0) { f_st(p.st1 + t + 1, ((u64)__double_as_longlong(tp + tot) & ~3ull) | ST1_INCL); f_trace(p, t, 4); }
            lap(1);
            int e0;
            const bool ca = clean_add(tp, tp + tot, p.eb, &e0);
            int kind = bad ? TK_BAD : ((ca || tot == 0.0) ? TK_CLEAN : TK_CROSS);
            i64 D = 0;
            if (kind == TK_CLEAN) {
                if (tot == 0.0) {
                    for (int i = 0; i < RPL; i++) sm.ex[s][i * 32 + lane] = 0;
                    any_tie = 0;
                } else if (e0 != eg) {
                    // the loaders' guess was wrong (first tiles, a new binade): row sums again, in binade e0
                    i64 rsum[RPL];
                    any_tie = f_row_sums<NW, STAGES>(sm, s, e0, 0, lane, rsum) ? 1 : 0;
#pragma unroll
                    for (int i = 0; i < RPL; i++) sm.ex[s][i * 32 + lane] = rsum[i];
                }
                if (tot != 0.0 && lane == 0) sm.e_last = e0;
                if (any_tie) kind = TK_TIES;               // ties: the general parity maps of the slow path
            }
            if (kind == TK_CLEAN) {
                // exclusive offsets per row (= consumer thread) and the tile's map D
                __syncwarp();
                i64 v[RPL], run = 0;
#pragma unroll
                for (int j = 0; j < RPL; j++) { const i64 x = sm.ex[s][lane * RPL + j]; v[j] = run; run += x; }
                const i64 inc = warp_incl_scan_i64(run, lane);
                const i64 lane_ex = inc - run;
#pragma unroll
                for (int j = 0; j < RPL; j++) sm.ex[s][lane * RPL + j] = lane_ex + v[j];
                D = __shfl_sync(FULL, inc, 31);
                if (lane == 0) { f_st(p.st2 + t + 1, st2_pack_agg(D, 0)); f_trace(p, t, 5); }
            }
            __syncwarp();
            if (lane == 0) {
                sm.mid[s].kind = kind; sm.mid[s].e0 = e0; sm.mid[s].tp = tp; sm.mid[s].tot = tot; sm.mid[s].D = D;
                f_mbar_arrive(&sm.mapped[s]);
            }
            lap(2);
        }
        if (p.prof && lane == 0)
            for (int i = 0; i < 3; i++) atomicAdd(reinterpret_cast<unsigned long long *>(&p.hdr->prof[3 + i]), (unsigned long long)pf[i]);
        return;
    }
    if (wid == NW + 3) {
        // ============================================================ chain warp C2 (stage 2)
        // exact state before the tile; tiles that may cross a binade are resolved here, exactly, from
        // that state (no margins): map arithmetic up to the row that leaves the binade, true adds
        // inside that row, map arithmetic of the next binade after it
        long long pf[3] = {0, 0, 0}, tk = clock64();
        auto lap = [&](int i) { const long long now = clock64(); pf[i] += now - tk; tk = now; };
        for (int q = 0;; q++) {
            const int s = q % STAGES, use = q / STAGES;
            f_mbar_wait(&sm.mapped[s], use & 1);
            lap(0);
            const int t = sm.tile_of[s];
            if (t < 0) {
                if (lane == 0) { sm.info[s].t = -1; f_mbar_arrive(&sm.ready[s]); }
                break;
            }
            const int kind = sm.mid[s].kind, e0 = sm.mid[s].e0;
            const double tot = sm.mid[s].tot;
            const i64 D = sm.mid[s].D;
            const i64 S_in = f_lookback_state(p, t, lane);
            lap(1);
            int mode, good = 1;
            i64 S_out = S_in, base = 0, lo = 0, cnt = 0;
            if (lane < NT / 32) sm.cross[s][lane] = 0;
            if (kind == TK_CLEAN) {
                S_out = S_in + D;
                good = (tot == 0.0) || ((int)(S_in >> 52) == e0 && (int)(S_out >> 52) == e0);
                base = S_in;
                mode = TM_FAST;
            } else if (kind == TK_BAD) {
                mode = TM_BAD;                             // invalid weights: the sequential kernel will produce the result
            } else if (kind == TK_CROSS && f_resolve_exact<NW, STAGES>(sm, s, S_in, &S_out, lane)) {
                mode = TM_FAST;                            // ex[] now holds the exact state before every row
                if (lane == 0) { atomicAdd(&p.hdr->n_unclean, 1); sm.e_last = (int)(S_out >> 52); }
            } else {
                mode = TM_SLOW;                            // ties / too many crossings: the consumers' general path publishes
            }
            if (lane == 0) {
                if (mode != TM_SLOW) { f_st(p.st2 + t + 1, ST2_INCL | (u64)S_out); f_trace(p, t, 6); }
                if (mode == TM_FAST) f_finish_tile<MODE>(p, t, S_in, S_out, good, lo, cnt);
                sm.info[s].t = t; sm.info[s].mode = mode; sm.info[s].good = good;
                sm.info[s].tp = __longlong_as_double(S_in); sm.info[s].S_in = S_in; sm.info[s].base = base;
                sm.info[s].lo = lo; sm.info[s].cnt = cnt;
                f_trace(p, t, 7);
            }
            __syncwarp();
            if (lane == 0) f_mbar_arrive(&sm.ready[s]);
            lap(2);
        }
        if (p.prof && lane == 0)
            for (int i = 0; i < 3; i++) atomicAdd(reinterpret_cast<unsigned long long *>(&p.hdr->prof[6 + i]), (unsigned long long)pf[i]);
        return;
    }

    // ================================================================ consumer warps
    // Fast tiles arrive with everything resolved (exact start state, output range): no global wait.
    const i64 out_begin = p.hdr->out_begin;
    const double Nd = (double)p.ng;
    long long cwait = 0, cwork = 0, ctk = clock64();
    for (int q = 0;; q++) {
        const int s = q % STAGES, use = q / STAGES;
        { const long long now = clock64(); cwork += now - ctk; ctk = now; }
        f_mbar_wait(&sm.ready[s], use & 1);
        { const long long now = clock64(); cwait += now - ctk; ctk = now; }
        const int t = sm.info[s].t;
        if (t < 0) {
            if (p.prof && tid == 0) {
                atomicAdd(reinterpret_cast<unsigned long long *>(&p.hdr->prof[9]), (unsigned long long)cwait);
                atomicAdd(reinterpret_cast<unsigned long long *>(&p.hdr->prof[10]), (unsigned long long)cwork);
            }
            break;
        }
        const int mode = sm.info[s].mode;
        if (mode == TM_BAD) {
            __syncwarp();
            if (lane == 0) f_mbar_arrive(&sm.empty[s]);
            continue;
        }
        if (tid == 0) { f_trace(p, t, 8); if (q > 0 && sm.last_t >= 0) f_trace(p, sm.last_t, 9); sm.last_t = t; }
        if (p.use_tma) f_mbar_wait(&sm.full_tma[s], use & 1);       // already complete: acquires the TMA writes directly
        unsigned char *sb = reinterpret_cast<unsigned char *>(sm.w[s]);
        const i64 jthread = (i64)t * TILE + (i64)tid * F_IPT;       // first particle of this thread (local numbering)
        i64 cb[F_IPT];
        i64 thread_start;                                   // exact state before this thread's first particle
        i64 tile_lo, tile_cnt;
        int good;
        if (mode == TM_FAST || p.wnorm_out) {
            double w[F_IPT];
#pragma unroll
            for (int c = 0; c < F_IPT / 2; c++) {
                const double2 v = *reinterpret_cast<const double2 *>(sb + f_swz(tid, c));
                w[2 * c] = v.x; w[2 * c + 1] = v.y;
            }
            if (p.wnorm_out) {                              // the producer left the normalised weights in the stage
                double *o = p.wnorm_out + jthread;
                if (jthread + F_IPT <= p.n && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                    for (int k = 0; k < F_IPT; k += 2) *reinterpret_cast<double2 *>(o + k) = make_double2(w[k], w[k + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < F_IPT; k++) if (jthread + k < p.n) o[k] = w[k];
                }
            }
            if (mode == TM_FAST) {
                good = sm.info[s].good; tile_lo = sm.info[s].lo; tile_cnt = sm.info[s].cnt;
                thread_start = sm.info[s].base + sm.ex[s][tid];
                i64 c = thread_start;
                if ((sm.cross[s][wid] >> lane) & 1) {
                    // this row's adds leave the binade of its start state: true adds
#pragma unroll
                    for (int k = 0; k < F_IPT; k++) {
                        c = __double_as_longlong(__dadd_rn(__longlong_as_double(c), w[k]));
                        cb[k] = c;
                    }
                } else {
                    // every add stays in the binade of the row's start state: bits(S) += d0
                    const i64 base = (thread_start >> 52) << 52;
                    const double B0 = __longlong_as_double(base);
#pragma unroll
                    for (int k = 0; k < F_IPT; k++) {
                        c += __double_as_longlong(__dadd_rn(B0, w[k])) - base;
                        cb[k] = c;
                    }
                }
            }
        }
        if (mode == TM_FAST) {
            // the stage is free once this warp holds everything it needs in registers
            __syncwarp();
            if (lane == 0) f_mbar_arrive(&sm.empty[s]);
        } else {
            good = f_slow_tile<NW, STAGES, MODE>(p, sm, s, t);
#pragma unroll
            for (int c = 0; c < F_IPT / 2; c++) {
                const longlong2 v = *reinterpret_cast<const longlong2 *>(sb + f_swz(tid, c));
                cb[2 * c] = v.x; cb[2 * c + 1] = v.y;
            }
            // state before the thread's first particle = the previous particle's c (the tile's S_in for thread 0)
            {
                const longlong2 v = *reinterpret_cast<const longlong2 *>(sb + f_swz(tid > 0 ? tid - 1 : 0, 7));
                thread_start = tid > 0 ? v.y : sm.bc_S_in;
            }
            tile_lo = sm.bc_lo; tile_cnt = sm.bc_cnt;
            __syncwarp();
            if (lane == 0) { f_fence_proxy_async(); f_mbar_arrive(&sm.empty[s]); }
        }
        if (!good) continue;
        if (MODE == F_CUMSUM) { f_store_cumsum(p, jthread, cb); continue; }

        // ---- output range end of every particle, relative to tile_lo: hv[k] = #{positions < c_k} - tile_lo
        int hv[F_IPT], hv_prev;
        if (MODE == F_SYS) {
            // branch-free: floor(c N - u) + 1 away from integers; the rare near-integer cases are redone exactly
            const double u = p.u, half_m = 0.5 - p.tau;
            const int n_m1 = (int)p.ng - 1, lo_m1 = (int)tile_lo - 1;
            unsigned slow = 0;
            auto count1 = [&](i64 cbits, unsigned bit) -> int {
                const double v = fma(__longlong_as_double(cbits), Nd, -u);    // >= -u > -1
                const double fl = floor(v);
                const double fr = v - fl;                                     // exact, in [0, 1)
                if (!(fabs(fr - 0.5) < half_m)) slow |= bit;                  // within tau of an integer
                return min(__double2int_rz(fl), n_m1) - lo_m1;                // floor(v) + 1 - tile_lo
            };
#pragma unroll
            for (int k = 0; k < F_IPT; k++) hv[k] = count1(cb[k], 1u << k);
            hv_prev = count1(thread_start, 1u << F_IPT);
            if (slow) {
#pragma unroll
                for (int k = 0; k < F_IPT; k++)
                    if (slow & (1u << k)) hv[k] = (int)(f_count_below<MODE>(p, __longlong_as_double(cb[k])) - tile_lo);
                if (slow & (1u << F_IPT)) hv_prev = (int)(f_count_below<MODE>(p, __longlong_as_double(thread_start)) - tile_lo);
            }
        } else {
#pragma unroll
            for (int k = 0; k < F_IPT; k++) hv[k] = (int)(f_count_below<MODE>(p, __longlong_as_double(cb[k])) - tile_lo);
            hv_prev = (int)(f_count_below<MODE>(p, __longlong_as_double(thread_start)) - tile_lo);
        }

        // ---- expansion
        const int base_j = (int)(p.j0 + (i64)t * TILE) - 1;                   // markers are local index + 1
        const i64 rel_lo = tile_lo - out_begin;
        const int mis = (int)(((reinterpret_cast<uintptr_t>(p.idx) >> 2) + (uintptr_t)rel_lo) & 3);
        if (tile_cnt + 3 <= WIN && rel_lo >= 0 && rel_lo + tile_cnt <= p.cap) {
            // one window; slot 0 is 16-byte aligned in the index array, the tile's first output is slot `mis`
            const int total = (int)tile_cnt + mis;
            int l = hv_prev + mis;
#pragma unroll
            for (int k = 0; k < F_IPT; k++) {
                const int h = hv[k] + mis;
                if (h > l) sm.win[l] = tid * F_IPT + k + 1;
                l = h;
            }
            f_bar<NT>();
            int m[F_SPT];
            f_window_scan<NW, STAGES>(sm, tid, lane, wid, m);
            const int s0 = tid * F_SPT;
            int *dst = p.idx + (rel_lo - mis) + s0;
            if (s0 >= mis && s0 + F_SPT <= total) {
#pragma unroll
                for (int i = 0; i < F_SPT; i += 4)
                    *reinterpret_cast<int4 *>(dst + i) = make_int4(base_j + m[i], base_j + m[i + 1], base_j + m[i + 2], base_j + m[i + 3]);
            } else if (s0 < total) {
#pragma unroll
                for (int i = 0; i < F_SPT; i++)
                    if (s0 + i >= mis && s0 + i < total) dst[i] = base_j + m[i];
            }
            continue;      // the next tile's first barrier separates these window reads from its marker writes
        }
        // general expansion: several windows, runs of BIGRUN or more copies go to the fill kernel
        if (tid == 0) atomicAdd(&p.hdr->n_general, 1);
        const int cnt = (int)tile_cnt;
        int cs = 0;
        while (cs < cnt) {
            if (tid == 0) sm.bc_skip = -1;
            f_bar<NT>();
            {
                int l = hv_prev;
#pragma unroll
                for (int k = 0; k < F_IPT; k++) {
                    const int h = hv[k];
                    if (l <= cs && cs < h && h - cs >= BIGRUN) {
                        sm.bc_skip = h;
                        const int r = atomicAdd(&p.hdr->n_runs, 1);
                        if (r < p.max_runs) p.runs[r] = Run{tile_lo + cs, tile_lo + h, base_j + tid * F_IPT + k + 1, 0};
                        else p.hdr->fallback = 1;
                    }
                    l = h;
                }
            }
            f_bar<NT>();
            const int skip = sm.bc_skip;
            if (skip >= 0) { cs = skip; f_bar<NT>(); continue; }
            const int ce = (cnt - cs > WIN) ? cs + WIN : cnt;
            {
                int l = hv_prev;
#pragma unroll
                for (int k = 0; k < F_IPT; k++) {
                    const int h = hv[k];
                    const int a0 = max(l, cs);
                    if (h > a0 && a0 < ce) sm.win[a0 - cs] = tid * F_IPT + k + 1;
                    l = h;
                }
            }
            f_bar<NT>();
            int m[F_SPT];
            f_window_scan<NW, STAGES>(sm, tid, lane, wid, m);
#pragma unroll
            for (int i = 0; i < F_SPT; i++) {
                const int sl = tid * F_SPT + i;
                if (sl < ce - cs) f_put_index(p, out_begin, tile_lo + cs + sl, base_j + m[i]);
            }
            f_bar<NT>();
            cs = ce;
        }
        f_bar<NT>();
    }
}

// ------------------------------------------------------------------ init / epilogue kernels
__global__ void __launch_bounds__(256) k_finit(FParams p)
{
    const i64 nst = (i64)p.T + 1;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x + 1; i < nst; i += (i64)gridDim.x * blockDim.x) { p.st1[i] = 0; p.st2[i] = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        FHeader h;
        memset(&h, 0, sizeof(h));
        const double ca = p.carry_approx ? *p.carry_approx : 0.0;
        const double ce = p.carry_exact ? *p.carry_exact : 0.0;
        const double Ngd = (double)p.ng;
        h.out_begin = p.cumsum_out ? 0 : (p.U ? count_below_str(ce, p.U, p.ng, Ngd) : count_below_sys(ce, p.u, p.ng, Ngd, p.tau));
        *p.hdr = h;
        p.st1[0] = ((u64)__double_as_longlong(ca) & ~3ull) | ST1_INCL;
        p.st2[0] = ST2_INCL | (u64)__double_as_longlong(ce);
    }
}

// long runs (one particle copied >= BIGRUN times), then — only if something failed — the literal
// sequential transcription of resampling.py:141-149, and the info block
__global__ void __launch_bounds__(256) k_fepilogue(FParams p)
{
    FHeader *hdr = p.hdr;
    if (!hdr->fallback && p.idx) {
        int nr = hdr->n_runs;
        if (nr > p.max_runs) nr = p.max_runs;
        const i64 ob = hdr->out_begin;
        for (int r = 0; r < nr; r++) {
            const Run run = p.runs[r];
            for (i64 i = run.lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < run.hi; i += (i64)gridDim.x * blockDim.x) {
                const i64 rel = i - ob;
                if (rel >= 0 && rel < p.cap) p.idx[rel] = run.j;
                else hdr->cap_overflow = 1;
            }
        }
    }
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    if (p.prof) {
        const double T = (double)p.T;
        printf("RSPROF tiles=%d cycles/tile: L0[stage-wait %.0f | tma-wait %.0f | sum-pass %.0f]  C1[wait-summed %.0f | lookback1 %.0f | maps %.0f]  C2[wait-mapped %.0f | lookback2 %.0f | finish %.0f]  cons[wait-ready %.0f | work %.0f]  crossing=%d slow=%d general=%d\n",
               p.T, hdr->prof[0] / T, hdr->prof[1] / T, hdr->prof[2] / T, hdr->prof[3] / T, hdr->prof[4] / T, hdr->prof[5] / T,
               hdr->prof[6] / T, hdr->prof[7] / T, hdr->prof[8] / T, hdr->prof[9] / T, hdr->prof[10] / T, hdr->n_unclean, hdr->n_slow, hdr->n_general);
    }
    auto write_info = [&](int overflow, int fb) {
        if (p.info) {
            p.info[0] = overflow; p.info[1] = fb; p.info[2] = hdr->n_unclean; p.info[3] = hdr->n_runs;
            p.info[4] = hdr->chain_bad | (hdr->timeout << 4); p.info[5] = hdr->n_seq; p.info[6] = hdr->cap_overflow; p.info[7] = hdr->n_slow;
        }
    };
    if (!hdr->fallback) { write_info(hdr->overflow, 0); return; }
    const double S = p.div ? *p.div : 1.0;
    auto W = [&](i64 q) { return p.div ? __ddiv_rn(p.w[q], S) : p.w[q]; };
    if (p.div && p.wnorm_out) for (i64 q = 0; q < p.n; q++) p.wnorm_out[q] = W(q);
    if (p.cumsum_out) {                          // cumsum mode: np.cumsum, one add at a time
        double c = 0.0;
        for (i64 q = 0; q < p.n; q++) { c = (q == 0) ? W(0) : __dadd_rn(c, W(q)); p.cumsum_out[q] = c; }
        if (p.cumsum_last) *p.cumsum_last = c;
        if (p.last_one) p.cumsum_out[p.n - 1] = 1.0;
        write_info(0, 1);
        return;
    }
    // resampling.py:141-149 — cumulative sum and two-pointer merge, one element at a time.  A shard
    // starts from the exact running sum of the earlier shards and owns the positions from
    // count_below(carry) up to count_below(its last cumulative sum).
    const double Ngd = (double)p.ng;
    const double carry = p.carry_exact ? *p.carry_exact : 0.0;
    auto pos = [&](i64 i) { return p.U ? pos_str(i, p.U, Ngd) : pos_sys(i, p.u, Ngd); };
    i64 lo = 0, hi = p.ng;                       // first i with pos_i >= carry (positions are non-decreasing)
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (pos(mid) < carry) lo = mid + 1; else hi = mid;
    }
    const i64 ob = lo;
    hdr->out_begin = ob;
    hdr->cap_overflow = 0;
    i64 i = ob, j = 0;
    double c = (carry == 0.0) ? W(0) : __dadd_rn(carry, W(0));
    int overflow = 0;
    while (i < p.ng) {
        if (pos(i) < c) {
            if (i - ob < p.cap) p.idx[i - ob] = (int)(p.j0 + j); else hdr->cap_overflow = 1;
            i++;
        } else {
            j++;
            if (j >= p.n) {
                if (p.is_last) {
                    overflow = (int)(p.ng - i);
                    for (; i < p.ng; i++) { if (i - ob < p.cap) p.idx[i - ob] = (int)(p.ng - 1); else hdr->cap_overflow = 1; }
                }
                break;
            }
            c = __dadd_rn(c, W(j));
        }
    }
    for (i64 q = j + 1; q < p.n; q++) c = __dadd_rn(c, W(q));
    if (p.cumsum_last) *p.cumsum_last = c;
    hdr->out_end = i;
    if (p.out_range) { p.out_range[0] = ob; p.out_range[1] = i; }
    write_info(overflow, 1);
}

// ------------------------------------------------------------------ host side
namespace {

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn f_get_encode()
{
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeFn)ptr;
    }
    return fn;
}

// the weights as rows of 16 doubles (128 bytes); a tile is a box of `box_rows` rows
bool f_make_map(CUtensorMap *m, const double *base, int64_t rows, int box_rows)
{
    EncodeFn enc = f_get_encode();
    if (!enc || rows < 1) return false;
    cuuint64_t gdim[2] = {16, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int f_env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int NW, int STAGES, int MODE>
int f_launch(const CUtensorMap &map, const FParams &p, cudaStream_t s)
{
    auto kern = k_fused<NW, STAGES, MODE>;
    const int smem = (int)sizeof(FSmem<NW, STAGES>);
    static bool configured[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "cudaFuncSetAttribute")) return BKE_ERR_CUDA;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int ctas_env = f_env_int("BKE_RS_CTAS", 0);
    const int per_sm = ctas_env > 0 ? ctas_env : f_ctas(NW, STAGES);
    int grid = sm_count() * per_sm;
    if (grid > p.T) grid = p.T;
    kern<<<grid, NW * 32 + 128, smem, s>>>(map, p);
    return check_cuda(cudaGetLastError(), "k_fused launch");
}

template <int NW, int STAGES>
int f_launch_mode(int mode, const CUtensorMap &map, const FParams &p, cudaStream_t s)
{
    if (mode == F_CUMSUM) return f_launch<NW, STAGES, F_CUMSUM>(map, p, s);
    if (mode == F_STRAT) return f_launch<NW, STAGES, F_STRAT>(map, p, s);
    return f_launch<NW, STAGES, F_SYS>(map, p, s);
}

}  // namespace

static unsigned long long *g_trace = nullptr;
void f_set_trace(void *buf) { g_trace = (unsigned long long *)buf; }

size_t f_carve(int64_t n, unsigned char *base, FParams *p)
{
    // the smallest tile any variant uses decides the number of status words
    const int64_t Tmax = (n + 4 * 32 * F_IPT - 1) / (4 * 32 * F_IPT);
    size_t off = 0;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return base ? base + o : nullptr; };
    unsigned char *q;
    q = take(sizeof(FHeader));               if (p) p->hdr = (FHeader *)q;
    q = take(sizeof(u64) * (Tmax + 1));      if (p) p->st1 = (u64 *)q;
    q = take(sizeof(u64) * (Tmax + 1));      if (p) p->st2 = (u64 *)q;
    const int64_t max_runs = n / BIGRUN + 8;
    q = take(sizeof(Run) * max_runs);        if (p) { p->runs = (Run *)q; p->max_runs = (int)max_runs; }
    return off;
}

int f_run(const FRunArgs &a, cudaStream_t s)
{
    const i64 n = a.n;
    if (n < 0 || a.ng < n || a.j0 < 0) { set_error("bad particle counts"); return BKE_ERR_BAD_ARG; }
    if (n == 0) return BKE_OK;
    if (a.ng >= ((i64)1 << 31)) { set_error("n must be < 2^31 (indexes are int32, resampling.py:141)"); return BKE_ERR_BAD_ARG; }
    if (!a.w || !(a.idx || a.cumsum_out) || !a.workspace) { set_error("weights, indexes and workspace must be non-NULL"); return BKE_ERR_BAD_ARG; }
    if (!a.U && !a.cumsum_out && !(a.u >= 0.0 && a.u < 1.0)) { set_error("u must be in [0, 1)"); return BKE_ERR_BAD_ARG; }
    const size_t need = f_carve(n, nullptr, nullptr);
    if (a.ws_bytes < need) { set_error("workspace too small: %zu < %zu", a.ws_bytes, need); return BKE_ERR_BAD_ARG; }
    if (reinterpret_cast<uintptr_t>(a.workspace) & 255) { set_error("workspace must be 256-byte aligned"); return BKE_ERR_BAD_ARG; }
    FParams p;
    memset(&p, 0, sizeof(p));
    f_carve(n, (unsigned char *)a.workspace, &p);
    const int nw_env = f_env_int("BKE_RS_WARPS", 8);
    const int NW = (nw_env == 4) ? 4 : 8;
    p.sleep_ns = f_env_int("BKE_RS_SLEEP", 0);
    p.prof = f_env_int("BKE_RS_PROF", 0);
    p.trace = g_trace;
    p.lbk = f_env_int("BKE_RS_LBK", 4);
    if (p.lbk < 1) p.lbk = 1;
    if (p.lbk > LBK_MAX) p.lbk = LBK_MAX;
    const int tile = NW * 32 * F_IPT;
    p.w = a.w; p.n = n; p.ng = a.ng; p.j0 = a.j0; p.cap = a.cap; p.is_last = a.is_last;
    p.carry_approx = a.carry_approx; p.carry_exact = a.carry_exact; p.out_range = a.out_range;
    p.u = a.u; p.U = a.U; p.idx = a.idx; p.info = a.info; p.cumsum_last = a.cumsum_last;
    p.cumsum_out = a.cumsum_out; p.last_one = a.last_one; p.div = a.div; p.wnorm_out = a.wnorm_out;
    p.T = (int)((n + tile - 1) / tile);
    // |exact sequential sum - approximate sum| in ulps of the running sum: N adds of the reference,
    // the tree sums inside a tile, the T sequential adds and the 2 flag bits per published word of
    // the look-back (and the division of a normalised call); doubled, plus slack.
    const i64 Tg = a.ng / tile + 2;
    p.eb = 2 * (a.ng + 16 * Tg + 2 * tile) + (a.ng >> 4);
    const double tau = ldexp((double)a.ng, -46);
    p.tau = tau > 1e-6 ? tau : 1e-6;
    // TMA path: 16-byte aligned base and at least one full row of 16 weights
    static thread_local CUtensorMap map;
    static thread_local const void *map_ptr = nullptr;
    static thread_local i64 map_n = -1;
    static thread_local int map_nw = 0;
    const int tma_env = f_env_int("BKE_RS_TMA", 1);
    p.use_tma = tma_env && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0 && (n >> 4) >= 1 && f_get_encode() != nullptr;
    if (p.use_tma && !(map_ptr == a.w && map_n == n && map_nw == NW)) {
        if (!f_make_map(&map, a.w, n >> 4, NW * 32)) p.use_tma = 0;
        else { map_ptr = a.w; map_n = n; map_nw = NW; }
    }
    if (!p.use_tma) memset(&map, 0, sizeof(map)), map_ptr = nullptr;
    const int init_blocks = (int)((p.T + 1 + 255) / 256) < 64 ? (int)((p.T + 1 + 255) / 256) : 64;
    k_finit<<<init_blocks, 256, 0, s>>>(p);
    int rc;
    const int mode = a.cumsum_out ? F_CUMSUM : (a.U ? F_STRAT : F_SYS);
    const int st = f_env_int("BKE_RS_STAGES", F_STAGES);
    if (NW == 4) rc = st <= 2 ? f_launch_mode<4, 2>(mode, map, p, s) : (st == 3 ? f_launch_mode<4, 3>(mode, map, p, s) : f_launch_mode<4, 4>(mode, map, p, s));
    else rc = st <= 2 ? f_launch_mode<8, 2>(mode, map, p, s) : f_launch_mode<8, 3>(mode, map, p, s);
    if (rc != BKE_OK) return rc;
    k_fepilogue<<<sm_count() * 4, 256, 0, s>>>(p);
    return check_cuda(cudaGetLastError(), "resample launch");
}

}  // namespace rs
}  // namespace bke
