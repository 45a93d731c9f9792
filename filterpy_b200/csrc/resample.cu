// resample.cu — systematic / stratified particle resampling, bit-exact against the reference's
// strictly sequential fp64 cumsum (filterpy/monte_carlo/resampling.py:117-150, :80-114).
//
//   indexes[i] = #{ j : c_j <= pos_i },   c_j = fl(c_{j-1} + w_j)  (np.cumsum, :142),
//   pos_i = fl(fl(u + i) / N)  (systematic, :139)   or   fl(fl(U_i + i) / N)  (stratified, :103)
//
// A parallel fp64 scan rounds in a different order than np.cumsum and flips output indices, so the
// running sum is reproduced EXACTLY instead.  While the running sum S stays inside one binade
// (exponent field e, ulp q = 2^(max(e,1)-1075)) it is an integer multiple of q, and adding a weight
// w is the integer map  S/q -> S/q + d[parity(S/q)]  with d0 = d1 = rne(w/q) except for an exact
// tie (fraction 1/2), which rounds to the even neighbour and therefore depends on the parity.
// Such parity maps compose associatively ((f;g)[p] = f[p] + g[(p + f[p]) & 1]), so a parallel scan
// over them reproduces the sequential rounding.  The few elements whose addition may leave the
// binade ("raw" elements, found with an approximate scan and a rigorous error margin) are applied
// by a true fp64 add in a tiny sequential chain.  Every assumption (start and end of a mapped
// segment in the assumed binade) is verified with the exact values; if one fails, or the weights
// contain negative / non-finite entries, a literal single-thread transcription of the reference
// loop produces the result instead (info[1] = 1), so the output is always the reference's.
//
// Passes (all on one stream, no host sync):
//   A  tile sums (approximate, fp64 tree)            reads w
//   B  exclusive scan of tile sums                   1 CTA
//   C  per-tile parity maps / raw-element lists      reads w
//   D  exact chain over tiles                        1 CTA
//   E  exact c_j, output ranges, index expansion     reads w, writes indexes
//   F  long runs (one particle copied >= 8192 times) writes indexes
//   G  sequential fallback (normally exits at once)
#include "bke_internal.cuh"

namespace bke {
namespace rs {

constexpr int BLOCK = 256;
constexpr int IPT = 8;
constexpr int TILE = BLOCK * IPT;        // 2048 particles per tile
constexpr int RMAX = 64;                 // raw elements per tile before giving up
constexpr int UMAX = 2048;               // tiles with raw elements before giving up
constexpr int EXPAND = 4096;             // outputs expanded per shared-memory pass
constexpr int BIGRUN = 2 * EXPAND;       // runs this long go to the fill kernel
constexpr int CHAIN_THREADS = 1024;
constexpr int CHAIN_BATCH = 8;          // unclean tiles staged in shared memory per round of the chain

typedef long long i64;
typedef unsigned long long u64;

struct Map { i64 d0, d1; };
// parity map since the last raw element + raw count + the binade the map was built for
// (k: -2 = identity / nothing yet, -3 = poisoned: elements of different binades were mixed)
struct SMap { i64 d0, d1; int cnt; int k; };

__device__ __forceinline__ Map compose(Map f, Map g)
{
    Map h;
    h.d0 = f.d0 + ((f.d0 & 1) ? g.d1 : g.d0);
    h.d1 = f.d1 + (((f.d1 + 1) & 1) ? g.d1 : g.d0);
    return h;
}
__device__ __forceinline__ int merge_k(int a, int b)
{
    if (a == -2) return b;
    if (b == -2) return a;
    return a == b ? a : -3;
}
__device__ __forceinline__ SMap combine(SMap a, SMap b)
{
    if (b.cnt > 0) { b.cnt += a.cnt; return b; }
    Map h = compose(Map{a.d0, a.d1}, Map{b.d0, b.d1});
    return SMap{h.d0, h.d1, a.cnt, merge_k(a.k, b.k)};
}

__device__ __forceinline__ int efield(double s) { return (int)((u64)__double_as_longlong(s) >> 52) & 0x7ff; }
__device__ __forceinline__ i64 sint(double s)
{
    u64 b = (u64)__double_as_longlong(s);
    u64 m = b & 0xFFFFFFFFFFFFFull;
    return (i64)(((b >> 52) & 0x7ff) ? (m | (1ull << 52)) : m);
}
__device__ __forceinline__ i64 limit_of(int e) { return e ? (1ll << 53) : (1ll << 52); }
__device__ __forceinline__ double rebuild(int e, i64 si)
{
    u64 b = ((u64)e << 52) | ((u64)si & 0xFFFFFFFFFFFFFull);
    return __longlong_as_double((i64)b);
}

// state S (exact, in binade e) advanced by a parity map; *bad is set when the assumption fails
__device__ __forceinline__ double apply_map(double S, Map m, int e, int *bad)
{
    if (efield(S) != e) { *bad = 1; return S; }
    i64 si = sint(S);
    i64 d = (si & 1) ? m.d1 : m.d0;
    si += d;
    if (si >= limit_of(e) || d < 0) { *bad = 1; return S; }
    return rebuild(e, si);
}

// parity map of adding w to a state in binade e (only called for elements classified clean)
__device__ __forceinline__ Map elem_map(double w, int e)
{
    u64 b = (u64)__double_as_longlong(w);
    int ew = (int)(b >> 52) & 0x7ff;
    u64 mw = (b & 0xFFFFFFFFFFFFFull) | (ew ? (1ull << 52) : 0ull);
    int sh = (ew > 1 ? ew : 1) - (e > 1 ? e : 1);       // exponent of w's ulp minus exponent of q
    Map m;
    if (sh >= 0) {
        i64 a = (sh < 10) ? (i64)(mw << sh) : (i64)(1ll << 62);   // clean elements have sh <= 1
        m.d0 = m.d1 = a;
    } else {
        int r = -sh;
        if (r >= 64) { m.d0 = m.d1 = 0; return m; }
        u64 a = mw >> r;
        u64 rem = mw & ((1ull << r) - 1ull);
        u64 half = 1ull << (r - 1);
        if (rem > half) { m.d0 = m.d1 = (i64)(a + 1); }
        else if (rem < half) { m.d0 = m.d1 = (i64)a; }
        else { m.d0 = (i64)(a + (a & 1)); m.d1 = (i64)(a + 1 - (a & 1)); }
    }
    return m;
}

// ------------------------------------------------------------------ block primitives (256 threads)
__device__ __forceinline__ double block_excl_scan_double(double v, double *total, double *sh /*[8]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    double base = 0.0, tot = 0.0;
#pragma unroll
    for (int k = 0; k < BLOCK / 32; k++) {
        double s = sh[k];
        if (k < wid) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + (inc - v);
}

__device__ __forceinline__ SMap shfl_up_smap(SMap v, int o)
{
    SMap r;
    r.d0 = __shfl_up_sync(FULL, v.d0, o);
    r.d1 = __shfl_up_sync(FULL, v.d1, o);
    r.cnt = __shfl_up_sync(FULL, v.cnt, o);
    r.k = __shfl_up_sync(FULL, v.k, o);
    return r;
}

// exclusive scan of SMap over the block; *total = aggregate of the whole block
__device__ __forceinline__ SMap block_excl_scan_smap(SMap v, SMap *total, SMap *sh /*[8]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    SMap inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        SMap t = shfl_up_smap(inc, o);
        if (lane >= o) inc = combine(t, inc);
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    SMap base = SMap{0, 0, 0, -2}, tot = SMap{0, 0, 0, -2};
#pragma unroll
    for (int k = 0; k < BLOCK / 32; k++) {
        SMap s = sh[k];
        if (k < wid) base = combine(base, s);
        tot = combine(tot, s);
    }
    __syncthreads();
    SMap prev = shfl_up_smap(inc, 1);                  // inclusive of the previous lane
    if (lane == 0) prev = SMap{0, 0, 0, -2};
    *total = tot;
    return combine(base, prev);
}

// ------------------------------------------------------------------ workspace
struct Header {
    int fallback;       // 1 -> the sequential kernel must produce the result
    int n_unclean;      // tiles with raw elements
    int n_runs;         // long runs queued for the fill kernel
    int overflow;       // positions >= cumsum[-1]
    int chain_bad;      // a verified assumption failed
    int pad[3];
};

struct Slot {           // one tile with raw elements
    int tile;
    int nraw;
    int segk[RMAX + 1];         // binade of segment s, -1 = empty segment
    i64 end0[RMAX + 1], end1[RMAX + 1];
    double wraw[RMAX];
};

struct Run { i64 lo, hi; int j; int pad; };

struct Ws {
    Header *hdr;
    double *tile_sum;       // [T]
    double *tile_prefix;    // [T+1] approximate exclusive prefix (+ carry)
    double *S_in;           // [T+1] exact state before each tile
    i64 *tile_map;          // [T][2]
    int *tile_k;            // [T]  binade, or -1 for a tile with raw elements
    int *tile_slot;         // [T]
    i64 *run_map;           // [T][2] composite of the clean tiles since the last unclean one
    int *run_id;            // [T]
    int *run_k;             // [T]
    Slot *slots;            // [UMAX]
    double *S_run;          // [UMAX+1]
    int *ord2tile;          // [UMAX]
    Run *runs;              // [max_runs]
    int max_runs;
    int T;
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t carve(int64_t n, unsigned char *base, Ws *w)
{
    const int64_t T = (n + TILE - 1) / TILE;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return base ? base + o : nullptr; };
    unsigned char *p;
    p = take(sizeof(Header));                 if (w) w->hdr = (Header *)p;
    p = take(sizeof(double) * T);             if (w) w->tile_sum = (double *)p;
    p = take(sizeof(double) * (T + 1));       if (w) w->tile_prefix = (double *)p;
    p = take(sizeof(double) * (T + 1));       if (w) w->S_in = (double *)p;
    p = take(sizeof(i64) * 2 * T);            if (w) w->tile_map = (i64 *)p;
    p = take(sizeof(int) * T);                if (w) w->tile_k = (int *)p;
    p = take(sizeof(int) * T);                if (w) w->tile_slot = (int *)p;
    p = take(sizeof(i64) * 2 * T);            if (w) w->run_map = (i64 *)p;
    p = take(sizeof(int) * T);                if (w) w->run_id = (int *)p;
    p = take(sizeof(int) * T);                if (w) w->run_k = (int *)p;
    p = take(sizeof(Slot) * UMAX);            if (w) w->slots = (Slot *)p;
    p = take(sizeof(double) * (UMAX + 1));    if (w) w->S_run = (double *)p;
    p = take(sizeof(int) * UMAX);             if (w) w->ord2tile = (int *)p;
    int64_t max_runs = n / BIGRUN + 8;
    p = take(sizeof(Run) * max_runs);         if (w) { w->runs = (Run *)p; w->max_runs = (int)max_runs; w->T = (int)T; }
    return off;
}

struct Params {
    const double *w;
    i64 n;                 // particles in this call
    double u;              // systematic offset
    const double *U;       // stratified uniforms (NULL = systematic)
    int *idx;
    double eps;            // relative margin of the approximate prefix
    double tau;            // fast-path margin of the position search
    int *info;             // user info[8] or NULL
    double *cumsum_last;   // or NULL
    Ws ws;
};

// ------------------------------------------------------------------ pass A: tile sums
__global__ void __launch_bounds__(BLOCK) k_tile_sums(Params p)
{
    __shared__ double sh[BLOCK / 32];
    const int t = blockIdx.x;
    const i64 base = (i64)t * TILE + (i64)threadIdx.x * IPT;
    double s = 0.0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        i64 j = base + k;
        double w = (j < p.n) ? p.w[j] : 0.0;
        if (!(w >= 0.0) || isinf(w)) bad = true;
        s += w;
    }
    double tot;
    block_excl_scan_double(s, &tot, sh);
    if (threadIdx.x == 0) p.ws.tile_sum[t] = tot;
    if (bad) p.ws.hdr->fallback = 1;
}

// ------------------------------------------------------------------ pass B: scan of tile sums
__global__ void __launch_bounds__(CHAIN_THREADS) k_scan_tiles(Params p)
{
    __shared__ double sh[CHAIN_THREADS];
    const int T = p.ws.T;
    const int per = (T + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const int a = threadIdx.x * per;
    const int b = min(T, a + per);
    double s = 0.0;
    for (int t = a; t < b; t++) s += p.ws.tile_sum[t];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (int k = 0; k < CHAIN_THREADS; k++) { double v = sh[k]; sh[k] = run; run += v; }
        p.ws.tile_prefix[T] = run;
    }
    __syncthreads();
    double run = sh[threadIdx.x];
    for (int t = a; t < b; t++) { p.ws.tile_prefix[t] = run; run += p.ws.tile_sum[t]; }
}

// ------------------------------------------------------------------ shared tile analysis (C and E)
struct TileAn {
    double w[IPT];
    SMap inc[IPT];         // inclusive segmented parity map at each element
    int ek[IPT];           // binade assumed for the element (clean), -2 (zero weight) or -1 (raw)
    int nraw;              // raw elements in the tile
    SMap total;            // tile aggregate
};

struct TileShared {
    double shd[BLOCK / 32];
    SMap shm[BLOCK / 32];
    int first_raw[BLOCK + 1];    // raw flag of each thread's first element (+ sentinel)
};

__device__ __forceinline__ void analyse_tile(const Params &p, int t, TileAn &an, TileShared &sm)
{
    const i64 base = (i64)t * TILE + (i64)threadIdx.x * IPT;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        i64 j = base + k;
        an.w[k] = (j < p.n) ? p.w[j] : 0.0;
        s += an.w[k];
    }
    double tot;
    double before = p.ws.tile_prefix[t] + block_excl_scan_double(s, &tot, sm.shd);
    // classify + element maps, thread-local inclusive segmented scan
    SMap run = SMap{0, 0, 0, -2};
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        double after = before + an.w[k];
        double lo = before * (1.0 - p.eps);
        double hi = after * (1.0 + p.eps);
        int e = efield(lo);
        bool clean = (efield(hi) == e) && (hi < 1.0e308);
        SMap el;
        if (an.w[k] == 0.0) {            // fl(S + 0) = S in every binade: a binade-neutral identity
            el = SMap{0, 0, 0, -2};
            an.ek[k] = -2;
        } else if (clean) {
            Map m = elem_map(an.w[k], e);
            el = SMap{m.d0, m.d1, 0, e};
            an.ek[k] = e;
        } else {
            el = SMap{0, 0, 1, -2};
            an.ek[k] = -1;
        }
        run = combine(run, el);
        an.inc[k] = run;
        before = after;
    }
    SMap total;
    SMap excl = block_excl_scan_smap(run, &total, sm.shm);
#pragma unroll
    for (int k = 0; k < IPT; k++) an.inc[k] = combine(excl, an.inc[k]);
    an.nraw = total.cnt;
    an.total = total;
}

// ------------------------------------------------------------------ pass C: per-tile maps
__global__ void __launch_bounds__(BLOCK) k_tile_maps(Params p)
{
    __shared__ TileShared sm;
    __shared__ int s_slot;
    const int t = blockIdx.x;
    if (p.ws.hdr->fallback) return;
    TileAn an;
    analyse_tile(p, t, an, sm);
    if (an.nraw == 0) {
        if (threadIdx.x == BLOCK - 1) {
            if (an.total.k == -3) p.ws.hdr->fallback = 1;    // mixed binades inside one map (never expected)
            p.ws.tile_k[t] = an.total.k;                     // binade, or -2 for an all-zero tile
            p.ws.tile_map[2 * t] = an.total.d0;
            p.ws.tile_map[2 * t + 1] = an.total.d1;
            p.ws.tile_slot[t] = -1;
        }
        return;
    }
    if (threadIdx.x == 0) {
        int s = -1;
        if (an.nraw <= RMAX) {
            s = atomicAdd(&p.ws.hdr->n_unclean, 1);
            if (s >= UMAX) s = -1;
        }
        if (s < 0) p.ws.hdr->fallback = 1;
        s_slot = s;
        p.ws.tile_k[t] = -1;
        p.ws.tile_slot[t] = s < 0 ? 0 : s;
    }
    // raw flag of the element that follows each thread's last element
    sm.first_raw[threadIdx.x] = (an.ek[0] == -1);
    if (threadIdx.x == 0) sm.first_raw[BLOCK] = 1;       // the tile end closes the last segment
    __syncthreads();
    const int s = s_slot;
    if (s < 0) return;
    Slot *sl = &p.ws.slots[s];
    for (int q = threadIdx.x; q <= RMAX; q += BLOCK) { sl->segk[q] = -1; sl->end0[q] = 0; sl->end1[q] = 0; }
    if (threadIdx.x == 0) { sl->tile = t; sl->nraw = an.nraw; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int seg = an.inc[k].cnt;
        if (an.ek[k] == -1) {
            sl->wraw[seg - 1] = an.w[k];                 // the raw element that opens segment `seg`
        } else {
            bool next_raw = (k + 1 < IPT) ? (an.ek[k + 1] == -1) : (sm.first_raw[threadIdx.x + 1] != 0);
            if (next_raw) {
                if (an.inc[k].k == -3) p.ws.hdr->fallback = 1;
                // k == -2: only zero weights in the segment -> identity, recorded as empty
                sl->segk[seg] = an.inc[k].k == -2 ? -1 : an.inc[k].k;
                sl->end0[seg] = an.inc[k].d0; sl->end1[seg] = an.inc[k].d1;
            }
        }
    }
}

// sequential walk through one tile with raw elements; returns the exact state after the tile.
// segstate (optional, shared memory) receives the state at the start of every segment.
__device__ double walk_slot(const int *segk, const i64 *end0, const i64 *end1, const double *wraw, int nraw,
                            double S, int *bad, double *segstate)
{
    for (int s = 0; s <= nraw; s++) {
        if (segstate) segstate[s] = S;
        if (segk[s] >= 0) S = apply_map(S, Map{end0[s], end1[s]}, segk[s], bad);
        if (s < nraw) S = __dadd_rn(S, wraw[s]);
    }
    return S;
}

// ------------------------------------------------------------------ pass D: exact chain over tiles
struct RunEl { i64 d0, d1; int cnt; int k; };       // k: binade of the composite, -2 = identity
__device__ __forceinline__ RunEl run_combine(RunEl a, RunEl b, int *bad)
{
    if (b.cnt > 0) { b.cnt += a.cnt; return b; }
    if (a.k == -2) { b.cnt = a.cnt; return b; }
    if (b.k == -2) return a;
    if (a.k != b.k) *bad = 1;                        // clean tiles of one run share the binade
    Map h = compose(Map{a.d0, a.d1}, Map{b.d0, b.d1});
    return RunEl{h.d0, h.d1, a.cnt, a.k};
}

__global__ void __launch_bounds__(CHAIN_THREADS) k_chain(Params p, double carry)
{
    __shared__ RunEl sh[CHAIN_THREADS];
    __shared__ int s_bad;
    const Ws &ws = p.ws;
    if (ws.hdr->fallback) return;
    const int T = ws.T;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    int bad = 0;
    const int per = (T + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const int a = threadIdx.x * per, b = min(T, a + per);
    // element of tile t: clean -> its map (cnt 0); unclean -> reset marker (cnt 1, identity)
    auto tile_el = [&](int t) {
        if (ws.tile_slot[t] >= 0) return RunEl{0, 0, 1, -2};
        return RunEl{ws.tile_map[2 * t], ws.tile_map[2 * t + 1], 0, ws.tile_k[t]};
    };
    RunEl agg = RunEl{0, 0, 0, -2};
    for (int t = a; t < b; t++) agg = run_combine(agg, tile_el(t), &bad);
    sh[threadIdx.x] = agg;
    __syncthreads();
    if (threadIdx.x == 0) {
        RunEl run = RunEl{0, 0, 0, -2};
        for (int k = 0; k < CHAIN_THREADS; k++) { RunEl v = sh[k]; sh[k] = run; run = run_combine(run, v, &bad); }
    }
    __syncthreads();
    RunEl run = sh[threadIdx.x];
    for (int t = a; t < b; t++) {
        // exclusive value at tile t: composite of the clean tiles since the last unclean tile
        ws.run_map[2 * t] = run.d0; ws.run_map[2 * t + 1] = run.d1;
        ws.run_id[t] = run.cnt; ws.run_k[t] = run.k;
        if (ws.tile_slot[t] >= 0) ws.ord2tile[run.cnt] = t;
        run = run_combine(run, tile_el(t), &bad);
    }
    __threadfence_block();
    __syncthreads();
    // sequential part: one thread walks the tiles that contain raw elements.  Their slot data is
    // staged into shared memory by the whole block first (a dependent chain of global loads would
    // cost ~1 us per hop), CHAIN_BATCH tiles at a time.
    {
        __shared__ Slot s_slots[CHAIN_BATCH];
        __shared__ i64 s_rm[CHAIN_BATCH][2];
        __shared__ int s_rk[CHAIN_BATCH];
        __shared__ double s_S;
        const int U = ws.hdr->n_unclean;
        if (threadIdx.x == 0) { s_S = carry; ws.S_run[0] = carry; }
        for (int i0 = 0; i0 < U; i0 += CHAIN_BATCH) {
            const int nb = min(CHAIN_BATCH, U - i0);
            __syncthreads();
            for (int q = threadIdx.x; q < nb * (int)(sizeof(Slot) / sizeof(int)); q += CHAIN_THREADS) {
                const int b = q / (int)(sizeof(Slot) / sizeof(int)), o = q % (int)(sizeof(Slot) / sizeof(int));
                const int t = ws.ord2tile[i0 + b];
                reinterpret_cast<int *>(&s_slots[b])[o] = reinterpret_cast<const int *>(&ws.slots[ws.tile_slot[t]])[o];
            }
            if (threadIdx.x < nb) {
                const int t = ws.ord2tile[i0 + threadIdx.x];
                s_rm[threadIdx.x][0] = ws.run_map[2 * t]; s_rm[threadIdx.x][1] = ws.run_map[2 * t + 1];
                s_rk[threadIdx.x] = ws.run_k[t];
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double S = s_S;
                for (int b = 0; b < nb; b++) {
                    if (s_rk[b] != -2) S = apply_map(S, Map{s_rm[b][0], s_rm[b][1]}, s_rk[b], &bad);
                    const Slot *sl = &s_slots[b];
                    S = walk_slot(sl->segk, sl->end0, sl->end1, sl->wraw, sl->nraw, S, &bad, nullptr);
                    ws.S_run[i0 + b + 1] = S;
                }
                s_S = S;
            }
        }
    }
    __threadfence_block();
    __syncthreads();
    // parallel part: exact state before every tile, with verification of the clean tiles
    double S_last = 0.0;
    for (int t = a; t < b; t++) {
        double S = ws.S_run[ws.run_id[t]];
        if (ws.run_k[t] != -2) S = apply_map(S, Map{ws.run_map[2 * t], ws.run_map[2 * t + 1]}, ws.run_k[t], &bad);
        ws.S_in[t] = S;
        if (ws.tile_slot[t] < 0) {
            double E = ws.tile_k[t] >= 0 ? apply_map(S, Map{ws.tile_map[2 * t], ws.tile_map[2 * t + 1]}, ws.tile_k[t], &bad) : S;
            if (t == T - 1) S_last = E;
        } else if (t == T - 1) {
            S_last = ws.S_run[ws.run_id[t] + 1];
        }
        if (t == T - 1) ws.S_in[T] = S_last;
    }
    if (bad) s_bad = 1;
    __syncthreads();
    if (threadIdx.x == 0 && s_bad) { ws.hdr->fallback = 1; ws.hdr->chain_bad = 1; }
}

// ------------------------------------------------------------------ positions
__device__ __forceinline__ double pos_sys(i64 i, double u, double Nd) { return __ddiv_rn(__dadd_rn(u, (double)i), Nd); }

// number of positions strictly below c (systematic)
__device__ __forceinline__ i64 count_below_sys(double c, double u, i64 N, double Nd, double tau)
{
    double v = __dadd_rn(__dmul_rn(c, Nd), -u);
    double fl = floor(v);
    double fr = v - fl;
    if (fr > tau && fr < 1.0 - tau && fabs(v) < 4.0e15) {
        double g = fl + 1.0;
        if (g < 0.0) g = 0.0;
        if (g > Nd) g = Nd;
        return (i64)g;
    }
    double g0d = fl + 1.0;
    if (!(g0d > 0.0)) g0d = 0.0;
    if (g0d > Nd) g0d = Nd;
    i64 g = (i64)g0d;
    while (g < N && pos_sys(g, u, Nd) < c) g++;
    while (g > 0 && !(pos_sys(g - 1, u, Nd) < c)) g--;
    return g;
}

__device__ __forceinline__ double pos_str(i64 i, const double *U, double Nd) { return __ddiv_rn(__dadd_rn(U[i], (double)i), Nd); }

// number of positions strictly below c (stratified; positions are non-decreasing in i)
__device__ __forceinline__ i64 count_below_str(double c, const double *U, i64 N, double Nd)
{
    double v = floor(__dmul_rn(c, Nd));
    if (!(v > 1.0)) v = 1.0;
    if (v > Nd) v = Nd;
    i64 g = (i64)v - 1;                      // candidates start two below the real-valued crossing
    if (g < 0) g = 0;
    while (g < N && pos_str(g, U, Nd) < c) g++;
    while (g > 0 && !(pos_str(g - 1, U, Nd) < c)) g--;
    return g;
}

// ------------------------------------------------------------------ pass E: emit indexes
struct EmitShared {
    TileShared ts;
    int hi[TILE];                 // output end (exclusive) of every element, relative to tile_lo
    int ebuf[EXPAND];
    double segstate[RMAX + 1];
    int segk[RMAX + 1];
    i64 end0[RMAX + 1], end1[RMAX + 1];
    double wraw[RMAX];
    int warp_max[BLOCK / 32];
    i64 tile_lo;
    int bad;
};

__global__ void __launch_bounds__(BLOCK) k_emit(Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitShared &sm = *reinterpret_cast<EmitShared *>(smem_raw);
    const Ws &ws = p.ws;
    if (ws.hdr->fallback) return;
    const int t = blockIdx.x;
    const int tid = threadIdx.x;
    const double Nd = (double)p.n;
    TileAn an;
    analyse_tile(p, t, an, sm.ts);
    const double S_in = ws.S_in[t];
    // segment start states (exact)
    if (an.nraw > 0) {
        sm.ts.first_raw[tid] = (an.ek[0] == -1);
        if (tid == 0) sm.ts.first_raw[BLOCK] = 1;
        for (int q = tid; q <= RMAX; q += BLOCK) { sm.segk[q] = -1; sm.end0[q] = 0; sm.end1[q] = 0; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int seg = an.inc[k].cnt;
            if (an.ek[k] == -1) sm.wraw[seg - 1] = an.w[k];
            else {
                bool next_raw = (k + 1 < IPT) ? (an.ek[k + 1] == -1) : (sm.ts.first_raw[tid + 1] != 0);
                if (next_raw) {
                    sm.segk[seg] = an.inc[k].k == -2 ? -1 : an.inc[k].k;
                    sm.end0[seg] = an.inc[k].d0; sm.end1[seg] = an.inc[k].d1;
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            int bad = 0;
            walk_slot(sm.segk, sm.end0, sm.end1, sm.wraw, an.nraw, S_in, &bad, sm.segstate);
            sm.bad = bad;
        }
    } else if (tid == 0) {
        sm.segstate[0] = S_in;
        sm.bad = 0;
    }
    if (tid == 0) {
        sm.tile_lo = p.U ? count_below_str(S_in, p.U, p.n, Nd) : count_below_sys(S_in, p.u, p.n, Nd, p.tau);
    }
    __syncthreads();
    const i64 tile_lo = sm.tile_lo;
    // exact c_j and the output range end of every element
    const i64 jbase = (i64)t * TILE + (i64)tid * IPT;
    int bad = 0;
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const int seg = an.inc[k].cnt;
        double S0 = sm.segstate[seg];
        // raw element: the segment it opens starts at its own result; only zeros so far: unchanged
        double c = (an.ek[k] == -1 || an.inc[k].k == -2) ? S0 : apply_map(S0, Map{an.inc[k].d0, an.inc[k].d1}, an.inc[k].k, &bad);
        i64 h;
        if (jbase + k < p.n) {
            h = p.U ? count_below_str(c, p.U, p.n, Nd) : count_below_sys(c, p.u, p.n, Nd, p.tau);
        } else {
            h = -1;       // filled in below: padding elements own no output
        }
        i64 rel = (h < 0) ? -1 : (h - tile_lo);
        sm.hi[tid * IPT + k] = (int)rel;
    }
    if (bad) ws.hdr->chain_bad = 2;      // cannot happen after pass D verified the tile; recorded for tests
    __syncthreads();
    // padding elements inherit the end of the last real element
    {
        const i64 last_real = p.n - 1 - (i64)t * TILE;      // index in tile of the last real element
        if (last_real < TILE - 1) {
            int hv = sm.hi[last_real];
            __syncthreads();
            for (int q = tid; q < TILE; q += BLOCK) if (q > last_real) sm.hi[q] = hv;
        }
    }
    __syncthreads();
    const int tile_cnt = sm.hi[TILE - 1];                   // outputs owned by this tile
    if (t == ws.T - 1 && tid == 0) {
        if (p.cumsum_last) *p.cumsum_last = ws.S_in[ws.T];
        i64 O1 = tile_lo + tile_cnt;
        if (O1 < p.n) {                                     // resampling.py:145 would raise IndexError
            ws.hdr->overflow = (int)(p.n - O1 > 0x7fffffff ? 0x7fffffff : p.n - O1);
            int r = atomicAdd(&ws.hdr->n_runs, 1);
            if (r < ws.max_runs) ws.runs[r] = Run{O1, p.n, (int)(p.n - 1), 0};
        }
    }
    // expansion: outputs [tile_lo + cs, tile_lo + ce) per pass
    int cs = 0;
    while (cs < tile_cnt) {
        // owner of output cs: first element with hi > cs
        int lo_s = 0, hi_s = TILE - 1;
        while (lo_s < hi_s) {
            int mid = (lo_s + hi_s) >> 1;
            if (sm.hi[mid] > cs) hi_s = mid; else lo_s = mid + 1;
        }
        const int owner = lo_s;
        const int owner_end = sm.hi[owner];
        if (owner_end - cs >= BIGRUN) {
            if (tid == 0) {
                int r = atomicAdd(&ws.hdr->n_runs, 1);
                if (r < ws.max_runs) ws.runs[r] = Run{tile_lo + cs, tile_lo + owner_end, (int)((i64)t * TILE + owner), 0};
                else ws.hdr->fallback = 1;
            }
            cs = owner_end;
            continue;
        }
        const int ce = min(tile_cnt, cs + EXPAND);
        for (int q = tid; q < EXPAND; q += BLOCK) sm.ebuf[q] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const int e = tid * IPT + k;
            const int h = sm.hi[e];
            const int l = (e == 0) ? 0 : sm.hi[e - 1];
            if (h > l && h > cs && l < ce) {
                int start = l > cs ? l : cs;
                sm.ebuf[start - cs] = e + 1;
            }
        }
        __syncthreads();
        // inclusive max-scan over ebuf[0 .. ce-cs): 16 consecutive entries per thread
        {
            constexpr int PER = EXPAND / BLOCK;
            int v[PER];
            int m = 0;
#pragma unroll
            for (int q = 0; q < PER; q++) { int x = sm.ebuf[tid * PER + q]; m = x > m ? x : m; v[q] = m; }
            const int lane = tid & 31, wid = tid >> 5;
            int inc = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc = y > inc ? y : inc; }
            if (lane == 31) sm.warp_max[wid] = inc;
            __syncthreads();
            int basem = 0;
            for (int k2 = 0; k2 < wid; k2++) basem = max(basem, sm.warp_max[k2]);
            int prev = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) prev = 0;
            basem = max(basem, prev);
#pragma unroll
            for (int q = 0; q < PER; q++) sm.ebuf[tid * PER + q] = max(v[q], basem);
        }
        __syncthreads();
        const int base_j = (int)((i64)t * TILE) - 1;
        for (int q = tid; q < ce - cs; q += BLOCK) p.idx[tile_lo + cs + q] = base_j + sm.ebuf[q];
        __syncthreads();
        cs = ce;
    }
}

// ------------------------------------------------------------------ pass F: long runs
__global__ void __launch_bounds__(256) k_fill_runs(Params p)
{
    const Ws &ws = p.ws;
    if (ws.hdr->fallback) return;
    int nr = ws.hdr->n_runs;
    if (nr > ws.max_runs) nr = ws.max_runs;
    for (int r = 0; r < nr; r++) {
        const Run run = ws.runs[r];
        for (i64 i = run.lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < run.hi; i += (i64)gridDim.x * blockDim.x)
            p.idx[i] = run.j;
    }
}

// ------------------------------------------------------------------ pass G: literal sequential fallback
__global__ void k_sequential(Params p, double carry)
{
    const Ws &ws = p.ws;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!ws.hdr->fallback) {
        if (p.info) { p.info[0] = ws.hdr->overflow; p.info[1] = 0; p.info[2] = ws.hdr->n_unclean; p.info[3] = ws.hdr->n_runs; p.info[4] = ws.hdr->chain_bad; }
        return;
    }
    // resampling.py:141-149 — cumulative sum and two-pointer merge, one element at a time
    const double Nd = (double)p.n;
    i64 i = 0, j = 0;
    double c = __dadd_rn(carry, p.w[0]);
    if (carry == 0.0) c = p.w[0];
    int overflow = 0;
    while (i < p.n) {
        double pos = p.U ? pos_str(i, p.U, Nd) : pos_sys(i, p.u, Nd);
        if (pos < c) { p.idx[i] = (int)j; i++; }
        else {
            j++;
            if (j >= p.n) { overflow = (int)(p.n - i); for (; i < p.n; i++) p.idx[i] = (int)(p.n - 1); break; }
            c = __dadd_rn(c, p.w[j]);
        }
    }
    if (p.cumsum_last) {
        for (i64 q = j + 1; q < p.n; q++) c = __dadd_rn(c, p.w[q]);
        *p.cumsum_last = c;
    }
    if (p.info) { p.info[0] = overflow; p.info[1] = 1; p.info[2] = ws.hdr->n_unclean; p.info[3] = ws.hdr->n_runs; p.info[4] = ws.hdr->chain_bad; }
}

// ------------------------------------------------------------------ weight sum / scale
__global__ void __launch_bounds__(BLOCK) k_scale(i64 n, const double *w, const double *div, double *out)
{
    const double d = *div;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = __ddiv_rn(w[i], d);
}

__global__ void __launch_bounds__(CHAIN_THREADS) k_sum_tiles(const double *tile_sum, int T, double *out)
{
    __shared__ double sh[CHAIN_THREADS];
    const int per = (T + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const int a = threadIdx.x * per, b = min(T, a + per);
    double s = 0.0;
    for (int t = a; t < b; t++) s += tile_sum[t];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = CHAIN_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

int run(i64 n, const double *w, double u, const double *U, int *idx, void *workspace, size_t ws_bytes,
        int *info, double *cumsum_last, cudaStream_t s)
{
    if (n < 0) { set_error("n < 0"); return BKE_ERR_BAD_ARG; }
    if (n == 0) return BKE_OK;
    if (n >= ((i64)1 << 31)) { set_error("n must be < 2^31 (indexes are int32, resampling.py:141)"); return BKE_ERR_BAD_ARG; }
    if (!w || !idx || !workspace) { set_error("weights, indexes and workspace must be non-NULL"); return BKE_ERR_BAD_ARG; }
    if (!U && !(u >= 0.0 && u < 1.0)) { set_error("u must be in [0, 1)"); return BKE_ERR_BAD_ARG; }
    const size_t need = carve(n, nullptr, nullptr);
    if (ws_bytes < need) { set_error("workspace too small: %zu < %zu", ws_bytes, need); return BKE_ERR_BAD_ARG; }
    if (reinterpret_cast<uintptr_t>(workspace) & 255) { set_error("workspace must be 256-byte aligned"); return BKE_ERR_BAD_ARG; }
    Params p;
    carve(n, (unsigned char *)workspace, &p.ws);
    p.w = w; p.n = n; p.u = u; p.U = U; p.idx = idx; p.info = info; p.cumsum_last = cumsum_last;
    p.eps = ldexp((double)n + 4096.0, -52);
    double tau = ldexp((double)n, -46);
    p.tau = tau > 1e-6 ? tau : 1e-6;
    if (check_cuda(cudaMemsetAsync(p.ws.hdr, 0, sizeof(Header), s), "memset header")) return BKE_ERR_CUDA;
    const int T = p.ws.T;
    const int emit_smem = (int)sizeof(EmitShared);
    if (check_cuda(cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, emit_smem), "cudaFuncSetAttribute")) return BKE_ERR_CUDA;
    k_tile_sums<<<T, BLOCK, 0, s>>>(p);
    k_scan_tiles<<<1, CHAIN_THREADS, 0, s>>>(p);
    k_tile_maps<<<T, BLOCK, 0, s>>>(p);
    k_chain<<<1, CHAIN_THREADS, 0, s>>>(p, 0.0);
    k_emit<<<T, BLOCK, emit_smem, s>>>(p);
    k_fill_runs<<<sm_count() * 4, 256, 0, s>>>(p);
    k_sequential<<<1, 32, 0, s>>>(p, 0.0);
    return check_cuda(cudaGetLastError(), "resample launch");
}

}  // namespace rs
}  // namespace bke

using namespace bke;

extern "C" {

size_t bke_resample_workspace_bytes(int64_t n)
{
    if (n <= 0) return 256;
    return rs::carve(n, nullptr, nullptr);
}

int bke_systematic_resample(int64_t n, const double *weights, double u, int32_t *indexes, void *workspace,
                            size_t workspace_bytes, int32_t *info, double *cumsum_last, void *stream)
{
    return rs::run(n, weights, u, nullptr, indexes, workspace, workspace_bytes, info, cumsum_last, (cudaStream_t)stream);
}

int bke_stratified_resample(int64_t n, const double *weights, const double *uniforms, int32_t *indexes,
                            void *workspace, size_t workspace_bytes, int32_t *info, double *cumsum_last, void *stream)
{
    if (n > 0 && !uniforms) { set_error("uniforms is NULL"); return BKE_ERR_BAD_ARG; }
    return rs::run(n, weights, 0.0, uniforms, indexes, workspace, workspace_bytes, info, cumsum_last, (cudaStream_t)stream);
}

int bke_weights_sum(int64_t n, const double *weights, double *sum_out, void *workspace, size_t workspace_bytes,
                    void *stream)
{
    if (n < 0 || !sum_out) { set_error("bad arguments"); return BKE_ERR_BAD_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) return check_cuda(cudaMemsetAsync(sum_out, 0, sizeof(double), s), "memset");
    if (!weights || !workspace) { set_error("weights and workspace must be non-NULL"); return BKE_ERR_BAD_ARG; }
    const size_t need = rs::carve(n, nullptr, nullptr);
    if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu", workspace_bytes, need); return BKE_ERR_BAD_ARG; }
    rs::Params p;
    rs::carve(n, (unsigned char *)workspace, &p.ws);
    p.w = weights; p.n = n;
    if (check_cuda(cudaMemsetAsync(p.ws.hdr, 0, sizeof(rs::Header), s), "memset header")) return BKE_ERR_CUDA;
    rs::k_tile_sums<<<p.ws.T, rs::BLOCK, 0, s>>>(p);
    rs::k_sum_tiles<<<1, rs::CHAIN_THREADS, 0, s>>>(p.ws.tile_sum, p.ws.T, sum_out);
    return check_cuda(cudaGetLastError(), "weights_sum launch");
}

int bke_weights_scale(int64_t n, const double *weights, const double *divisor, double *weights_out, void *stream)
{
    if (n < 0) { set_error("n < 0"); return BKE_ERR_BAD_ARG; }
    if (n == 0) return BKE_OK;
    if (!weights || !divisor || !weights_out) { set_error("NULL argument"); return BKE_ERR_BAD_ARG; }
    int64_t blocks = (n + rs::BLOCK * 8 - 1) / (rs::BLOCK * 8);
    int64_t cap = (int64_t)sm_count() * 16;
    rs::k_scale<<<(unsigned)(blocks < cap ? blocks : cap), rs::BLOCK, 0, (cudaStream_t)stream>>>(n, weights, divisor, weights_out);
    return check_cuda(cudaGetLastError(), "weights_scale launch");
}

}  // extern "C"
