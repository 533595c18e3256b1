// K2 -- ball query (reference: model/pointnet2_utils.py:70-90 on top of square_distance :5-26), plus the
// stand-alone square_distance kernel.
//
// One CTA handles one cloud x a tile of centres.  The cloud (SoA coordinates + |p|^2, computed once per CTA
// in the layout-dependent summation order) is staged in shared memory; a warp owns four centres at a time (two for
// C > 4) and scans the points in index order.  The "first K in ascending index order" semantics of the reference's
// sort (:86) falls out of ballots without any sort, and the scan stops as soon as every list in flight holds K
// neighbours.  The [B,S,N] distance matrix and the [B,S,N] int64 index tensor of the reference never exist.
//
// Kernels of this file:
//   ball_query_masks_kernel / ball_query_multi_masks_kernel  (C <= 4, cloud staged whole: the default)  -- a trip of 128
//       points only parks its four ballots in lanes; the lists are resolved once per 1024 points (deferred bookkeeping);
//   ball_query_kernel / ball_query_multi_kernel  (C = 10; clouds beyond shared memory; PN2_BQ_V1=1)  -- 64 points per
//       trip, ballot + popc give each in-radius lane its output slot inside the trip;
//   ball_query_cell_kernel  (PN2_BQ_CELLS=1)  -- exact grid cull, measured not faster at the reference's densities;
//   ball_query_generic_kernel  -- any C <= 16, lists beyond 1024 entries: one thread per centre, no staging.
//
// Exactness contract: dot = fl(s0*p0), dot = fma(s_k, p_k, dot) ascending k (MKL sgemm, :23);
// d = fl(fl(-2*dot + |s|^2) + |p|^2); in radius <=> !(d > r2).  Every kernel evaluates a pair with exactly this chain.
#include <stdlib.h>
#include <type_traits>
#include "pn2_common.cuh"

namespace pn2 {

constexpr int BQ_WARPS_SMALL = 8, BQ_WARPS_BIG = 32;     // warps per CTA; 32 when the staged cloud allows only one CTA per SM
// centres scanned together by one warp (register budget: CPB * (2C + 2))
template <int C> struct BQ_CPB { static constexpr int value = C <= 4 ? 4 : 2; };

template <int C>
__device__ __forceinline__ float sqnorm(const float (&v)[C], bool contig, bool tail)
{
    float t[C];
#pragma unroll
    for (int k = 0; k < C; ++k) t[k] = __fmul_rn(v[k], v[k]);
    return contig ? sum_order<C, true>(t) : sum_order<C, false>(t, tail);
}

// distances of centre s to the two consecutive points n, n+1 (n even)
template <int C>
__device__ __forceinline__ float2 pair_dist2(const f32x2 (&s)[C], const float *sp, int stride, int n, f32x2 s2, f32x2 m2)
{
    f32x2 dot = mul2(s[0], *reinterpret_cast<const f32x2 *>(sp + n));
#pragma unroll
    for (int k = 1; k < C; ++k) dot = fma2(s[k], *reinterpret_cast<const f32x2 *>(sp + k * stride + n), dot);
    f32x2 d = mul2(m2, dot);
    d = add2(d, s2);
    return unpk2(add2(d, *reinterpret_cast<const f32x2 *>(sp + (size_t)C * stride + n)));
}

template <int C>
__device__ __forceinline__ float pair_dist(const float (&s)[C], const float *sp, int stride, int n, float s2, float p2)
{
    float dot = __fmul_rn(s[0], sp[n]);
#pragma unroll
    for (int k = 1; k < C; ++k) dot = __fmaf_rn(s[k], sp[k * stride + n], dot);
    float d = __fmul_rn(-2.0f, dot);
    d = __fadd_rn(d, s2);
    return __fadd_rn(d, p2);
}

// smem: [C+1][NCH] floats (coords SoA then |p|^2), then per-warp staging rows [BQ_WARPS][K] int
template <int C, int BQ_WARPS>
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float *__restrict__ centres, int64_t csb, int64_t csn, int64_t csc,
                  const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                  int S, int N, int NCH, int cpw /*centres per warp*/, float r2, int K,
                  int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, int32_t *__restrict__ empty_flag)
{
    constexpr int BQ_THREADS = BQ_WARPS * 32;
    extern __shared__ float smem[];
    float *sp = smem;                       // [C][NCH]
    float *sp2 = smem + (size_t)C * NCH;    // [NCH]
    constexpr int CPB = BQ_CPB<C>::value;
    int *stage = (int *)(sp2 + NCH) + (threadIdx.x >> 5) * (CPB * K);      // [CPB][K] per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int s_begin = (blockIdx.x * BQ_WARPS + warp) * cpw;
    const float *P = points + (int64_t)b * sb;
    const bool pcontig = (sc == 1);
    const unsigned lt_mask = (1u << lane) - 1u;

    const bool single_chunk = (NCH >= N);
    // per-warp state for the multi-chunk path lives in registers for cpw == 1 only; otherwise the host
    // guarantees single_chunk (see launcher).
    int cnt = 0, first = N;
    float s[C];
    float s2 = 0.f;
    const bool have = s_begin < S;
    if (!single_chunk && have) {
#pragma unroll
        for (int k = 0; k < C; ++k) s[k] = __ldg(centres + (int64_t)b * csb + (int64_t)s_begin * csn + (int64_t)k * csc);
        s2 = sqnorm<C>(s, csc == 1, s_begin >= (S & ~31));
    }

    for (int n_base = 0; n_base < N; n_base += NCH) {
        const int nc = min(NCH, N - n_base);
        if (n_base) __syncthreads();
        // The pose heads' "coordinates" are xyz + a one-hot category vector (rotation_ssg.py:27): channels 3.. are the same
        // finite constants e_k for every point (and centre) of the cloud, at most one of them non-zero.  The dot-product chain
        // then continues past k = 2 with fma(e_k, e_k, dot): a no-op for e_k = 0, one fused add of e_k^2 for the non-zero one --
        // no shared-memory reads, three channels of arithmetic, bit-identical.  |p|^2 and |s|^2 keep their full sums.
        bool tail_same = C > 4 && single_chunk;
        // one CTA per SM (large clouds): four points in flight per thread, the staging is a chain of global-load latencies
        // (configs[2]: 135 -> 131 us; with several CTAs per SM the other CTAs hide it and the unrolled loop is 2 % slower)
#pragma unroll((C <= 4 && BQ_WARPS == BQ_WARPS_BIG) ? 4 : 1)
        for (int n = tid; n < nc; n += BQ_THREADS) {
            float v[C];
#pragma unroll
            for (int k = 0; k < C; ++k) {
                v[k] = __ldg(P + (int64_t)(n_base + n) * sn + (int64_t)k * sc);
                sp[k * NCH + n] = v[k];
                if (C > 4 && k >= 3) {
                    const float ref = __ldg(P + (int64_t)k * sc);       // point 0
                    tail_same = tail_same && fabsf(ref) <= 3.0e38f && v[k] == ref;
                }
            }
            sp2[n] = sqnorm<C>(v, pcontig, n_base + n >= (N & ~31));
        }
        if (single_chunk) {
            // pad the rows to a multiple of 64 points with zero coordinates and |p|^2 = +inf: d = +inf, never in
            // radius (a NaN centre finds its K neighbours among the real points first), so the scan needs no bounds
            for (int n = nc + tid; n < NCH; n += BQ_THREADS) {
#pragma unroll
                for (int k = 0; k < C; ++k) sp[k * NCH + n] = 0.f;
                sp2[n] = INFINITY;
            }
        }
        const bool tail_uniform = __syncthreads_and(tail_same ? 1 : 0) != 0 && C > 4;
        // the tail's constants: usable by the short chain when at most one is non-zero
        float e_nz = 0.f;
        bool tail_short = tail_uniform;
        if (C > 4 && tail_uniform) {
            int nz = 0;
#pragma unroll
            for (int k = 3; k < C; ++k) { const float ek = sp[k * NCH]; if (ek != 0.f) { ++nz; e_nz = ek; } }
            tail_short = nz <= 1;
        }

        if (single_chunk) {
            // CPB centres per pass: every pair of points read from shared memory is tested against all of them
            // (shared-memory traffic, the limiter of the one-centre scan, drops by CPB; so do the loop instructions)
            for (int ci = 0; ci < cpw; ci += CPB) {
                f32x2 sv[CPB][C], s2v[CPB];
                int cn[CPB];
                bool grp_short = tail_short;
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    const int sidx = s_begin + ci + c;
                    const bool act = ci + c < cpw && sidx < S;
                    float sc_[C];
#pragma unroll
                    for (int k = 0; k < C; ++k) sc_[k] = act ? __ldg(centres + (int64_t)b * csb + (int64_t)sidx * csn + (int64_t)k * csc) : 0.f;
                    const float q2 = sqnorm<C>(sc_, csc == 1, sidx >= (S & ~31));
#pragma unroll
                    for (int k = 0; k < C; ++k) sv[c][k] = pk2(sc_[k], sc_[k]);
                    s2v[c] = pk2(q2, q2);
                    cn[c] = act ? 0 : K;              // an inactive slot counts as finished
                    if (C > 4 && act) {               // the centres must carry the cloud's tail as well (they do when they are its points)
#pragma unroll
                        for (int k = 3; k < C; ++k) grp_short = grp_short && sc_[k] == sp[k * NCH];
                    }
                }
                const f32x2 m2v = pk2(-2.0f, -2.0f);
                if (C > 4 && grp_short) {
                    // ---- three channels from shared memory + one constant fused add (see the staging comment) ----------------
                    const f32x2 ev = pk2(e_nz, e_nz);
                    for (int n0 = 0; n0 < NCH; n0 += 64) {
                        bool all_done = true;
#pragma unroll
                        for (int c = 0; c < CPB; ++c) all_done = all_done && cn[c] >= K;
                        if (all_done) break;
                        const int na = n0 + 2 * lane;
                        f32x2 pv[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) pv[k] = *reinterpret_cast<const f32x2 *>(sp + k * NCH + na);
                        const f32x2 p2v = *reinterpret_cast<const f32x2 *>(sp2 + na);
#pragma unroll
                        for (int c = 0; c < CPB; ++c) {
                            f32x2 dot = mul2(sv[c][0], pv[0]);
                            dot = fma2(sv[c][1], pv[1], dot);
                            dot = fma2(sv[c][2], pv[2], dot);
                            if (e_nz != 0.f) dot = fma2(ev, ev, dot);
                            const float2 d = unpk2(add2(fma2(m2v, dot, s2v[c]), p2v));   // fl(-2 dot + |s|^2): -2 dot is exact, one rounding either way
                            const bool ina = !(d.x > r2), inb = !(d.y > r2);
                            const unsigned ma = __ballot_sync(0xffffffffu, ina);
                            const unsigned mb = __ballot_sync(0xffffffffu, inb);
                            if (ma | mb) {
                                int *stg = stage + c * K;
                                const int pa = cn[c] + __popc(ma & lt_mask) + __popc(mb & lt_mask);
                                if (ina && pa < K) stg[pa] = na;
                                if (inb && pa + (int)ina < K) stg[pa + (int)ina] = na + 1;
                                cn[c] += __popc(ma) + __popc(mb);
                            }
                        }
                    }
                } else {
                const float *prow = sp + 2 * lane;            // this lane's pair of points in channel 0; bumped by 64 per trip
                for (int n0 = 0; n0 < NCH; n0 += 64, prow += 64) {
                    bool all_done = true;
#pragma unroll
                    for (int c = 0; c < CPB; ++c) all_done = all_done && cn[c] >= K;
                    if (all_done) break;
                    // 64 points per trip, two consecutive ones per lane (rows are padded to a multiple of 64)
                    const int na = n0 + 2 * lane;
                    f32x2 pv[C];
#pragma unroll
                    for (int k = 0; k < C; ++k) pv[k] = *reinterpret_cast<const f32x2 *>(prow + k * NCH);
                    const f32x2 p2v = *reinterpret_cast<const f32x2 *>(prow + C * NCH);
#pragma unroll
                    for (int c = 0; c < CPB; ++c) {
                        f32x2 dot = mul2(sv[c][0], pv[0]);
#pragma unroll
                        for (int k = 1; k < C; ++k) dot = fma2(sv[c][k], pv[k], dot);
                        const float2 d = unpk2(add2(fma2(m2v, dot, s2v[c]), p2v));   // fl(-2 dot + |s|^2): -2 dot is exact, one rounding either way
                        const bool ina = !(d.x > r2), inb = !(d.y > r2);
                        const unsigned ma = __ballot_sync(0xffffffffu, ina);
                        const unsigned mb = __ballot_sync(0xffffffffu, inb);
                        if (ma | mb) {                                                              // warp-uniform; most trips have no hit
                            int *stg = stage + c * K;
                            const int pa = cn[c] + __popc(ma & lt_mask) + __popc(mb & lt_mask);     // hits at lower indices
                            if (ina && pa < K) stg[pa] = na;
                            if (inb && pa + (int)ina < K) stg[pa + (int)ina] = na + 1;
                            cn[c] += __popc(ma) + __popc(mb);
                        }
                    }
                }
                }
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    const int sidx = s_begin + ci + c;
                    if (!(ci + c < cpw && sidx < S)) continue;
                    const int *stg = stage + c * K;
                    const int cc = cn[c] > K ? K : cn[c];
                    // hits are staged in index order; a NaN centre also "hits" the padding slots (indices >= N, always last)
                    const int frst = (cc && stg[0] < N) ? stg[0] : N;
                    if (frst == N && lane == 0 && empty_flag) atomicOr(empty_flag, 1);
                    const int64_t o = ((int64_t)b * S + sidx) * K;
                    for (int j = lane; j < K; j += 32) {
                        const int v = (j < cc && stg[j] < N) ? stg[j] : frst;
                        if (idx64) idx64[o + j] = v;
                        if (idx32) idx32[o + j] = v;
                    }
                }
                __syncwarp();
            }
        } else if (have) {
            for (int n0 = 0; n0 < nc && cnt < K; n0 += 32) {
                const int na = n0 + lane;
                const float da = (na < nc) ? pair_dist<C>(s, sp, NCH, na, s2, sp2[na]) : 0.f;
                const bool ina = (na < nc) && !(da > r2);
                const unsigned ma = __ballot_sync(0xffffffffu, ina);
                if (first == N && ma) first = n_base + n0 + __ffs(ma) - 1;
                const int pa = cnt + __popc(ma & lt_mask);
                if (ina && pa < K) stage[pa] = n_base + na;
                cnt += __popc(ma);
            }
        }
    }
    if (!single_chunk && have) {
        if (cnt > K) cnt = K;
        if (cnt == 0 && lane == 0 && empty_flag) atomicOr(empty_flag, 1);
        __syncwarp();
        const int64_t o = ((int64_t)b * S + s_begin) * K;
        for (int j = lane; j < K; j += 32) {
            const int v = (j < cnt) ? stage[j] : first;
            if (idx64) idx64[o + j] = v;
            if (idx32) idx32[o + j] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Ball query with DEFERRED hit bookkeeping (C <= 4, cloud staged whole; the default for those shapes).
//
// The scan kernel above is issue-bound and a third of what it issues is bookkeeping: per (trip, centre) with a hit it computes
// two prefix popcounts, two slot addresses and two predicated stores (~19 instructions, three trips in four at N = 1024,
// r = 0.2), on top of 13 for the distances and ballots.  Here a trip only RECORDS: a trip covers 128 points, four per lane
// (one LDS.128 per channel), staged so that component j of lane l is point 128 t + 32 j + l -- ballot j of trip t is then the
// hit mask of the 32 CONSECUTIVE points of word 4 t + j, and it is parked in lane 4 t + j (one predicated move).  After 8 trips
// = 1024 points every lane holds one word of the 1024-bit hit mask of each of the four centres in flight, in index order, and
// the lists are resolved once: prefix sum of the popcounts over the lanes, then every lane emits its own set bits at its base
// slot.  "First K in ascending index order" (:86-89) is unchanged -- the same pairs get the same distance arithmetic, only the
// place where a hit's slot is computed moves.  Early exit: after every trip the per-centre totals (one REDUX over the parked
// words) are compared with K.  Per 1024 points x 4 centres: ~1050 instructions against ~1800.
template <int C, int BQ_WARPS>
__global__ void __launch_bounds__(BQ_WARPS * 32, BQ_WARPS == BQ_WARPS_BIG ? 1 : 4)
ball_query_masks_kernel(const float *__restrict__ centres, int64_t csb, int64_t csn, int64_t csc,
                        const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                        int S, int N, int NCH /* multiple of 128 */, int cpw, float r2, int K,
                        int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, int32_t *__restrict__ empty_flag)
{
    static_assert(C <= 4, "one float4 of points per channel per lane");
    constexpr int BQ_THREADS = BQ_WARPS * 32;
    constexpr int CPB = 4;
    extern __shared__ float smem[];
    float *sp = smem;                       // [C][NCH], permuted inside every block of 128 points (see above)
    float *sp2 = smem + (size_t)C * NCH;    // [NCH]
    int *stage = (int *)(sp2 + NCH) + (threadIdx.x >> 5) * (CPB * K);      // [CPB][K] per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int s_begin = (blockIdx.x * BQ_WARPS + warp) * cpw;
    const float *P = points + (int64_t)b * sb;
    const bool pcontig = (sc == 1);

    // ---- stage the cloud: same arithmetic as ball_query_kernel; point n of a 128-block goes to slot 4 (n % 32) + n / 32 -----
#pragma unroll(BQ_WARPS == BQ_WARPS_BIG ? 4 : 1)
    for (int n = tid; n < N; n += BQ_THREADS) {
        float v[C];
        const int pos = (n & ~127) | ((n & 31) << 2) | ((n >> 5) & 3);
#pragma unroll
        for (int k = 0; k < C; ++k) {
            v[k] = __ldg(P + (int64_t)n * sn + (int64_t)k * sc);
            sp[k * NCH + pos] = v[k];
        }
        sp2[pos] = sqnorm<C>(v, pcontig, n >= (N & ~31));
    }
    // padding up to NCH: zero coordinates and |p|^2 = +inf (d = +inf: never in radius; a NaN centre's hits beyond N are
    // masked off when the lists are resolved)
    for (int n = N + tid; n < NCH; n += BQ_THREADS) {
        const int pos = (n & ~127) | ((n & 31) << 2) | ((n >> 5) & 3);
#pragma unroll
        for (int k = 0; k < C; ++k) sp[k * NCH + pos] = 0.f;
        sp2[pos] = INFINITY;
    }
    __syncthreads();

    const f32x2 m2v = pk2(-2.0f, -2.0f);
    const bool odd = (lane & 1) != 0, upper = (lane & 2) != 0;
    for (int ci = 0; ci < cpw; ci += CPB) {
        if (s_begin + ci >= S) break;                                   // warp-uniform
        f32x2 sv[CPB][C], s2v[CPB];
        int cn[CPB];                                                    // hits resolved so far (warp-uniform)
        unsigned hm[CPB];                                               // this lane's word of the current 1024-point window
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            const int sidx = s_begin + ci + c;
            const bool act = ci + c < cpw && sidx < S;
            float sc_[C];
#pragma unroll
            for (int k = 0; k < C; ++k) sc_[k] = act ? __ldg(centres + (int64_t)b * csb + (int64_t)sidx * csn + (int64_t)k * csc) : 0.f;
            const float q2 = sqnorm<C>(sc_, csc == 1, sidx >= (S & ~31));
#pragma unroll
            for (int k = 0; k < C; ++k) sv[c][k] = pk2(sc_[k], sc_[k]);
            s2v[c] = pk2(q2, q2);
            cn[c] = act ? 0 : K;                                        // an inactive slot counts as finished
        }
        for (int w0 = 0; w0 < NCH; w0 += 1024) {                        // windows of 1024 points = 32 mask words
            const int trips = min(8, (NCH - w0) >> 7);
#pragma unroll
            for (int c = 0; c < CPB; ++c) hm[c] = 0u;
            const float *prow = sp + w0 + 4 * lane;
            for (int t = 0; t < trips; ++t, prow += 128) {
                float4 pv[C];
#pragma unroll
                for (int k = 0; k < C; ++k) pv[k] = *reinterpret_cast<const float4 *>(prow + k * NCH);
                const float4 p2 = *reinterpret_cast<const float4 *>(prow + C * NCH);
                const f32x2 p2a = pk2(p2.x, p2.y), p2b = pk2(p2.z, p2.w);
                const bool mine = (lane >> 2) == t;                     // lanes 4 t .. 4 t + 3 park the four ballots of this trip
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    f32x2 da = mul2(sv[c][0], pk2(pv[0].x, pv[0].y));
                    f32x2 db = mul2(sv[c][0], pk2(pv[0].z, pv[0].w));
#pragma unroll
                    for (int k = 1; k < C; ++k) {
                        da = fma2(sv[c][k], pk2(pv[k].x, pv[k].y), da);
                        db = fma2(sv[c][k], pk2(pv[k].z, pv[k].w), db);
                    }
                    const float2 d01 = unpk2(add2(fma2(m2v, da, s2v[c]), p2a));     // fl(-2 dot + |s|^2): -2 dot is exact, one rounding either way
                    const float2 d23 = unpk2(add2(fma2(m2v, db, s2v[c]), p2b));
                    const unsigned m0 = __ballot_sync(0xffffffffu, !(d01.x > r2));
                    const unsigned m1 = __ballot_sync(0xffffffffu, !(d01.y > r2));
                    const unsigned m2 = __ballot_sync(0xffffffffu, !(d23.x > r2));
                    const unsigned m3 = __ballot_sync(0xffffffffu, !(d23.y > r2));
                    const unsigned pick = odd ? (upper ? m3 : m1) : (upper ? m2 : m0);       // ballot (lane & 3): selects, no branches
                    hm[c] = mine ? pick : hm[c];
                }
                bool all_done = true;
#pragma unroll
                for (int c = 0; c < CPB; ++c) all_done = all_done && cn[c] + (int)__reduce_add_sync(0xffffffffu, (unsigned)__popc(hm[c])) >= K;
                if (all_done) break;
            }
            // ---- resolve the window: lane l owns points w0 + 32 l .. + 31 ------------------------------------------------------
            const int nb = w0 + 32 * lane;
            const unsigned valid = nb + 32 <= N ? 0xffffffffu : (nb >= N ? 0u : (1u << (N - nb)) - 1u);
            bool all_full = true;
#pragma unroll
            for (int c = 0; c < CPB; ++c) {
                unsigned m = hm[c] & valid;
                const int pc = __popc(m);
                int inc = pc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int total = __shfl_sync(0xffffffffu, inc, 31);
                if (cn[c] < K && total) {                               // warp-uniform
                    int pos = cn[c] + inc - pc;
                    int *stg = stage + c * K;
                    while (m && pos < K) { stg[pos++] = nb + __ffs(m) - 1; m &= m - 1; }
                }
                cn[c] = min(K, cn[c] + total);
                all_full = all_full && cn[c] >= K;
            }
            if (all_full) break;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            const int sidx = s_begin + ci + c;
            if (!(ci + c < cpw && sidx < S)) continue;
            const int *stg = stage + c * K;
            const int cc = cn[c];
            const int frst = cc ? stg[0] : N;                           // hits are staged in index order
            if (cc == 0 && lane == 0 && empty_flag) atomicOr(empty_flag, 1);
            const int64_t o = ((int64_t)b * S + sidx) * K;
            for (int j = lane; j < K; j += 32) {
                const int v = (j < cc) ? stg[j] : frst;
                if (idx64) idx64[o + j] = v;
                if (idx32) idx32[o + j] = v;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// Multi-radius ball query: the MSG layers (pointnet2_utils.py:203-204) query the same centres against the same cloud once
// per radius.  The distances do not depend on the radius, so one scan serves all NR of them: d is evaluated once per
// (centre, point) and compared against the ascending radii, largest first -- a point outside the largest ball is outside
// all (the common case costs one compare + two ballots, like the single-radius kernel), and the hit bookkeeping of the
// smaller balls runs only on trips that have a hit in the largest.  The scan of a centre group stops when every list of
// every centre is full.  Same staging, distance arithmetic and output convention as ball_query_kernel (single-chunk
// path: the cloud fits in shared memory), so each list is bit-identical to a separate query.
struct BqMulti {
    float r2[3];
    int K[3];
    int koff[3];            // offset of radius r's staging row inside a centre's [ktot] block
    int ktot;
    int64_t *idx64[3];
    int32_t *idx32[3];
};

template <int C, int BQ_WARPS, int NR>
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_multi_kernel(const float *__restrict__ centres, int64_t csb, int64_t csn, int64_t csc,
                        const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                        int S, int N, int NCH, int cpw, const BqMulti q, int32_t *__restrict__ empty_flag)
{
    constexpr int BQ_THREADS = BQ_WARPS * 32;
    constexpr int CPB = BQ_CPB<C>::value;
    extern __shared__ float smem[];
    float *sp = smem;                       // [C][NCH]
    float *sp2 = smem + (size_t)C * NCH;    // [NCH]
    int *stage = (int *)(sp2 + NCH) + (threadIdx.x >> 5) * (CPB * q.ktot);       // [CPB][ktot] per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int s_begin = (blockIdx.x * BQ_WARPS + warp) * cpw;
    const float *P = points + (int64_t)b * sb;
    const bool pcontig = (sc == 1);
    const unsigned lt_mask = (1u << lane) - 1u;

    // staging: identical to ball_query_kernel's single-chunk path (one-hot tail detection, rows padded to 64 with +inf)
    bool tail_same = C > 4;
    for (int n = tid; n < N; n += BQ_THREADS) {
        float v[C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            v[k] = __ldg(P + (int64_t)n * sn + (int64_t)k * sc);
            sp[k * NCH + n] = v[k];
            if (C > 4 && k >= 3) {
                const float ref = __ldg(P + (int64_t)k * sc);
                tail_same = tail_same && fabsf(ref) <= 3.0e38f && v[k] == ref;
            }
        }
        sp2[n] = sqnorm<C>(v, pcontig, n >= (N & ~31));
    }
    for (int n = N + tid; n < NCH; n += BQ_THREADS) {
#pragma unroll
        for (int k = 0; k < C; ++k) sp[k * NCH + n] = 0.f;
        sp2[n] = INFINITY;
    }
    const bool tail_uniform = __syncthreads_and(tail_same ? 1 : 0) != 0 && C > 4;
    float e_nz = 0.f;
    bool tail_short = tail_uniform;
    if (C > 4 && tail_uniform) {
        int nz = 0;
#pragma unroll
        for (int k = 3; k < C; ++k) { const float ek = sp[k * NCH]; if (ek != 0.f) { ++nz; e_nz = ek; } }
        tail_short = nz <= 1;
    }

    for (int ci = 0; ci < cpw; ci += CPB) {
        f32x2 sv[CPB][C], s2v[CPB];
        int cn[CPB][NR];
        bool grp_short = tail_short;
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            const int sidx = s_begin + ci + c;
            const bool act = ci + c < cpw && sidx < S;
            float sc_[C];
#pragma unroll
            for (int k = 0; k < C; ++k) sc_[k] = act ? __ldg(centres + (int64_t)b * csb + (int64_t)sidx * csn + (int64_t)k * csc) : 0.f;
            const float q2 = sqnorm<C>(sc_, csc == 1, sidx >= (S & ~31));
#pragma unroll
            for (int k = 0; k < C; ++k) sv[c][k] = pk2(sc_[k], sc_[k]);
            s2v[c] = pk2(q2, q2);
#pragma unroll
            for (int r = 0; r < NR; ++r) cn[c][r] = act ? 0 : q.K[r];
            if (C > 4 && act) {
#pragma unroll
                for (int k = 3; k < C; ++k) grp_short = grp_short && sc_[k] == sp[k * NCH];
            }
        }
        const f32x2 m2v = pk2(-2.0f, -2.0f);
        const f32x2 ev = pk2(e_nz, e_nz);

        auto scan = [&](auto short_tag) {
            constexpr bool SHORT = decltype(short_tag)::value;
            constexpr int CE = SHORT ? 3 : C;
            const float *prow = sp + 2 * lane;
            for (int n0 = 0; n0 < NCH; n0 += 64, prow += 64) {
                bool all_done = true;
#pragma unroll
                for (int c = 0; c < CPB; ++c)
#pragma unroll
                    for (int r = 0; r < NR; ++r) all_done = all_done && cn[c][r] >= q.K[r];
                if (all_done) break;
                const int na = n0 + 2 * lane;
                f32x2 pv[CE];
#pragma unroll
                for (int k = 0; k < CE; ++k) pv[k] = *reinterpret_cast<const f32x2 *>(prow + k * NCH);
                const f32x2 p2v = *reinterpret_cast<const f32x2 *>(prow + C * NCH);
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    f32x2 dot = mul2(sv[c][0], pv[0]);
#pragma unroll
                    for (int k = 1; k < CE; ++k) dot = fma2(sv[c][k], pv[k], dot);
                    if (SHORT && e_nz != 0.f) dot = fma2(ev, ev, dot);
                    const float2 d = unpk2(add2(fma2(m2v, dot, s2v[c]), p2v));
                    bool ina = !(d.x > q.r2[NR - 1]), inb = !(d.y > q.r2[NR - 1]);
                    unsigned ma = __ballot_sync(0xffffffffu, ina);
                    unsigned mb = __ballot_sync(0xffffffffu, inb);
                    if (ma | mb) {                                          // warp-uniform
#pragma unroll
                        for (int r = NR - 1; r >= 0; --r) {
                            if (r < NR - 1) {                               // next smaller ball: a subset of the previous hits
                                ina = !(d.x > q.r2[r]); inb = !(d.y > q.r2[r]);
                                ma = __ballot_sync(0xffffffffu, ina);
                                mb = __ballot_sync(0xffffffffu, inb);
                                if (!(ma | mb)) break;
                            }
                            if (cn[c][r] < q.K[r]) {
                                int *stg = stage + c * q.ktot + q.koff[r];
                                const int pa = cn[c][r] + __popc(ma & lt_mask) + __popc(mb & lt_mask);
                                if (ina && pa < q.K[r]) stg[pa] = na;
                                if (inb && pa + (int)ina < q.K[r]) stg[pa + (int)ina] = na + 1;
                                cn[c][r] += __popc(ma) + __popc(mb);
                            }
                        }
                    }
                }
            }
        };
        if (C > 4 && grp_short) scan(std::true_type{});
        else scan(std::false_type{});
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            const int sidx = s_begin + ci + c;
            if (!(ci + c < cpw && sidx < S)) continue;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int K = q.K[r];
                const int *stg = stage + c * q.ktot + q.koff[r];
                const int cc = cn[c][r] > K ? K : cn[c][r];
                const int frst = (cc && stg[0] < N) ? stg[0] : N;            // padding slots (NaN centre) are never reported
                if (frst == N && lane == 0 && empty_flag) atomicOr(empty_flag, 1);
                const int64_t o = ((int64_t)b * S + sidx) * K;
                for (int j = lane; j < K; j += 32) {
                    const int v = (j < cc && stg[j] < N) ? stg[j] : frst;
                    if (q.idx64[r]) q.idx64[r][o + j] = v;
                    if (q.idx32[r]) q.idx32[r][o + j] = v;
                }
            }
        }
        __syncwarp();
    }
}

// Multi-radius scan with deferred hit bookkeeping (C <= 4): ball_query_masks_kernel's scheme with one parked mask word per
// (centre, radius).  A trip evaluates the distances once, ballots against the largest radius, and only when some lane hit does
// it ballot the next smaller one (the balls are nested); the NR lists of the four centres in flight are resolved per 1024-point
// window.  The lists fill at very different rates (the MSG layers pair a 16-entry list at r = 0.1 with a 128-entry list at
// r = 0.4), so the scan practically always covers the whole cloud: the all-lists-full test runs in the middle and at the end of
// a window only.
template <int C, int BQ_WARPS, int NR>
__global__ void __launch_bounds__(BQ_WARPS * 32, 3)
ball_query_multi_masks_kernel(const float *__restrict__ centres, int64_t csb, int64_t csn, int64_t csc,
                              const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                              int S, int N, int NCH /* multiple of 128 */, int cpw, const BqMulti q, int32_t *__restrict__ empty_flag)
{
    static_assert(C <= 4, "one float4 of points per channel per lane");
    constexpr int BQ_THREADS = BQ_WARPS * 32;
    constexpr int CPB = 4;
    extern __shared__ float smem[];
    float *sp = smem;                       // [C][NCH], permuted inside every block of 128 points (see ball_query_masks_kernel)
    float *sp2 = smem + (size_t)C * NCH;    // [NCH]
    int *stage = (int *)(sp2 + NCH) + (threadIdx.x >> 5) * (CPB * q.ktot);       // [CPB][ktot] per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int s_begin = (blockIdx.x * BQ_WARPS + warp) * cpw;
    const float *P = points + (int64_t)b * sb;
    const bool pcontig = (sc == 1);

    for (int n = tid; n < N; n += BQ_THREADS) {
        float v[C];
        const int pos = (n & ~127) | ((n & 31) << 2) | ((n >> 5) & 3);
#pragma unroll
        for (int k = 0; k < C; ++k) {
            v[k] = __ldg(P + (int64_t)n * sn + (int64_t)k * sc);
            sp[k * NCH + pos] = v[k];
        }
        sp2[pos] = sqnorm<C>(v, pcontig, n >= (N & ~31));
    }
    for (int n = N + tid; n < NCH; n += BQ_THREADS) {
        const int pos = (n & ~127) | ((n & 31) << 2) | ((n >> 5) & 3);
#pragma unroll
        for (int k = 0; k < C; ++k) sp[k * NCH + pos] = 0.f;
        sp2[pos] = INFINITY;
    }
    __syncthreads();

    const f32x2 m2v = pk2(-2.0f, -2.0f);
    const bool odd = (lane & 1) != 0, upper = (lane & 2) != 0;
    for (int ci = 0; ci < cpw; ci += CPB) {
        if (s_begin + ci >= S) break;                                   // warp-uniform
        f32x2 sv[CPB][C], s2v[CPB];
        int cn[CPB][NR];                                                // hits resolved so far (warp-uniform)
        unsigned hm[CPB][NR];                                           // this lane's word of the current window, per radius
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            const int sidx = s_begin + ci + c;
            const bool act = ci + c < cpw && sidx < S;
            float sc_[C];
#pragma unroll
            for (int k = 0; k < C; ++k) sc_[k] = act ? __ldg(centres + (int64_t)b * csb + (int64_t)sidx * csn + (int64_t)k * csc) : 0.f;
            const float q2 = sqnorm<C>(sc_, csc == 1, sidx >= (S & ~31));
#pragma unroll
            for (int k = 0; k < C; ++k) sv[c][k] = pk2(sc_[k], sc_[k]);
            s2v[c] = pk2(q2, q2);
#pragma unroll
            for (int r = 0; r < NR; ++r) cn[c][r] = act ? 0 : q.K[r];
        }
        auto all_full = [&](bool with_window) -> bool {
            bool full = true;
#pragma unroll
            for (int c = 0; c < CPB; ++c)
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    int have = cn[c][r];
                    if (with_window) have += (int)__reduce_add_sync(0xffffffffu, (unsigned)__popc(hm[c][r]));
                    full = full && have >= q.K[r];
                }
            return full;
        };
        for (int w0 = 0; w0 < NCH; w0 += 1024) {                        // windows of 1024 points = 32 mask words
            const int trips = min(8, (NCH - w0) >> 7);
#pragma unroll
            for (int c = 0; c < CPB; ++c)
#pragma unroll
                for (int r = 0; r < NR; ++r) hm[c][r] = 0u;
            const float *prow = sp + w0 + 4 * lane;
            for (int t = 0; t < trips; ++t, prow += 128) {
                float4 pv[C];
#pragma unroll
                for (int k = 0; k < C; ++k) pv[k] = *reinterpret_cast<const float4 *>(prow + k * NCH);
                const float4 p2 = *reinterpret_cast<const float4 *>(prow + C * NCH);
                const f32x2 p2a = pk2(p2.x, p2.y), p2b = pk2(p2.z, p2.w);
                const bool mine = (lane >> 2) == t;
#pragma unroll
                for (int c = 0; c < CPB; ++c) {
                    f32x2 da = mul2(sv[c][0], pk2(pv[0].x, pv[0].y));
                    f32x2 db = mul2(sv[c][0], pk2(pv[0].z, pv[0].w));
#pragma unroll
                    for (int k = 1; k < C; ++k) {
                        da = fma2(sv[c][k], pk2(pv[k].x, pv[k].y), da);
                        db = fma2(sv[c][k], pk2(pv[k].z, pv[k].w), db);
                    }
                    const float2 d01 = unpk2(add2(fma2(m2v, da, s2v[c]), p2a));
                    const float2 d23 = unpk2(add2(fma2(m2v, db, s2v[c]), p2b));
#pragma unroll
                    for (int r = NR - 1; r >= 0; --r) {                 // largest ball first: the smaller ones are subsets
                        const float rr = q.r2[r];
                        const unsigned m0 = __ballot_sync(0xffffffffu, !(d01.x > rr));
                        const unsigned m1 = __ballot_sync(0xffffffffu, !(d01.y > rr));
                        const unsigned m2 = __ballot_sync(0xffffffffu, !(d23.x > rr));
                        const unsigned m3 = __ballot_sync(0xffffffffu, !(d23.y > rr));
                        const unsigned pick = odd ? (upper ? m3 : m1) : (upper ? m2 : m0);
                        hm[c][r] = mine ? pick : hm[c][r];
                        if (!(m0 | m1 | m2 | m3)) break;                // warp-uniform: nothing in this ball, nothing in the smaller ones
                    }
                }
                if (t == 3 && trips > 4 && all_full(true)) break;
            }
            // ---- resolve the window: lane l owns points w0 + 32 l .. + 31 ------------------------------------------------------
            const int nb = w0 + 32 * lane;
            const unsigned valid = nb + 32 <= N ? 0xffffffffu : (nb >= N ? 0u : (1u << (N - nb)) - 1u);
#pragma unroll
            for (int c = 0; c < CPB; ++c)
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const int K = q.K[r];
                    unsigned m = hm[c][r] & valid;
                    const int pc = __popc(m);
                    const int total = (int)__reduce_add_sync(0xffffffffu, (unsigned)pc);
                    if (cn[c][r] < K && total) {                        // warp-uniform
                        int inc = pc;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                        int pos = cn[c][r] + inc - pc;
                        int *stg = stage + c * q.ktot + q.koff[r];
                        while (m && pos < K) { stg[pos++] = nb + __ffs(m) - 1; m &= m - 1; }
                    }
                    cn[c][r] = min(K, cn[c][r] + total);
                }
            if (all_full(false)) break;
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < CPB; ++c) {
            const int sidx = s_begin + ci + c;
            if (!(ci + c < cpw && sidx < S)) continue;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int K = q.K[r];
                const int *stg = stage + c * q.ktot + q.koff[r];
                const int cc = cn[c][r];
                const int frst = cc ? stg[0] : N;
                if (cc == 0 && lane == 0 && empty_flag) atomicOr(empty_flag, 1);
                const int64_t o = ((int64_t)b * S + sidx) * K;
                for (int j = lane; j < K; j += 32) {
                    const int v = (j < cc) ? stg[j] : frst;
                    if (q.idx64[r]) q.idx64[r][o + j] = v;
                    if (q.idx32[r]) q.idx32[r][o + j] = v;
                }
            }
        }
        __syncwarp();
    }
}

// Run-time-C fallback (C <= PN2_MAX_COORD): one thread per centre, points streamed from global memory.
__global__ void ball_query_generic_kernel(const float *__restrict__ centres, int64_t csb, int64_t csn, int64_t csc,
                                          const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                                          int B, int S, int N, int C, float r2, int K,
                                          int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, int32_t *__restrict__ empty_flag)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)B * S) return;
    const int b = (int)(g / S);
    const float *P = points + (int64_t)b * sb;
    const float *cp = centres + (int64_t)b * csb + (g % S) * csn;
    float s[PN2_MAX_COORD], t[PN2_MAX_COORD];
    for (int k = 0; k < C; ++k) { s[k] = cp[k * csc]; t[k] = __fmul_rn(s[k], s[k]); }
    const float s2 = sum_order_rt(t, C, csc == 1, (int)(g % S) >= (S & ~31));
    const bool pcontig = (sc == 1);
    int cnt = 0, first = N;
    for (int n = 0; n < N && cnt < K; ++n) {
        float dot = 0.f;
        for (int k = 0; k < C; ++k) {
            const float x = __ldg(P + (int64_t)n * sn + (int64_t)k * sc);
            t[k] = __fmul_rn(x, x);
            dot = (k == 0) ? __fmul_rn(s[0], x) : __fmaf_rn(s[k], x, dot);
        }
        const float p2 = sum_order_rt(t, C, pcontig, n >= (N & ~31));
        float d = __fmul_rn(-2.0f, dot);
        d = __fadd_rn(d, s2);
        d = __fadd_rn(d, p2);
        if (!(d > r2)) {
            if (cnt == 0) first = n;
            if (idx64) idx64[g * K + cnt] = n;
            if (idx32) idx32[g * K + cnt] = n;
            ++cnt;
        }
    }
    if (cnt == 0 && empty_flag) atomicOr(empty_flag, 1);
    for (int j = cnt; j < K; ++j) {
        if (idx64) idx64[g * K + j] = first;
        if (idx32) idx32[g * K + j] = first;
    }
}

// ---------------------------------------------------------------------------------------------------
// Ball query with an exact spatial cull (small radii): the points of the cloud are bucketed into a uniform grid over their
// bounding box whose cells are at least R_c = 1.01 * sqrt(r^2 + e) wide, e an upper bound of the rounding error of the
// reference's expanded-form distance (so a point the fp32 formula puts inside the ball is geometrically within R_c of the
// centre and therefore in one of the 27 cells around the centre's cell; the 1 % slack absorbs the rounding of the cell index
// itself).  A warp evaluates one centre against the points of those cells only -- with the SAME distance arithmetic as the
// scan kernel -- and sets a bit per hit in a per-warp N-bit mask; the first K set bits in index order are the reference's
// "first K in ascending index order" (:86-89).  The cull only decides WHICH pairs are evaluated, never the result of a pair:
// bit-identical rows.  Clouds with non-finite coordinates, and centres far outside the cloud, take every point as candidate.
constexpr int BQC_WARPS = 16;
constexpr int BQC_THREADS = BQC_WARPS * 32;
constexpr int BQC_GMAX = 16;
constexpr int BQC_CELLS = BQC_GMAX * BQC_GMAX * BQC_GMAX;

struct BqGrid {
    float lo[3], inv_w[3];
    int G[3];
    int full;            // every point is a candidate (non-finite cloud)
    float m2cap;         // 4 * max |p|^2: centres beyond it scan everything
};

template <int C>
__global__ void __launch_bounds__(BQC_THREADS)
ball_query_cell_kernel(const float *__restrict__ centres, int64_t csb, int64_t csn, int64_t csc,
                       const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                       int S, int N, int cpc /*centres per CTA*/, float r2, int K,
                       int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, int32_t *__restrict__ empty_flag)
{
    extern __shared__ float smem[];
    const int NW = (N + 31) >> 5;                       // words of a hit mask
    float *sp = smem;                                   // [C][N]
    float *sp2 = sp + (size_t)C * N;                    // [N]
    int *order = reinterpret_cast<int *>(sp2 + N);      // [N] point indices sorted by cell
    int *cstart = order + N;                            // [BQC_CELLS + 1]
    unsigned *words = reinterpret_cast<unsigned *>(cstart + BQC_CELLS + 1);    // [BQC_WARPS][NW]
    int *stage = reinterpret_cast<int *>(words + (size_t)BQC_WARPS * NW);       // [BQC_WARPS][K]
    int *rs = stage + (size_t)BQC_WARPS * K;                                     // [BQC_WARPS][20]: 9 x (start, exclusive prefix)
    __shared__ float red[BQC_WARPS][8];
    __shared__ BqGrid grid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float *P = points + (int64_t)b * sb;
    const bool pcontig = (sc == 1);

    // ---- stage the cloud (same arithmetic as ball_query_kernel), bounding box, largest norm, finiteness ---------------------
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, m2 = 0.f;
    bool fin = true;
    for (int n = tid; n < N; n += BQC_THREADS) {
        float v[C];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            v[k] = __ldg(P + (int64_t)n * sn + (int64_t)k * sc);
            sp[k * N + n] = v[k];
            fin = fin && fabsf(v[k]) <= 3.0e38f;
        }
        const float q = sqnorm<C>(v, pcontig, n >= (N & ~31));
        sp2[n] = q;
        m2 = fmaxf(m2, q);
#pragma unroll
        for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], v[k]); mx[k] = fmaxf(mx[k], v[k]); }
    }
    for (int i = tid; i < BQC_WARPS * NW; i += BQC_THREADS) words[i] = 0u;
    for (int i = tid; i <= BQC_CELLS; i += BQC_THREADS) cstart[i] = 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    const bool all_fin = __syncthreads_and(fin ? 1 : 0) != 0;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { red[warp][k] = mn[k]; red[warp][3 + k] = mx[k]; }
        red[warp][6] = m2;
    }
    __syncthreads();
    if (tid == 0) {
        float lo[3], hi[3], M2 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        for (int w = 0; w < BQC_WARPS; ++w) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], red[w][k]); hi[k] = fmaxf(hi[k], red[w][3 + k]); }
            M2 = fmaxf(M2, red[w][6]);
        }
        const bool ok = all_fin && M2 <= 1.0e30f && r2 >= 0.f && r2 <= 1.0e30f;
        grid.full = ok ? 0 : 1;
        grid.m2cap = 4.0f * M2;
        // |d_fp32 - |s - p|^2| <= u (4C + 8) max(|s|^2, |p|^2) with u = 2^-24; centres are admitted up to |s|^2 = 4 max|p|^2
        // and the bound is doubled once more
        const float err = 1.1920929e-7f * (4 * C + 8) * 4.0f * M2;
        const float h = 1.01f * sqrtf(r2 + err) + 1.0e-30f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ext = ok ? hi[k] - lo[k] : 0.f;
            float g = floorf(fminf(ext / h, (float)BQC_GMAX));
            if (!(g >= 1.f)) g = 1.f;
            grid.G[k] = (int)g;
            grid.lo[k] = ok ? lo[k] : 0.f;
            grid.inv_w[k] = (ok && ext > 0.f) ? g / ext : 0.f;
        }
    }
    __syncthreads();
    const BqGrid gr = grid;
    const int G0 = gr.G[0], G1 = gr.G[1], G2 = gr.G[2];
    auto cell_coord = [&](float x, int k, int Gk) -> int {       // unclamped cell coordinate, limited to [-2, Gk + 1]
        float c = floorf((x - gr.lo[k]) * gr.inv_w[k]);
        c = fminf(fmaxf(c, -2.f), (float)(Gk + 1));
        return (int)c;
    };
    auto point_cell = [&](int n) -> int {
        const int cx = min(G0 - 1, max(0, cell_coord(sp[n], 0, G0)));
        const int cy = min(G1 - 1, max(0, cell_coord(sp[N + n], 1, G1)));
        const int cz = min(G2 - 1, max(0, cell_coord(sp[2 * N + n], 2, G2)));
        return (cz * G1 + cy) * G0 + cx;
    };
    // ---- counting sort of the point indices by cell ---------------------------------------------------------------------------
    const int ncell = G0 * G1 * G2;
    if (!gr.full) {
        for (int n = tid; n < N; n += BQC_THREADS) atomicAdd(&cstart[point_cell(n)], 1);
        __syncthreads();
        // exclusive scan of the counts (ncell <= 4096 = 512 threads x 8)
        int loc[8], sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = tid * 8 + i; loc[i] = c < ncell ? cstart[c] : 0; sum += loc[i]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __shared__ int wsum[BQC_WARPS];
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        int run = base + inc - sum;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int c = tid * 8 + i; if (c < ncell) cstart[c] = run; run += loc[i]; }
        __syncthreads();
        // scatter: the running cursor of cell c ends at the start of cell c + 1
        for (int n = tid; n < N; n += BQC_THREADS) order[atomicAdd(&cstart[point_cell(n)], 1)] = n;
        __syncthreads();
    }
    // start of cell c = (c ? cstart[c - 1] : 0); end of cell c = cstart[c]
    auto cell_begin = [&](int c) -> int { return c ? cstart[c - 1] : 0; };

    // ---- centres: one warp per centre -----------------------------------------------------------------------------------------
    unsigned *wm = words + (size_t)warp * NW;
    int *stg = stage + (size_t)warp * K;
    int *wr = rs + warp * 20;
    const int s_begin = blockIdx.x * cpc;
    const int s_end = min(S, s_begin + cpc);
    const int WPL = (NW + 31) >> 5;                     // mask words per lane at emission
    for (int sidx = s_begin + warp; sidx < s_end; sidx += BQC_WARPS) {
        float sc_[C];
#pragma unroll
        for (int k = 0; k < C; ++k) sc_[k] = __ldg(centres + (int64_t)b * csb + (int64_t)sidx * csn + (int64_t)k * csc);
        const float s2 = sqnorm<C>(sc_, csc == 1, sidx >= (S & ~31));
        const bool full = gr.full || !(s2 <= gr.m2cap);            // NaN / far centres: every point is a candidate
        int T = N;
        if (!full) {
            const int cx = cell_coord(sc_[0], 0, G0), cy = cell_coord(sc_[1], 1, G1), cz = cell_coord(sc_[2], 2, G2);
            int start = 0, len = 0;
            if (lane < 9) {
                const int y = cy + (lane % 3) - 1, z = cz + (lane / 3) - 1;
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, G0 - 1);
                if (y >= 0 && y < G1 && z >= 0 && z < G2 && x0 <= x1) {
                    const int c0 = (z * G1 + y) * G0 + x0, c1 = (z * G1 + y) * G0 + x1;
                    start = cell_begin(c0);
                    len = cstart[c1] - start;
                }
            }
            int inc = len;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane < 9) { wr[2 * lane] = start; wr[2 * lane + 1] = inc - len; }
            T = __shfl_sync(0xffffffffu, inc, 8);
            __syncwarp();
        }
        for (int i0 = 0; i0 < T; i0 += 32) {
            const int i = i0 + lane;
            if (i < T) {
                int n = i;
                if (!full) {
                    int j = 0;
#pragma unroll
                    for (int t = 1; t < 9; ++t) j += (i >= wr[2 * t + 1]) ? 1 : 0;      // exclusive prefixes are non-decreasing
                    n = order[wr[2 * j] + (i - wr[2 * j + 1])];
                }
                float dot = __fmul_rn(sc_[0], sp[n]);
#pragma unroll
                for (int k = 1; k < C; ++k) dot = __fmaf_rn(sc_[k], sp[k * N + n], dot);
                float d = __fmaf_rn(-2.0f, dot, s2);        // fl(-2 dot + |s|^2): -2 dot is exact, one rounding either way
                d = __fadd_rn(d, sp2[n]);
                if (!(d > r2)) atomicOr(&wm[n >> 5], 1u << (n & 31));
            }
        }
        __syncwarp();
        // ---- the first K set bits in index order ------------------------------------------------------------------------------
        int cnt = 0;
        for (int w = 0; w < WPL; ++w) { const int wi = lane * WPL + w; if (wi < NW) cnt += __popc(wm[wi]); }
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        int pos = inc - cnt;
        for (int w = 0; w < WPL; ++w) {
            const int wi = lane * WPL + w;
            if (wi < NW) {
                unsigned m = wm[wi];
                wm[wi] = 0u;
                while (m && pos < K) { const int bit = __ffs(m) - 1; stg[pos++] = (wi << 5) + bit; m &= m - 1; }
            }
        }
        __syncwarp();
        const int cc = total > K ? K : total;
        const int frst = cc ? stg[0] : N;
        if (cc == 0 && lane == 0 && empty_flag) atomicOr(empty_flag, 1);
        const int64_t o = ((int64_t)b * S + sidx) * K;
        for (int j = lane; j < K; j += 32) {
            const int v = (j < cc) ? stg[j] : frst;
            if (idx64) idx64[o + j] = v;
            if (idx32) idx32[o + j] = v;
        }
        __syncwarp();
    }
}

static size_t bq_cell_smem(int C, int N, int K)
{
    const int NW = (N + 31) >> 5;
    return sizeof(float) * (size_t)(C + 1) * N + sizeof(int) * ((size_t)N + BQC_CELLS + 1 + (size_t)BQC_WARPS * NW + (size_t)BQC_WARPS * K + BQC_WARPS * 20);
}

// returns -100 when the culled kernel does not apply (the caller scans)
template <int C>
static int launch_bq_cells(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                           const float *points, int64_t sb, int64_t sn, int64_t sc,
                           int B, int S, int N, float r2, int K,
                           int64_t *idx64, int32_t *idx32, int32_t *empty_flag, cudaStream_t st)
{
    // PN2_BQ_CELLS=1 selects this kernel (where it fits); OFF by default.  Measured on B200 (profiles/r02_bq_cells.txt): the
    // reference's clouds are normalised to the unit sphere, where a radius of 0.2 leaves a 5^3 grid and the 27 cells around a
    // centre still hold ~22 % of the cloud; one centre per warp on those candidates costs as much as the scan kernel, which
    // tests four centres against 64 points per trip: 665 vs 666 us (B = 2048, N = 1024), 212 vs 180 us (B = 512), 62 vs 70 us
    // (B = 16, N = 4096), 234 vs 131 us (configs[2], N = 8192: one CTA per SM, grid rebuilt by every tile of centres).
    static const int env = getenv("PN2_BQ_CELLS") ? atoi(getenv("PN2_BQ_CELLS")) : 0;
    if (env <= 0) return -100;
    const size_t smem = bq_cell_smem(C, N, K);
    if (smem > 200 * 1024 || N < 64) return -100;
    auto kern = ball_query_cell_kernel<C>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("ball_query(cells): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    // one CTA per cloud when there are enough clouds; otherwise tiles of >= 64 centres, each rebuilding the grid (O(N))
    int tiles = 1;
    if (B < 296) { tiles = (296 + B - 1) / B; const int tmax = (S + 63) / 64; if (tiles > tmax) tiles = tmax; if (tiles < 1) tiles = 1; }
    const int cpc = ((S + tiles - 1) / tiles + BQC_WARPS - 1) / BQC_WARPS * BQC_WARPS;
    dim3 g((S + cpc - 1) / cpc, B);
    kern<<<g, BQC_THREADS, smem, st>>>(centres, csb, csn, csc, points, sb, sn, sc, S, N, cpc, r2, K, idx64, idx32, empty_flag);
    return launch_status("pn2_ball_query_f32(cells)");
}

template <int C, int BQ_WARPS>
static int launch_bq_w(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                       const float *points, int64_t sb, int64_t sn, int64_t sc,
                       int B, int S, int N, float r2, int K,
                       int64_t *idx64, int32_t *idx32, int32_t *empty_flag, cudaStream_t st)
{
    const size_t stage_bytes = sizeof(int) * (size_t)BQ_WARPS * BQ_CPB<C>::value * K;
    const size_t budget = 200 * 1024 - stage_bytes;
    // deferred hit bookkeeping (ball_query_masks_kernel) wherever the cloud is staged whole and C <= 4; PN2_BQ_V1=1 keeps the
    // per-trip bookkeeping of ball_query_kernel (the round-1 kernel) for re-measurement
    static const bool v1_env = getenv("PN2_BQ_V1") && atoi(getenv("PN2_BQ_V1")) == 1;
    bool masks = false;
    if constexpr (C <= 4) masks = !v1_env && (size_t)(C + 1) * ((N + 127) & ~127) * sizeof(float) <= budget;
    int NCH = masks ? ((N + 127) & ~127) : ((N + 63) & ~63), cpw;          // rows padded to whole trips (128 / 64 points): the inner loops are unguarded
    if ((size_t)(C + 1) * NCH * sizeof(float) > budget) {
        NCH = (int)(budget / ((C + 1) * sizeof(float))) & ~31;
        cpw = 1;  // multi-chunk path keeps one centre per warp in registers
    } else {
        // enough CTAs for ~4 waves of 148 SMs, but amortise the staging of the cloud over >= 8 centres/warp
        cpw = 16;
        while (cpw > 1 && (int64_t)B * ((S + BQ_WARPS * cpw - 1) / (BQ_WARPS * cpw)) < 592) cpw >>= 1;
        // at least one FULL pass per warp: a pass costs the same with idle centre slots (few clouds used to run passes of one or
        // two centres in four slots on up to four times the CTAs; B = 16, N = 4096, S = 1024: 88 -> 57 us at 8 per warp)
        if (cpw < BQ_CPB<C>::value) cpw = BQ_CPB<C>::value;
        static const int cpw_env = getenv("PN2_BQ_CPW") ? atoi(getenv("PN2_BQ_CPW")) : 0;
        if (cpw_env > 0) cpw = cpw_env;
    }
    const size_t smem = (size_t)(C + 1) * NCH * sizeof(float) + stage_bytes;
    dim3 grid((S + BQ_WARPS * cpw - 1) / (BQ_WARPS * cpw), B);
    if constexpr (C <= 4) {
        if (masks) {
            auto kern = ball_query_masks_kernel<C, BQ_WARPS>;
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { set_error("ball_query(masks): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
            }
            kern<<<grid, BQ_WARPS * 32, smem, st>>>(centres, csb, csn, csc, points, sb, sn, sc, S, N, NCH, cpw, r2, K,
                                                    idx64, idx32, empty_flag);
            return launch_status("pn2_ball_query_f32(masks)");
        }
    }
    auto kern = ball_query_kernel<C, BQ_WARPS>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("ball_query: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    kern<<<grid, BQ_WARPS * 32, smem, st>>>(centres, csb, csn, csc, points, sb, sn, sc, S, N, NCH, cpw, r2, K,
                                            idx64, idx32, empty_flag);
    return launch_status("pn2_ball_query_f32");
}

template <int C>
static int launch_bq_c(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                       const float *points, int64_t sb, int64_t sn, int64_t sc,
                       int B, int S, int N, float r2, int K,
                       int64_t *idx64, int32_t *idx32, int32_t *empty_flag, cudaStream_t st)
{
    // a staged cloud of more than ~100 KB leaves room for one CTA per SM: give it 32 warps instead of 8 (C <= 4 only: the
    // C = 10 kernel needs more than the 64 registers a 1024-thread CTA can have)
    {
        const int rc = launch_bq_cells<C>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, r2, K, idx64, idx32, empty_flag, st);
        if (rc != -100) return rc;
    }
    const size_t cloud = (size_t)(C + 1) * ((N + 63) & ~63) * sizeof(float);
    if (C <= 4 && cloud > 100 * 1024 && cloud + sizeof(int) * (size_t)BQ_WARPS_BIG * BQ_CPB<C>::value * K <= 200 * 1024)
        return launch_bq_w<C, BQ_WARPS_BIG>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, r2, K, idx64, idx32, empty_flag, st);
    return launch_bq_w<C, BQ_WARPS_SMALL>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, r2, K, idx64, idx32, empty_flag, st);
}

int launch_ball_query(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                      const float *points, int64_t sb, int64_t sn, int64_t sc,
                      int B, int S, int N, int C, float r2, int K,
                      int64_t *idx64, int32_t *idx32, int32_t *empty_flag, cudaStream_t st)
{
    if (B == 0 || S == 0) return 0;
    PN2_REQUIRE(B <= 65535, "pn2_ball_query_f32: B=%d exceeds gridDim.y", B);
    // the scan kernels stage K indices per centre in flight (8 warps x 4 or 2 centres): lists beyond 1024 entries (the reference's
    // longest is 128) take the one-thread-per-centre kernel, which needs no staging
    const bool staged = (size_t)K * sizeof(int) * BQ_WARPS_SMALL * (C <= 4 ? 4 : 2) <= 128 * 1024;
    if (C == 3 && staged) return launch_bq_c<3>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, r2, K, idx64, idx32, empty_flag, st);
    if (C == 10 && staged) return launch_bq_c<10>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, r2, K, idx64, idx32, empty_flag, st);
    const int64_t G = (int64_t)B * S;
    ball_query_generic_kernel<<<(unsigned)((G + 127) / 128), 128, 0, st>>>(centres, csb, csn, csc, points, sb, sn, sc,
                                                                            B, S, N, C, r2, K, idx64, idx32, empty_flag);
    return launch_status("pn2_ball_query_f32(generic)");
}

template <int C, int NR>
static int launch_bq_multi(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                           const float *points, int64_t sb, int64_t sn, int64_t sc,
                           int B, int S, int N, int NCH, const BqMulti &q, int32_t *empty_flag, cudaStream_t st)
{
    constexpr int W = BQ_WARPS_SMALL;
    int cpw = 16;
    while (cpw > BQ_CPB<C>::value && (int64_t)B * ((S + W * cpw - 1) / (W * cpw)) < 592) cpw >>= 1;      // never less than one full pass
    if constexpr (C <= 4) {
        // deferred hit bookkeeping (ball_query_multi_masks_kernel) unless PN2_BQ_V1=1; rows padded to whole 128-point trips
        static const bool v1_env = getenv("PN2_BQ_V1") && atoi(getenv("PN2_BQ_V1")) == 1;
        const int NCH2 = (N + 127) & ~127;
        const size_t smem2 = (size_t)(C + 1) * NCH2 * sizeof(float) + sizeof(int) * (size_t)W * 4 * q.ktot;
        if (!v1_env && smem2 <= 200 * 1024) {
            auto kern2 = ball_query_multi_masks_kernel<C, W, NR>;
            if (smem2 > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                if (e != cudaSuccess) { set_error("ball_query_multi(masks): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
            }
            dim3 grid2((S + W * cpw - 1) / (W * cpw), B);
            kern2<<<grid2, W * 32, smem2, st>>>(centres, csb, csn, csc, points, sb, sn, sc, S, N, NCH2, cpw, q, empty_flag);
            return launch_status("pn2_ball_query_multi_f32(masks)");
        }
    }
    const size_t smem = (size_t)(C + 1) * NCH * sizeof(float) + sizeof(int) * (size_t)W * BQ_CPB<C>::value * q.ktot;
    auto kern = ball_query_multi_kernel<C, W, NR>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("ball_query_multi: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    dim3 grid((S + W * cpw - 1) / (W * cpw), B);
    kern<<<grid, W * 32, smem, st>>>(centres, csb, csn, csc, points, sb, sn, sc, S, N, NCH, cpw, q, empty_flag);
    return launch_status("pn2_ball_query_multi_f32");
}

// Host arrays r2[NR] (ascending), K[NR], idx64[NR] / idx32[NR] (device pointers, entries or the arrays may be NULL).
int launch_ball_query_multi(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                            const float *points, int64_t sb, int64_t sn, int64_t sc,
                            int B, int S, int N, int C, int NR, const float *r2, const int *K,
                            int64_t *const *idx64, int32_t *const *idx32, int32_t *empty_flag, cudaStream_t st)
{
    if (B == 0 || S == 0) return 0;
    PN2_REQUIRE(B <= 65535, "pn2_ball_query_multi_f32: B=%d exceeds gridDim.y", B);
    BqMulti q;
    q.ktot = 0;
    bool ascending = true;
    for (int r = 0; r < NR; ++r) {
        q.r2[r] = r2[r]; q.K[r] = K[r]; q.koff[r] = q.ktot; q.ktot += K[r];
        q.idx64[r] = idx64 ? idx64[r] : nullptr;
        q.idx32[r] = idx32 ? idx32[r] : nullptr;
        if (r && !(r2[r - 1] <= r2[r])) ascending = false;
    }
    for (int r = NR; r < 3; ++r) { q.r2[r] = 0.f; q.K[r] = 0; q.koff[r] = 0; q.idx64[r] = nullptr; q.idx32[r] = nullptr; }
    const int NCH = (N + 63) & ~63;
    const size_t cloud = (size_t)(C + 1) * NCH * sizeof(float);
    const size_t stage = sizeof(int) * (size_t)BQ_WARPS_SMALL * (C <= 4 ? 4 : 2) * q.ktot;
    const bool fused = ascending && (NR == 2 || NR == 3) && (C == 3 || C == 10) && cloud <= 100 * 1024 && cloud + stage <= 200 * 1024;
    if (!fused) {         // one query per radius (unsorted radii, other widths, clouds that need the chunked / 32-warp kernel)
        for (int r = 0; r < NR; ++r) {
            const int rc = launch_ball_query(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, C, r2[r], K[r],
                                             q.idx64[r], q.idx32[r], empty_flag, st);
            if (rc) return rc;
        }
        return 0;
    }
    if (C == 3) return NR == 2 ? launch_bq_multi<3, 2>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, NCH, q, empty_flag, st)
                               : launch_bq_multi<3, 3>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, NCH, q, empty_flag, st);
    return NR == 2 ? launch_bq_multi<10, 2>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, NCH, q, empty_flag, st)
                   : launch_bq_multi<10, 3>(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, NCH, q, empty_flag, st);
}

// ---------------------------------------------------------------------------------------------------
// square_distance (pointnet2_utils.py:5-26): materialises [B,S,N]; only the stand-alone function uses it.
__global__ void square_distance_kernel(const float *__restrict__ src, int64_t ssb, int64_t ssn, int64_t ssc,
                                       const float *__restrict__ dst, int64_t dsb, int64_t dsn, int64_t dsc,
                                       int S, int N, int C, float *__restrict__ out)
{
    const int b = blockIdx.z, s = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float sv[PN2_MAX_COORD], t[PN2_MAX_COORD], u[PN2_MAX_COORD];
    const float *sp = src + (int64_t)b * ssb + (int64_t)s * ssn;
    const float *dp = dst + (int64_t)b * dsb + (int64_t)n * dsn;
    float dot = 0.f;
    for (int k = 0; k < C; ++k) {
        sv[k] = __ldg(sp + (int64_t)k * ssc);
        const float x = __ldg(dp + (int64_t)k * dsc);
        t[k] = __fmul_rn(sv[k], sv[k]);
        u[k] = __fmul_rn(x, x);
        dot = (k == 0) ? __fmul_rn(sv[0], x) : __fmaf_rn(sv[k], x, dot);
    }
    float d = __fmul_rn(-2.0f, dot);
    d = __fadd_rn(d, sum_order_rt(t, C, ssc == 1, s >= (S & ~31)));
    d = __fadd_rn(d, sum_order_rt(u, C, dsc == 1, n >= (N & ~31)));
    out[((int64_t)b * S + s) * N + n] = d;
}

int launch_square_distance(const float *src, int64_t ssb, int64_t ssn, int64_t ssc,
                           const float *dst, int64_t dsb, int64_t dsn, int64_t dsc,
                           int B, int S, int N, int C, float *out, cudaStream_t st)
{
    if (B == 0 || S == 0 || N == 0) return 0;
    PN2_REQUIRE(B <= 65535 && S <= 65535, "pn2_square_distance_f32: B=%d / S=%d exceed the grid limits", B, S);
    dim3 grid((N + 255) / 256, S, B);
    square_distance_kernel<<<grid, 256, 0, st>>>(src, ssb, ssn, ssc, dst, dsb, dsn, dsc, S, N, C, out);
    return launch_status("pn2_square_distance_f32");
}

}  // namespace pn2
