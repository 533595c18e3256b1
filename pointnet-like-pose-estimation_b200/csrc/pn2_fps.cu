// K1 -- farthest point sampling, one CTA per cloud (reference: model/pointnet2_utils.py:47-68).
//
// The whole cloud lives on chip for the duration of the S sequential iterations: every thread keeps its
// PPT points and their running min-distances in registers, a structure-of-arrays copy of the cloud in
// shared memory serves the (broadcast) read of the newly chosen centroid, and the per-iteration argmax is
// two REDUX instructions per warp + one 8-byte shared-memory slot per warp + ONE __syncthreads (slots are
// double-buffered by iteration parity).  No global-memory round trip inside the loop; the only global
// traffic per iteration is the fire-and-forget store of the chosen index / centroid row.
//
// Exactness contract: d = sum_order_k( fl( fl(p_k - c_k)^2 ) ) with separately rounded sub, mul
// and adds; dist = d if d < dist (NaN never replaces -> fminf); argmax with lowest index on ties.
#include <stdlib.h>
#include "pn2_common.cuh"

namespace pn2 {

__device__ __forceinline__ unsigned redux_max(unsigned v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned redux_min(unsigned v) { return __reduce_min_sync(0xffffffffu, v); }

template <int C, int PPT, bool CONTIG, int TMAX>
__global__ void __launch_bounds__(TMAX) fps_kernel(const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc, int N,
                           const int64_t *__restrict__ start, int S,
                           int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ new_points)
{
    extern __shared__ float s_pts[];            // [C][Npad]  (Npad = PPT * blockDim.x)
    __shared__ uint2 s_slot[2][32];             // per-warp (dist bits, index), double-buffered
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int Npad = PPT * T;
    const int n_body = N & ~31;                 // strided operand: points >= n_body reduce in the leftover order
    // packed (two points per instruction) unless some point needs the leftover order
    const bool packed = CONTIG || C <= 4 || n_body == N;
    const int b = blockIdx.x;
    const float *P = points + (int64_t)b * sb;

    static_assert(PPT % 2 == 0, "points are kept as register pairs");
    f32x2 pp[PPT / 2][C];                       // this thread's points (tid + j*T), two per 64-bit register
    float dist[PPT];
    // The pose heads feed xyz + a one-hot category vector as "coordinates" (rotation_ssg.py:27, provider.py:166-180): the
    // channels from 3 on are the same finite constants for every point of a cloud, so p_k - c_k = +0 exactly there and the
    // squared distance is the three-channel one -- adding +0 terms changes no fp32 sum, and every summation order of
    // DESIGN.md section 2 restricted to its non-zero terms is ((t0 + t1) + t2).  Detected per cloud, bit-identical.
    bool tail_same = C > 4;
#pragma unroll
    for (int j = 0; j < PPT; j += 2) {
        const int n0 = tid + j * T, n1 = n0 + T;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const float v0 = n0 < N ? __ldg(P + (int64_t)n0 * sn + (int64_t)k * sc) : 0.f;
            const float v1 = n1 < N ? __ldg(P + (int64_t)n1 * sn + (int64_t)k * sc) : 0.f;
            pp[j / 2][k] = pk2(v0, v1);
            s_pts[k * Npad + n0] = v0;
            s_pts[k * Npad + n1] = v1;
            if (C > 4 && k >= 3) {
                const float ref = __ldg(P + (int64_t)k * sc);          // point 0
                const bool fin = fabsf(ref) <= 3.0e38f;                 // finite: x - x = +0 (false for NaN / inf)
                tail_same = tail_same && fin && (n0 >= N || v0 == ref) && (n1 >= N || v1 == ref);
            }
        }
        dist[j] = n0 < N ? 1e10f : 0.f;         // padding can never win: dist 0 and index >= N
        dist[j + 1] = n1 < N ? 1e10f : 0.f;
    }
    int far = (int)min(max(start[b], (int64_t)0), (int64_t)N - 1);   // contract: start in [0,N); never index out of the cloud
    // shared-memory addresses of the argmax slots, computed once (the compiler otherwise rebuilds them every iteration)
    const uint32_t slot_wr = (uint32_t)__cvta_generic_to_shared(&s_slot[0][warp]);
    const uint32_t slot_rd = (uint32_t)__cvta_generic_to_shared(&s_slot[0][lane < nwarps ? lane : 0]);
    const bool slot_live = lane < nwarps;
    const bool tail_uniform = __syncthreads_and(tail_same ? 1 : 0) != 0 && C > 4;

    if (tail_uniform) {
        // three-channel loop (see above); sequential sum = every layout's order restricted to the non-zero terms
        constexpr int CE = C > 4 ? 3 : C;
        for (int i = 0; i < S; ++i) {
            f32x2 nc[CE];
#pragma unroll
            for (int k = 0; k < CE; ++k) { const float ck = s_pts[k * Npad + far]; nc[k] = pk2(-ck, -ck); }
            if (tid == 0) {
                if (idx64) idx64[(int64_t)b * S + i] = far;
                if (idx32) idx32[(int64_t)b * S + i] = far;
            }
            if (new_points && tid < C) new_points[((int64_t)b * S + i) * C + tid] = s_pts[tid * Npad + far];
            float best = -1.f;
            unsigned besti = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < PPT; j += 2) {
                f32x2 t[CE];
#pragma unroll
                for (int k = 0; k < CE; ++k) {
                    const f32x2 df = add2(pp[j / 2][k], nc[k]);
                    t[k] = mul2(df, df);
                }
                const float2 d = unpk2(sum_order2<CE, false>(t));
                dist[j] = fminf(dist[j], d.x);
                dist[j + 1] = fminf(dist[j + 1], d.y);
                if (dist[j] > best) { best = dist[j]; besti = (unsigned)(tid + j * T); }
                if (dist[j + 1] > best) { best = dist[j + 1]; besti = (unsigned)(tid + (j + 1) * T); }
            }
            const unsigned bits = __float_as_uint(best);
            const unsigned wmax = redux_max(bits);
            const unsigned widx = redux_min(bits == wmax ? besti : 0xffffffffu);
            const uint32_t par_off = (uint32_t)(i & 1) * (uint32_t)sizeof(s_slot[0]);
            if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot_wr + par_off), "r"(wmax), "r"(widx) : "memory");
            __syncthreads();
            uint2 sl;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(sl.x), "=r"(sl.y) : "r"(slot_rd + par_off) : "memory");
            if (!slot_live) sl = make_uint2(0u, 0xffffffffu);
            const unsigned gmax = redux_max(sl.x);
            far = (int)redux_min(sl.x == gmax ? sl.y : 0xffffffffu);
        }
        return;
    }

    for (int i = 0; i < S; ++i) {
        float c[C];
#pragma unroll
        for (int k = 0; k < C; ++k) c[k] = s_pts[k * Npad + far];
        if (tid == 0) {
            if (idx64) idx64[(int64_t)b * S + i] = far;
            if (idx32) idx32[(int64_t)b * S + i] = far;
        }
        if (new_points && tid < C) new_points[((int64_t)b * S + i) * C + tid] = s_pts[tid * Npad + far];

        float best = -1.f;
        unsigned besti = 0xffffffffu;
        if (packed) {
            // two of this thread's points per instruction (sub as add of the negated centroid: exact)
            f32x2 nc[C];
#pragma unroll
            for (int k = 0; k < C; ++k) nc[k] = pk2(-c[k], -c[k]);
#pragma unroll
            for (int j = 0; j < PPT; j += 2) {
                f32x2 t[C];
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    const f32x2 df = add2(pp[j / 2][k], nc[k]);
                    t[k] = mul2(df, df);
                }
                const float2 d = unpk2(sum_order2<C, CONTIG>(t));
                dist[j] = fminf(dist[j], d.x);          // d < dist ? d : dist, NaN never replaces
                dist[j + 1] = fminf(dist[j + 1], d.y);
                if (dist[j] > best) { best = dist[j]; besti = (unsigned)(tid + j * T); }
                if (dist[j + 1] > best) { best = dist[j + 1]; besti = (unsigned)(tid + (j + 1) * T); }
            }
        } else {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            float t[C];
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const float2 pv = unpk2(pp[j / 2][k]);
                const float df = __fsub_rn((j & 1) ? pv.y : pv.x, c[k]);
                t[k] = __fmul_rn(df, df);
            }
            const float d = sum_order<C, CONTIG>(t, tid + j * T >= n_body);
            dist[j] = fminf(dist[j], d);        // d < dist ? d : dist, NaN never replaces
            if (dist[j] > best) { best = dist[j]; besti = (unsigned)(tid + j * T); }
        }
        }
        // warp argmax: dist >= +0 so the bit pattern orders like the value; lowest index among equals
        const unsigned bits = __float_as_uint(best);
        const unsigned wmax = redux_max(bits);
        const unsigned widx = redux_min(bits == wmax ? besti : 0xffffffffu);
        const uint32_t par_off = (uint32_t)(i & 1) * (uint32_t)sizeof(s_slot[0]);
        if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot_wr + par_off), "r"(wmax), "r"(widx) : "memory");
        __syncthreads();
        uint2 sl;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(sl.x), "=r"(sl.y) : "r"(slot_rd + par_off) : "memory");
        if (!slot_live) sl = make_uint2(0u, 0xffffffffu);
        const unsigned gmax = redux_max(sl.x);
        far = (int)redux_min(sl.x == gmax ? sl.y : 0xffffffffu);
    }
}

// ------------------------------------------------------------------------------------------------------
// Cluster variant: ONE CLOUD PER THREAD-BLOCK CLUSTER of CS (4 or 8) CTAs, for few large clouds (BASELINE configs[2]: B = 16,
// N = 8192 -> 2048: sixteen single CTAs would leave 132 SMs idle and spend 1650 cycles per iteration on 8 points per
// thread).  Every CTA keeps the whole cloud in shared memory (centroid lookup) and 1/CS of the points + running distances
// in registers.  Per iteration: argmax inside each warp as above, then (DIRECT) lanes 0..CS-1 of EVERY WARP store the warp's
// (distance bits, index) into that warp's slot in the slot array of every CTA of the cluster through distributed shared
// memory (st.shared::cluster, double-buffered by iteration parity, tagged with the iteration number); every warp polls the
// CS x 8 slots of its own CTA (one or two per lane) until all carry the current tag and reduces them: same maximum, lowest
// index among equals, so the result is bit-identical to the one-CTA kernel.  There is no block or cluster barrier inside the
// loop.  (DIRECT = false is the first version: a block-level reduction, one slot per CTA; kept for PN2_FPS_CLUSTER=1.)
// ------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int C, int PPT, bool CONTIG, bool DIRECT, int CS, int CW>     // CS CTAs per cluster, CW warps per CTA
__global__ void __launch_bounds__(CW * 32) fps_cluster_kernel(const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc, int N,
                                                          const int64_t *__restrict__ start, int S,
                                                          int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ new_points)
{
    extern __shared__ float s_pts[];            // [C][Nall]: the WHOLE cloud (Nall = CS * PPT * blockDim.x >= N)
    __shared__ uint2 s_slot[2][32];             // per-warp (dist bits, index), double-buffered
    // DIRECT: one slot per WARP of the cluster (the CTA-level reduction, its barrier and two REDUX leave the dependent chain)
    constexpr int NSLOT = DIRECT ? CS * CW : CS;
    constexpr int SPL = NSLOT > 32 ? 2 : 1;      // DIRECT: slots polled per lane (fewer than 32 slots: lanes wrap around, duplicates do not change a maximum)
    static_assert(!DIRECT || NSLOT == 16 || NSLOT == 32 || NSLOT == 64, "one or two slots per lane");
    __shared__ __align__(16) uint2 c_slot[2][NSLOT];        // results of the cluster, written by the peers
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int b = blockIdx.x / CS;
    const int Nc = PPT * T, Nall = CS * Nc;
    const float *P = points + (int64_t)b * sb;
    static_assert(PPT % 2 == 0 && C <= 4, "two points per register pair; packed arithmetic needs C <= 4 on strided clouds");

    for (int n = tid; n < Nall; n += T) {
#pragma unroll
        for (int k = 0; k < C; ++k) s_pts[k * Nall + n] = n < N ? __ldg(P + (int64_t)n * sn + (int64_t)k * sc) : 0.f;
    }
    __syncthreads();
    const int base = (int)rank * Nc;            // this CTA's points: base + tid + j*T
    f32x2 pp[PPT / 2][C];
    float dist[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j += 2) {
        const int n0 = base + tid + j * T, n1 = n0 + T;
#pragma unroll
        for (int k = 0; k < C; ++k) pp[j / 2][k] = pk2(s_pts[k * Nall + n0], s_pts[k * Nall + n1]);
        dist[j] = n0 < N ? 1e10f : 0.f;
        dist[j + 1] = n1 < N ? 1e10f : 0.f;
    }
    int far = (int)min(max(start[b], (int64_t)0), (int64_t)N - 1);
    const uint32_t slot_wr = (uint32_t)__cvta_generic_to_shared(&s_slot[0][warp]);
    const uint32_t slot_rd = (uint32_t)__cvta_generic_to_shared(&s_slot[0][lane < nwarps ? lane : 0]);
    const uint32_t cslot_rd = (uint32_t)__cvta_generic_to_shared(DIRECT ? &c_slot[0][(SPL * lane) % NSLOT] : &c_slot[0][lane < CS ? lane : 0]);
    uint32_t cslot_peer = 0;                    // address of this CTA's (DIRECT: this warp's) slot inside CTA `tid` (DIRECT: `lane`) of the cluster
    if (DIRECT ? lane < CS : tid < CS) {
        const uint32_t mine = (uint32_t)__cvta_generic_to_shared(DIRECT ? &c_slot[0][rank * CW + warp] : &c_slot[0][rank]);
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(cslot_peer) : "r"(mine), "r"((uint32_t)(DIRECT ? lane : tid)));
    }
    for (int t = tid; t < 2 * NSLOT; t += T) (&c_slot[0][0])[t] = make_uint2(0u, 0u);       // tag 0 = "no iteration yet"
    cluster_sync_all();                         // every CTA of the cluster is resident (and its slots cleared) before anyone writes to a peer

    for (int i = 0; i < S; ++i) {
        float c[C];
#pragma unroll
        for (int k = 0; k < C; ++k) c[k] = s_pts[k * Nall + far];
        if (rank == 0) {
            if (tid == 0) {
                if (idx64) idx64[(int64_t)b * S + i] = far;
                if (idx32) idx32[(int64_t)b * S + i] = far;
            }
            if (new_points && tid < C) new_points[((int64_t)b * S + i) * C + tid] = s_pts[tid * Nall + far];
        }
        float best = -1.f;
        unsigned besti = 0xffffffffu;
        f32x2 nc[C];
#pragma unroll
        for (int k = 0; k < C; ++k) nc[k] = pk2(-c[k], -c[k]);
#pragma unroll
        for (int j = 0; j < PPT; j += 2) {
            f32x2 t[C];
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const f32x2 df = add2(pp[j / 2][k], nc[k]);
                t[k] = mul2(df, df);
            }
            const float2 d = unpk2(sum_order2<C, CONTIG>(t));
            dist[j] = fminf(dist[j], d.x);
            dist[j + 1] = fminf(dist[j + 1], d.y);
            if (dist[j] > best) { best = dist[j]; besti = (unsigned)(base + tid + j * T); }
            if (dist[j + 1] > best) { best = dist[j + 1]; besti = (unsigned)(base + tid + (j + 1) * T); }
        }
        const unsigned bits = __float_as_uint(best);
        const unsigned wmax = redux_max(bits);
        const unsigned widx = redux_min(bits == wmax ? besti : 0xffffffffu);
        const uint32_t cpar = (uint32_t)(i & 1) * (uint32_t)sizeof(c_slot[0]);
        const unsigned tag = ((unsigned)(i + 1) & 0xffffu) << 16;
        if constexpr (DIRECT) {
            // every warp posts its own result to all CTAs (lane l -> CTA l) and polls the CS * CW slots of its own CTA,
            // two per lane.  No block barrier in the loop: a warp rewrites a parity two iterations later, which needs every
            // warp's post of the iteration in between, and a warp posts only after it has read the previous iteration.
            if (lane < CS) asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(cslot_peer + cpar), "r"(wmax), "r"(widx | tag) : "memory");
            unsigned m2, i2;
            if constexpr (SPL == 2) {
                uint4 cl;
                bool ok;
                do {
                    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(cl.x), "=r"(cl.y), "=r"(cl.z), "=r"(cl.w) : "r"(cslot_rd + cpar) : "memory");
                    ok = (cl.y & 0xffff0000u) == tag && (cl.w & 0xffff0000u) == tag;
                } while (!__all_sync(0xffffffffu, ok));
                cl.y &= 0xffffu;
                cl.w &= 0xffffu;
                m2 = max(cl.x, cl.z);
                i2 = min(cl.x == m2 ? cl.y : 0xffffffffu, cl.z == m2 ? cl.w : 0xffffffffu);
            } else {
                uint2 cl;
                bool ok;
                do {
                    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cl.x), "=r"(cl.y) : "r"(cslot_rd + cpar) : "memory");
                    ok = (cl.y & 0xffff0000u) == tag;
                } while (!__all_sync(0xffffffffu, ok));
                m2 = cl.x;
                i2 = cl.y & 0xffffu;
            }
            const unsigned gmax = redux_max(m2);
            far = (int)redux_min(m2 == gmax ? i2 : 0xffffffffu);
        } else {
            const uint32_t par_off = (uint32_t)(i & 1) * (uint32_t)sizeof(s_slot[0]);
            if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot_wr + par_off), "r"(wmax), "r"(widx) : "memory");
            __syncthreads();
            uint2 sl;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(sl.x), "=r"(sl.y) : "r"(slot_rd + par_off) : "memory");
            if (lane >= nwarps) sl = make_uint2(0u, 0xffffffffu);
            const unsigned cmax = redux_max(sl.x);
            const unsigned cidx = redux_min(sl.x == cmax ? sl.y : 0xffffffffu);
            // this CTA's result -> slot [parity][rank] of every CTA of the cluster (thread t writes to CTA t)
            // (no cluster barrier: ~1000 cycles.)  The 8-byte slot carries the iteration number in the index word's upper half;
            // every warp polls its own CTA's slots until all CS of them show this iteration.  A slot of parity p is only
            // rewritten two iterations later, which needs every CTA's result of the iteration in between, i.e. every CTA has
            // finished reading this one.
            if (tid < CS) asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(cslot_peer + cpar), "r"(cmax), "r"(cidx | tag) : "memory");
            uint2 cl;
            bool ok;
            do {
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cl.x), "=r"(cl.y) : "r"(cslot_rd + cpar) : "memory");
                ok = (cl.y & 0xffff0000u) == tag;
            } while (!__all_sync(0xffffffffu, ok));
            cl.y &= 0xffffu;
            if (lane >= CS) cl = make_uint2(0u, 0xffffffffu);
            const unsigned gmax = redux_max(cl.x);
            far = (int)redux_min(cl.x == gmax ? cl.y : 0xffffffffu);
        }
    }
    cluster_sync_all();                         // no CTA may exit while a peer can still store into its slots
}

// returns -100 when the cluster variant does not apply
template <int C>
static int launch_fps_cluster(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                              const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, cudaStream_t st)
{
    // PN2_FPS_CLUSTER: 0 = off, 1 = 8 CTAs with per-CTA slots, 2 = 8 CTAs with per-warp slots, 4 = 4 CTAs with per-warp slots,
    // default (-1) = per-warp slots, cluster size by the number of clouds
    static const int env = getenv("PN2_FPS_CLUSTER") ? atoi(getenv("PN2_FPS_CLUSTER")) : -1;
    // few, large clouds only: the exchange costs more than it saves when a single CTA already has little to do
    // up to 64 clouds = 256 CTAs, two per SM (N = 8192 -> 2048: 778 us up to 37 clouds, 1050 - 1080 us at 48 - 64 against 1702 us on
    // one CTA per cloud; 74 clouds no longer fit one wave: 1675 us)
    static const int bmax = getenv("PN2_FPS_CLUSTER_BMAX") ? atoi(getenv("PN2_FPS_CLUSTER_BMAX")) : 64;
    if (!env || N < 4096 || N > 8192 || B > bmax) return -100;
    if (sc != 1 && (N & 31)) return -100;       // strided clouds with leftover points use the scalar order (single-CTA kernel)
    // 8-CTA clusters of more than a few clouds no longer find an SM per CTA (two CTAs share one: x1.15 per iteration at 12
    // clouds, x1.3 at 16); 4-CTA clusters of twice the points per thread run every B <= 37 at the same speed
    // (configs[2], B = 16, N = 8192 -> 2048: 1047 us with 8 CTAs and per-CTA slots, 981 us with per-warp slots, 778 us with 4 CTAs)
    // Other shapes measured at B = 16 (tools/fps_micro.py, profiles/r01_k_fps_cluster_variants.txt): 4 CTAs x 16 warps x 4 points
    // 909 us, 2 CTAs x 16 warps x 8 points 1074 us, 8 CTAs x 4 warps x 8 points 817 us (755 us up to 14 clouds), 4 CTAs x 4 warps x
    // 16 points 884 us: more warps per CTA lengthen the poll, fewer CTAs or warps lengthen the arithmetic.
    const int cs = env == 4 ? 4 : (env == 1 || env == 2) ? 8 : (B > 4 ? 4 : 8);
    const int threads = 256;
    const int per_cta = (N + cs - 1) / cs;
    const int ppt = per_cta <= threads * 2 ? 2 : per_cta <= threads * 4 ? 4 : 8;
    const size_t smem = sizeof(float) * (size_t)C * cs * ppt * threads;
    auto run = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("fps(cluster): smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * cs)); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, points, sb, sn, sc, N, start, S, idx64, idx32, new_points);
        if (e != cudaSuccess) { set_error("fps(cluster): launch: %s", cudaGetErrorString(e)); return (int)e; }
        return launch_status("pn2_fps_f32(cluster)");
    };
#define PN2_FPS_CL(P, D, CSZ, CWP) \
    return sc == 1 ? run(fps_cluster_kernel<C, P, true, D, CSZ, CWP>) : run(fps_cluster_kernel<C, P, false, D, CSZ, CWP>)
    if (cs == 4) {
        if (ppt == 2) PN2_FPS_CL(2, true, 4, 8);
        if (ppt == 4) PN2_FPS_CL(4, true, 4, 8);
        PN2_FPS_CL(8, true, 4, 8);
    }
    if (env == 1) {
        if (ppt == 2) PN2_FPS_CL(2, false, 8, 8);
        PN2_FPS_CL(4, false, 8, 8);
    }
    if (ppt == 2) PN2_FPS_CL(2, true, 8, 8);
    PN2_FPS_CL(4, true, 8, 8);
#undef PN2_FPS_CL
}

// Generic fallback: any C <= PN2_MAX_COORD, any N whose distance array fits in shared memory; points are
// re-read from global memory (L1/L2) every iteration.
__global__ void fps_generic_kernel(const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc, int N, int C,
                                   const int64_t *__restrict__ start, int S,
                                   int64_t *__restrict__ idx64, int32_t *__restrict__ idx32, float *__restrict__ new_points)
{
    extern __shared__ float s_dist[];           // [N]
    __shared__ uint2 s_slot[2][32];
    __shared__ float s_c[PN2_MAX_COORD];
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int b = blockIdx.x;
    const bool contig = (sc == 1);
    const float *P = points + (int64_t)b * sb;
    for (int n = tid; n < N; n += T) s_dist[n] = 1e10f;
    int far = (int)min(max(start[b], (int64_t)0), (int64_t)N - 1);
    __syncthreads();
    for (int i = 0; i < S; ++i) {
        if (tid < C) {
            const float v = __ldg(P + (int64_t)far * sn + (int64_t)tid * sc);
            s_c[tid] = v;
            if (new_points) new_points[((int64_t)b * S + i) * C + tid] = v;
        }
        if (tid == 0) {
            if (idx64) idx64[(int64_t)b * S + i] = far;
            if (idx32) idx32[(int64_t)b * S + i] = far;
        }
        __syncthreads();
        float best = -1.f;
        unsigned besti = 0xffffffffu;
        for (int n = tid; n < N; n += T) {
            float t[PN2_MAX_COORD];
            for (int k = 0; k < C; ++k) {
                const float df = __fsub_rn(__ldg(P + (int64_t)n * sn + (int64_t)k * sc), s_c[k]);
                t[k] = __fmul_rn(df, df);
            }
            const float d = sum_order_rt(t, C, contig, n >= (N & ~31));
            const float nd = fminf(s_dist[n], d);
            s_dist[n] = nd;
            if (nd > best) { best = nd; besti = (unsigned)n; }
        }
        const unsigned bits = (besti == 0xffffffffu) ? 0u : __float_as_uint(best);
        const unsigned wmax = redux_max(bits);
        const unsigned widx = redux_min(bits == wmax ? besti : 0xffffffffu);
        if (lane == 0) s_slot[i & 1][warp] = make_uint2(wmax, widx);
        __syncthreads();
        const uint2 sl = (lane < nwarps) ? s_slot[i & 1][lane] : make_uint2(0u, 0xffffffffu);
        const unsigned gmax = redux_max(sl.x);
        far = (int)redux_min(sl.x == gmax ? sl.y : 0xffffffffu);
    }
}

template <int C, int PPT, int TMAX>
static int launch_fps_t(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, int threads,
                        const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, cudaStream_t st)
{
    const size_t smem = sizeof(float) * (size_t)C * PPT * threads;
    auto run = [&](auto kern) -> int {
        if (smem > 40 * 1024) {   // static shared memory counts against the 48 KB default as well
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("fps: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        }
        kern<<<B, threads, smem, st>>>(points, sb, sn, sc, N, start, S, idx64, idx32, new_points);
        return launch_status("pn2_fps_f32");
    };
    return (sc == 1) ? run(fps_kernel<C, PPT, true, TMAX>) : run(fps_kernel<C, PPT, false, TMAX>);
}

// threads and points-per-thread: keep the per-thread serial work small (the loop is latency-bound) but do
// not pay for more warps in the barrier than the cloud needs.  Returns -100 when the cloud is not covered
// by the register-resident kernel (register budget: 64K / threads).
template <int C>
static int launch_fps_c(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                        const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, cudaStream_t st)
{
#define PN2_FPS_GO(P, TM, T) return launch_fps_t<C, P, TM>(points, sb, sn, sc, B, N, T, start, S, idx64, idx32, new_points, st)
    // Many clouds (the SMs are full either way): the kernel is issue-bound, and the per-warp cost of the argmax / barrier
    // (~60 instructions per iteration) outweighs the arithmetic of 4 points per thread (~34): 8 points per thread and half
    // the warps execute a third fewer instructions per cloud.  Few clouds stay latency-optimised (more threads).
    if constexpr (C <= 4) {
        if (B >= 4 * 148 && N > 256 && N <= 512)  PN2_FPS_GO(8, 512, 64);
        if (B >= 4 * 148 && N > 512 && N <= 1024) PN2_FPS_GO(8, 512, 128);      // 0.78 -> 0.54 ms at B = 2048 (16 per thread: 0.58)
    }
    if (N <= 128)  PN2_FPS_GO(2, 128, 64);
    if (N <= 256)  PN2_FPS_GO(2, 128, 128);
    if (N <= 512)  PN2_FPS_GO(4, 512, 128);
    if (N <= 1024) PN2_FPS_GO(4, 512, 256);
    if (N <= 2048) PN2_FPS_GO(4, 512, 512);
    if constexpr (C <= 4) {
        if (N <= 4096) PN2_FPS_GO(4, 1024, 1024);
        if (N <= 8192) PN2_FPS_GO(8, 1024, 1024);
    }
#undef PN2_FPS_GO
    return -100;
}

int launch_fps(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
               const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, cudaStream_t st)
{
    if (B == 0 || S == 0) return 0;
    int rc = -100;
    if (C == 3) rc = launch_fps_cluster<3>(points, sb, sn, sc, B, N, start, S, idx64, idx32, new_points, st);
    if (rc != -100) return rc;
    if (C == 3) rc = launch_fps_c<3>(points, sb, sn, sc, B, N, start, S, idx64, idx32, new_points, st);
    else if (C == 10) rc = launch_fps_c<10>(points, sb, sn, sc, B, N, start, S, idx64, idx32, new_points, st);
    if (rc != -100) return rc;
    const size_t smem = sizeof(float) * (size_t)N;
    PN2_REQUIRE(smem <= 200 * 1024, "pn2_fps_f32: N=%d too large for the generic kernel (C=%d)", N, C);
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fps_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("fps: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const int threads = N <= 256 ? 128 : (N <= 2048 ? 256 : 1024);
    fps_generic_kernel<<<B, threads, smem, st>>>(points, sb, sn, sc, N, C, start, S, idx64, idx32, new_points);
    return launch_status("pn2_fps_f32(generic)");
}

}  // namespace pn2
