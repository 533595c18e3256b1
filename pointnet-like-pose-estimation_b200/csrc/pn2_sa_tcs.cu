// K3 (fp32 mode) on the tensor cores -- fused set abstraction with SPLIT operands:
//   gather -> [4 x tcgen05.mma per K step on (hi, lo) operand pairs -> fp32 in TMEM -> +bias -> ReLU -> re-split to smem] x L -> max over K
// (reference: model/pointnet2_utils.py:166-172 and :204-218 at the reference's own precision; BatchNorm folded by pn2_pack_weights).
//
// Precision.  Every fp32 value v (activation or folded weight) is carried as v = hi + lo + d with hi = bf16(v) and
// lo = bf16(v - hi): |v - hi| <= 2^-9 |v| and the second rounding keeps 8 more bits, so |d| <= 2^-18 |v| -- both parts have fp32's
// exponent range, so nothing can overflow or flush that would not in fp32.  A product x w is the sum of the four exact
// tensor-core products xh wh + xh wl + xl wh + xl wl accumulated in fp32 in TMEM, leaving only the d terms: <= 2^-17 relative
// per product, random in sign over the K terms of a sum -- measured ~3e-6 of max|y| through three layers (the tolerance of this
// mode is 1e-5; fp32 FFMA sums in another order differ by 1e-7 .. 3e-7).  With PN2_F32_PASSES=3 the xl wl pass is dropped
// (another 2^-18 per product, ~5e-6).  (bf16 hi + fp16 lo would reach 5e-7, but kind::f16 rejects operands of different
// formats in one instruction: illegal instruction on sm_100a.)
//
// Shape.  Same tile / TMEM / epilogue structure as the streaming bf16 kernel (pn2_sa_tc.cu): 128 grouped rows = the 128 TMEM
// lanes, one row per epilogue thread, weights either resident or streamed through an mbarrier ring by a producer warp (one
// TMA multicast per chunk for the CTAs of a cluster).  Four passes make the kernel MMA-bound, which is what lets this simple
// lock-step schedule run near the tensor roof: per 128-row tile of sa2 the MMAs take ~8.4 k cycles against ~2 k for the
// gather and the three epilogues.  Operand bytes double (hi + lo), so wide stacks (-> 512 hidden channels) use 64-row
// tiles: the A descriptor still spans 128 rows (rows 64..127 alias the next k-group: garbage accumulator lanes nobody
// reads), the activations of a tile then fit beside the weight ring.
#include <stdlib.h>
#include <type_traits>
#include "pn2_sa_pipe.cuh"

namespace pn2 {

namespace {

struct TcsParams {
    const float *points; int64_t sb, sn, sc;
    const float *feat;   int64_t fb, fn, fc;
    const float *centres; const int32_t *idx32;
    int N, S, K, C, D, order;
    TcLayout lay;                    // split layout: every chunk = [bf16 hi image | bf16 lo image]
    const unsigned char *packed;
    float *out; int64_t ob, os; int co_off;
    int64_t G, tiles, n_iter;
    int TR;                          // rows per tile: 128, or 64 (wide stacks)
    int a_lbo;                       // bytes between k-groups of the A operand = TR * 16
    int gpt, cpg, Kl, direct, use_atomic, feat_fast;
    int tmem_cols, resident, nstage, stage_bytes, cs;
    int a_bytes[2];                  // activation buffers (hi + lo); a_bytes[1] == 0: one buffer, overwritten in place
    int a_lo[2];                     // offset of the lo copy inside each buffer
    int pool_cols, passes;
    int U;                           // tiles in flight per CTA: 1, or 2 teams of 8 compute warps sharing every streamed weight chunk
    int gather_depth;                // gather: 0 = cp.async + in-place conversion, 4 = register-staged loads, four 32-byte segments in flight
    unsigned long long *trace;       // pn2_debug_trace buffer (role 0 = compute warp 0, role 1 = producer warp; CTA 0 only)
};

// bf16 hi parts and bf16 lo parts (packed pairs) of two fp32 values
__device__ __forceinline__ void split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    hi = pack_bf16x2(a, b);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(a - ah, b - bh);                            // a - ah is exact in fp32
}
__device__ __forceinline__ void split8(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7, uint4 &hi, uint4 &lo)
{
    split2(v0, v1, hi.x, lo.x); split2(v2, v3, hi.y, lo.y);
    split2(v4, v5, hi.z, lo.z); split2(v6, v7, hi.w, lo.w);
}

// U = 1 (default): one tile per CTA at a time (two CTAs per SM overlap each other).  U = 2 (opt-in, measured slower, see
// tcs_plan): TWO TEAMS of eight compute warps, each with its own tile (activation buffers, TMEM half, pooling scratch), walk
// the layers in lock step and share every streamed weight chunk: team 0's first warp issues the chunk's MMAs for both tiles
// before the ring slot is released -- half the weight traffic per tile and a deeper ring.
template <int U>
__global__ void __launch_bounds__((U * TC_CWARPS + 1) * 32, U == 1 ? 2 : 1) sa_tcs_kernel(const TcsParams p)
{
    constexpr int NCT = U * TC_CTHREADS;                     // compute threads
    constexpr int NTHREADS = NCT + 32;
    extern __shared__ __align__(128) unsigned char smem_all[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp_all = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int team = (U == 2 && warp_all >= TC_CWARPS && warp_all < 2 * TC_CWARPS) ? 1 : 0;
    const int warp = warp_all - team * TC_CWARPS;            // warp inside the team (the producer warp reads as TC_CWARPS of team 0 ... U*TC_CWARPS)
    const int ttid = tid - team * TC_CTHREADS;               // thread inside the team
    const int a_team = p.a_bytes[0] + p.a_bytes[1];
    unsigned char *smem = smem_all + (size_t)team * a_team;  // this team's activation buffers
    const uint32_t a1_off = p.a_bytes[1] ? (uint32_t)p.a_bytes[0] : 0u;
    unsigned char *Wsm = smem_all + (size_t)U * a_team;
    const int w_region = p.resident ? p.lay.w_bytes : p.nstage * p.stage_bytes;
    float *bias_s = reinterpret_cast<float *>(Wsm + w_region);
    float *pool_s = bias_s + p.lay.bias_floats + (size_t)team * 4 * p.pool_cols;      // [4][pool_cols] per team
    int *s_n = reinterpret_cast<int *>(bias_s + p.lay.bias_floats + (size_t)U * 4 * p.pool_cols) + team * 2 * TC_ROWS;
    int *s_g = s_n + TC_ROWS;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<int *>(bias_s + p.lay.bias_floats + (size_t)U * 4 * p.pool_cols) + U * 2 * TC_ROWS);
    uint64_t *empty_bar = full_bar + TC_MAX_STAGES;
    uint64_t *acc_bar = empty_bar + TC_MAX_STAGES;
    uint64_t *slot_free = acc_bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(slot_free + TC_MAX_STAGES);

    const TcLayout &lay = p.lay;
    const int TR = p.TR, a_lbo = p.a_lbo, kblk = p.lay.kblk;
    const bool is_producer = warp_all == U * TC_CWARPS;
    // barriers: 1 + team = the team's eight warps; 3 = every compute warp of the CTA
    auto team_sync = [&]() {
        if (U == 1) cbar_sync();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(TC_CTHREADS) : "memory");
    };
    auto all_sync = [&]() {
        if (U == 1) cbar_sync();
        else asm volatile("bar.sync 3, %0;" ::"n"(NCT) : "memory");
    };

    if (tid == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); mbar_init(slot_free + s, (uint32_t)p.cs); }
        mbar_init(acc_bar, 1);
        fence_mbar_init();
    }
    if (warp_all == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    {
        const float *bsrc = reinterpret_cast<const float *>(p.packed + lay.w_bytes);
        for (int i = tid; i < lay.bias_floats; i += NTHREADS) bias_s[i] = __ldg(bsrc + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_all = *tmem_slot;
    const uint32_t tmem_base = tmem_all + (uint32_t)(team * (p.tmem_cols / U));       // this team's accumulator columns
    uint32_t cta_rank = 0;
    if (p.cs > 1) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }

    if (is_producer) {
        // ================================ producer warp: weight chunks (hi | lo) -> shared memory ==============================
        Tracer tr; tr.init(p.trace, 1, blockIdx.x == 0 && lane == 0);
        if (lane == 0) {
            if (p.resident) {
                mbar_expect_tx(full_bar, (uint32_t)lay.w_bytes);
                int off = 0;
                for (int l = 0; l < lay.L; ++l)
                    for (int sl = 0; sl < lay.ly[l].nslab; ++sl)
                        for (int kb = 0; kb < lay.ly[l].nkb; ++kb) {
                            const int cb = 2 * tc_chunk_bytes(lay.ly[l], kb, lay.kblk);
                            bulk_g2s(Wsm + off, p.packed + off, (uint32_t)cb, full_bar);
                            off += cb;
                        }
            } else {
                uint32_t q = 0;
                uint32_t free_remote = 0;
                if (p.cs > 1) asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(free_remote) : "r"(smem_u32(slot_free)));
                const uint16_t mask = (uint16_t)((1u << p.cs) - 1u);
                // this CTA's (cluster's) copy of the weight image (tc_weight_replicas)
                const unsigned char *wsrc = p.packed + (size_t)((blockIdx.x / (unsigned)p.cs) % (unsigned)lay.replicas) * lay.rep_stride;
                for (int64_t it = 0; it < p.n_iter; ++it) {
                    int off = 0;
                    for (int l = 0; l < lay.L; ++l)
                        for (int sl = 0; sl < lay.ly[l].nslab; ++sl)
                            for (int kb = 0; kb < lay.ly[l].nkb; ++kb, ++q) {
                                const int cb = 2 * tc_chunk_bytes(lay.ly[l], kb, lay.kblk);
                                const int s = q % p.nstage;
                                const uint32_t ph = (q / p.nstage) & 1u;
                                mbar_wait(empty_bar + s, ph ^ 1u);
                                tr.mark(0x900);
                                mbar_expect_tx(full_bar + s, (uint32_t)cb);
                                if (p.cs == 1) {
                                    bulk_g2s(Wsm + (size_t)s * p.stage_bytes, wsrc + off, (uint32_t)cb, full_bar + s);
                                } else {
                                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(free_remote + (uint32_t)(s * 8)) : "memory");
                                    if (cta_rank == 0) {
                                        mbar_wait_cluster(slot_free + s, ph);
                                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                                     ::"r"(smem_u32(Wsm + (size_t)s * p.stage_bytes)), "l"(wsrc + off), "r"((uint32_t)cb),
                                                       "r"(smem_u32(full_bar + s)), "h"(mask) : "memory");
                                    }
                                }
                                off += cb;
                            }
                }
            }
        }
    } else {
        // ================================ compute warps ========================================================================
        const int q4 = warp & 3;
        const int half = warp >> 2;
        const int row = q4 * 32 + lane;
        const bool row_in_tile = row < TR;
        const int K = p.K, C = p.C, D = p.D;
        uint32_t acc_phase = 0, wq = 0;
        if (p.resident && warp_all == 0) mbar_wait(full_bar, 0);
        Tracer tr; tr.init(p.trace, 0, blockIdx.x == 0 && warp_all == 0 && lane == 0);

        for (int64_t it = 0; it < p.n_iter; ++it) {
            const int64_t tile = (it * gridDim.x + blockIdx.x) * U + team;
            team_sync();
            tr.mark(0x100);
            if (ttid < TC_ROWS) {
                const int r = ttid;
                int64_t g = p.G; int j = 0;
                if (r < TR && tile < p.tiles) {
                    if (p.cpg == 1) { g = tile * p.gpt + r / K; j = r % K; if (r >= p.gpt * K) g = p.G; }
                    else            { g = tile / p.cpg; j = (int)(tile % p.cpg) * TR + r; }
                }
                int n = -1;
                if (g < p.G && j < K) n = (p.order == PN2_ORDER_ALL) ? j : __ldg(p.idx32 + g * K + j);
                if (n >= p.N) n = -1;            // 'no neighbour' marker of an empty ball (pointnet2_utils.py:85)
                s_n[r] = n; s_g[r] = (int)g;
            }
            team_sync();
            // ---- gather: layer-0 A operand (hi, lo), internal channel order [feat(D), xyz(C), zero pad] --------------------------
            {
                unsigned char *A0 = smem;
                const uint32_t lo_off = (uint32_t)p.a_lo[0];
                const int K0 = lay.ly[0].K;
                if (p.feat_fast) {
                    const int cpr = D / 8;
                    // DEPTH 32-byte segments per thread in flight before the first store: the loads are L2 / HBM latency (traced on
                    // B200: the gather was 13 k of the 44 k cycles of an sa2 tile with four in flight)
                    auto gather_feat = [&](auto depth_tag) {
                        constexpr int DEPTH = decltype(depth_tag)::value;
                        for (int i0 = ttid; i0 < TR * cpr; i0 += DEPTH * TC_CTHREADS) {
                            float4 x[DEPTH], y[DEPTH];
#pragma unroll
                            for (int u = 0; u < DEPTH; ++u) {
                                const int i = i0 + u * TC_CTHREADS;
                                x[u] = y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (i < TR * cpr) {
                                    const int r = i % TR, kg = i / TR;
                                    const int n = s_n[r];
                                    if (n >= 0) {
                                        const float *f = p.feat + (int64_t)(s_g[r] / p.S) * p.fb + (int64_t)n * p.fn + kg * 8;
                                        x[u] = __ldg(reinterpret_cast<const float4 *>(f));
                                        y[u] = __ldg(reinterpret_cast<const float4 *>(f) + 1);
                                    }
                                }
                            }
#pragma unroll
                            for (int u = 0; u < DEPTH; ++u) {
                                const int i = i0 + u * TC_CTHREADS;
                                if (i < TR * cpr) {
                                    const int r = i % TR, kg = i / TR;
                                    uint4 hi, lo;
                                    split8(x[u].x, x[u].y, x[u].z, x[u].w, y[u].x, y[u].y, y[u].z, y[u].w, hi, lo);
                                    unsigned char *dst = A0 + (size_t)kg * a_lbo + r * 16;
                                    *reinterpret_cast<uint4 *>(dst) = hi;
                                    *reinterpret_cast<uint4 *>(dst + lo_off) = lo;
                                }
                            }
                        }
                    };
                    if (p.gather_depth == 0) {
                        // asynchronous form: the raw fp32 segment (32 B = 8 channels of one row) is copied by cp.async straight into
                        // the two 16-byte slots its (hi, lo) images will occupy -- no register is held while the load is in flight, so
                        // ALL sixteen segments of a thread travel at once (one memory round trip per tile instead of four) -- and the
                        // segment is converted in place after the wait (one thread per segment reads both slots, then rewrites them)
                        // lane pairs cover the two 16-byte halves of one 32-byte sector in the SAME instruction (a lane that copied
                        // both halves in two instructions made L2 send every sector twice: cp.async.cg does not allocate in L1)
                        for (int i = ttid; i < 2 * TR * cpr; i += TC_CTHREADS) {
                            const int half = i & 1, seg = i >> 1;
                            const int r = seg % TR, kg = seg / TR;
                            const int n = s_n[r];
                            const float *f = p.feat + (int64_t)(s_g[r] / p.S) * p.fb + (int64_t)(n < 0 ? 0 : n) * p.fn + kg * 8 + half * 4;
                            unsigned char *dst = A0 + (size_t)kg * a_lbo + r * 16 + (half ? lo_off : 0u);
                            cp_async16_zfill(dst, n < 0 ? p.feat : f, n >= 0);
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    } else gather_feat(std::integral_constant<int, 4>{});
                    if (ttid < TR) {
                        const int r = ttid, n = s_n[r];
                        const int64_t g = s_g[r];
                        for (int kg = D / 8; kg < K0 / 8; ++kg) {
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int c = kg * 8 + i - D;
                                float x = 0.f;
                                if (n >= 0 && c < C) {
                                    x = __ldg(p.points + (g / p.S) * p.sb + (int64_t)n * p.sn + (int64_t)c * p.sc);
                                    if (p.order != PN2_ORDER_ALL) x = __fsub_rn(x, __ldg(p.centres + g * C + c));
                                }
                                v[i] = x;
                            }
                            uint4 hi, lo;
                            split8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], hi, lo);
                            unsigned char *dst = A0 + (size_t)kg * a_lbo + r * 16;
                            *reinterpret_cast<uint4 *>(dst) = hi;
                            *reinterpret_cast<uint4 *>(dst + lo_off) = lo;
                        }
                    }
                    if (p.gather_depth == 0) {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                        team_sync();                 // the two halves of a segment were copied by two lanes
                        for (int i = ttid; i < TR * cpr; i += TC_CTHREADS) {
                            const int r = i % TR, kg = i / TR;
                            unsigned char *dst = A0 + (size_t)kg * a_lbo + r * 16;
                            const float4 x = *reinterpret_cast<const float4 *>(dst), y = *reinterpret_cast<const float4 *>(dst + lo_off);
                            uint4 hi, lo;
                            split8(x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, hi, lo);
                            *reinterpret_cast<uint4 *>(dst) = hi;
                            *reinterpret_cast<uint4 *>(dst + lo_off) = lo;
                        }
                    }
                } else {
                    for (int i = ttid; i < TR * K0; i += TC_CTHREADS) {
                        const int r = i % TR, k = i / TR;
                        const int n = s_n[r];
                        float x = 0.f;
                        if (n >= 0) {
                            const int64_t g = s_g[r];
                            const int64_t b = g / p.S;
                            if (k < D) x = __ldg(p.feat + b * p.fb + (int64_t)n * p.fn + (int64_t)k * p.fc);
                            else if (k < D + C) {
                                const int c = k - D;
                                x = __ldg(p.points + b * p.sb + (int64_t)n * p.sn + (int64_t)c * p.sc);
                                if (p.order != PN2_ORDER_ALL) x = __fsub_rn(x, __ldg(p.centres + g * C + c));
                            }
                        }
                        const __nv_bfloat16 h = __float2bfloat16_rn(x);
                        unsigned char *dst = A0 + (size_t)(k / 8) * a_lbo + r * 16 + (k % 8) * 2;
                        *reinterpret_cast<__nv_bfloat16 *>(dst) = h;
                        *reinterpret_cast<__nv_bfloat16 *>(dst + lo_off) = __float2bfloat16_rn(x - __bfloat162float(h));
                    }
                }
            }
            fence_proxy_async();
            tr.mark(0x200);

            const bool row_valid = s_n[row] >= 0;
            for (int l = 0; l < lay.L; ++l) {
                const TcLayer ly = lay.ly[l];
                const bool last = (l == lay.L - 1);
                const int ib = (l & 1) ? 1 : 0;                     // input buffer of this layer
                const uint32_t in_off = ib ? a1_off : 0u, out_off = ib ? 0u : a1_off;
                const uint32_t a_in = smem_u32(smem + in_off);
                const uint32_t a_in_lo = a_in + (uint32_t)p.a_lo[a1_off ? ib : 0];
                unsigned char *a_out = smem + out_off;
                const uint32_t out_lo = (uint32_t)p.a_lo[a1_off ? (ib ^ 1) : 0];
                const int slabs_per_round = (p.tmem_cols / U) / ly.NS;
                const uint32_t idesc = instr_desc_bf16(ly.NS);
                for (int sl0 = 0; sl0 < ly.nslab; sl0 += slabs_per_round) {
                    const int sl1 = min(ly.nslab, sl0 + slabs_per_round);
                    tc_fence_before();
                    all_sync();            // the A operands of every team are complete / the previous round's TMEM reads and pool scratch drained
                    tr.mark(0x300 + l);
                    if (warp_all == 0) {
                        tc_fence_after();
                        for (int sl = sl0; sl < sl1; ++sl) {
                            const uint32_t d_tmem = tmem_all + (uint32_t)((sl - sl0) * ly.NS);
                            for (int kb = 0; kb < ly.nkb; ++kb) {
                                const int cb = tc_chunk_bytes(ly, kb, kblk);   // bytes of one (hi or lo) image of the chunk
                                uint32_t w_addr;
                                int s = 0;
                                if (p.resident) {
                                    w_addr = smem_u32(Wsm) + (uint32_t)(ly.w_off + 2 * (sl * ly.NS * ly.K * 2 + kb * ly.NS * kblk * 2));
                                } else {
                                    s = wq % p.nstage;
                                    mbar_wait(full_bar + s, (wq / p.nstage) & 1u);
                                    tc_fence_after();
                                    tr.mark(0x800);
                                    w_addr = smem_u32(Wsm) + (uint32_t)(s * p.stage_bytes);
                                    ++wq;
                                }
                                const int kw = min(kblk, ly.K - kb * kblk);
                                // team 0's addresses; team u's operands and accumulator lie u * a_team bytes / tmem_cols / U columns further
                                const uint32_t ah = desc_lo_ns(a_in + (uint32_t)(kb * (kblk / 8) * a_lbo), (uint32_t)a_lbo);
                                const uint32_t al = desc_lo_ns(a_in_lo + (uint32_t)(kb * (kblk / 8) * a_lbo), (uint32_t)a_lbo);
                                const uint32_t u_a16 = (uint32_t)(a_team >> 4), u_cols = (uint32_t)(p.tmem_cols / U);
                                const uint32_t bh = desc_lo_ns(w_addr, (uint32_t)(ly.NS * 16));
                                const uint32_t bl = desc_lo_ns(w_addr + (uint32_t)cb, (uint32_t)(ly.NS * 16));
                                const uint32_t a_step = (uint32_t)(2 * a_lbo / 16), b_step = (uint32_t)(2 * ly.NS);
                                if (elect_one()) {
#pragma unroll
                                    for (int u = 0; u < U; ++u) {
                                        const uint32_t dt = d_tmem + (uint32_t)u * u_cols;
                                        for (int j = 0; j < kw / 16; ++j) {
                                            const uint64_t dah = mk_desc64(ah + (uint32_t)u * u_a16 + (uint32_t)j * a_step, DESC_HI_NS);
                                            const uint64_t dal = mk_desc64(al + (uint32_t)u * u_a16 + (uint32_t)j * a_step, DESC_HI_NS);
                                            const uint64_t dbh = mk_desc64(bh + (uint32_t)j * b_step, DESC_HI_NS), dbl = mk_desc64(bl + (uint32_t)j * b_step, DESC_HI_NS);
                                            umma_bf16(dt, dah, dbh, idesc, (kb | j) ? 1u : 0u);
                                            umma_bf16(dt, dah, dbl, idesc, 1u);
                                            umma_bf16(dt, dal, dbh, idesc, 1u);
                                            if (p.passes > 3) umma_bf16(dt, dal, dbl, idesc, 1u);
                                        }
                                    }
                                    if (!p.resident) umma_commit(empty_bar + s);
                                }
                                __syncwarp();
                            }
                        }
                        if (elect_one()) umma_commit(acc_bar);
                    }
                    __syncwarp();
                    tr.mark(0x400 + l);
                    mbar_wait(acc_bar, acc_phase);
                    acc_phase ^= 1u;
                    tc_fence_after();
                    tr.mark(0x500 + l);

                    const int ncols = (sl1 - sl0) * ly.NS;
                    const int col_base = sl0 * ly.NS;
                    if (!last && !row_in_tile) {
                        // 64-row tiles: TMEM lanes 64..127 hold garbage (their A rows alias the next k-group); nothing to store
                    } else
                    for (int blk = half; blk < ncols / 32; blk += 2) {
                        float v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(blk * 32), v);
                        const int c0 = col_base + blk * 32;
                        if (!last) {
                            const float *bs = bias_s + ly.b_off + c0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) { const float t = v[i] + bs[i]; v[i] = t > 0.f ? t : (t != t ? t : 0.f); }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 hi, lo;
                                split8(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7], hi, lo);
                                unsigned char *dst = a_out + (size_t)(c0 / 8 + i) * a_lbo + row * 16;
                                *reinterpret_cast<uint4 *>(dst) = hi;
                                *reinterpret_cast<uint4 *>(dst + out_lo) = lo;
                            }
                        } else {
                            if (!row_valid) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = -INFINITY;
                            }
                            int o = p.Kl >> 1;
                            if (o) { rs_step<32>(v, o, (lane & o) != 0); o >>= 1; }
                            if (o) { rs_step<16>(v, o, (lane & o) != 0); o >>= 1; }
                            if (o) { rs_step<8>(v, o, (lane & o) != 0); o >>= 1; }
                            if (o) { rs_step<4>(v, o, (lane & o) != 0); o >>= 1; }
                            if (o) { rs_step<2>(v, o, (lane & o) != 0); }
                            const int nf = 32 / p.Kl;
                            const int cl = (lane & (p.Kl - 1)) * nf;
                            if (p.direct) {
                                const int64_t g = s_g[row];
                                if (g < p.G) {
                                    float *o_ptr = p.out + (g / p.S) * p.ob + (g % p.S) * p.os + p.co_off + c0 + cl;
                                    const float *bs = bias_s + ly.b_off + c0 + cl;
#pragma unroll
                                    for (int j = 0; j < 32; ++j)
                                        if (j < nf && c0 + cl + j < ly.Cout) o_ptr[j] = fmaxf(v[j] + bs[j], 0.f);
                                }
                            } else {
                                pool_s[q4 * p.pool_cols + blk * 32 + lane] = v[0];
                            }
                        }
                    }
                    if (last && !p.direct) {
                        team_sync();
                        for (int i = ttid; i < ncols * p.gpt; i += TC_CTHREADS) {
                            const int c = i % ncols, gl = i / ncols;
                            float m;
                            if (p.gpt == 2) m = fmaxf(pool_s[(2 * gl) * p.pool_cols + c], pool_s[(2 * gl + 1) * p.pool_cols + c]);
                            else m = fmaxf(fmaxf(pool_s[c], pool_s[p.pool_cols + c]), fmaxf(pool_s[2 * p.pool_cols + c], pool_s[3 * p.pool_cols + c]));
                            const int r0 = gl * (TC_ROWS / p.gpt);
                            const int64_t g = s_g[r0];
                            if (g >= p.G || col_base + c >= ly.Cout) continue;
                            m = fmaxf(m + bias_s[ly.b_off + col_base + c], 0.f);
                            float *o_ptr = p.out + (g / p.S) * p.ob + (g % p.S) * p.os + p.co_off + col_base + c;
                            if (p.use_atomic) atomicMax(reinterpret_cast<int *>(o_ptr), __float_as_int(m));
                            else *o_ptr = m;
                        }
                    }
                    if (!last) fence_proxy_async();
                    tr.mark(0x600 + l);
                }
            }
        }
        tc_fence_before();
        all_sync();
        if (warp_all == 0) tmem_dealloc(tmem_all, (uint32_t)p.tmem_cols);
    }
    if (p.cs > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct TcsPlan {
    TcLayout lay;
    int TR, a_bytes[2], a_lo[2], pool_cols, tmem_cols, resident, nstage, stage_bytes, ctas_per_sm, U;
    size_t smem;
    int gpt, cpg, Kl, direct, use_atomic;
};

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

bool tcs_plan(int K, int L, const int *dims, TcsPlan *pl)
{
    pl->lay = tc_layout(L, dims, true);
    const TcLayout &t = pl->lay;
    int max_ns = 32, max_mid = 0;
    for (int l = 0; l < L; ++l) {
        if (t.ly[l].NS > max_ns) max_ns = t.ly[l].NS;
        if (l < L - 1 && t.ly[l].Npad > max_mid) max_mid = t.ly[l].Npad;
    }
    const int last_npad = t.ly[L - 1].Npad;
    const size_t misc = (size_t)t.bias_floats * 4 + 2 * TC_ROWS * 4 + (3 * TC_MAX_STAGES + 1) * 8 + 16 + 16;
    for (int TR = TC_ROWS; TR >= 64; TR >>= 1) {
        // grouping mode (rows of a group inside a tile)
        if (K <= 32 && is_pow2(K)) { pl->gpt = TR / K; pl->cpg = 1; pl->Kl = K; pl->direct = 1; pl->use_atomic = 0; }
        else if (2 * K == TR)      { pl->gpt = 2; pl->cpg = 1; pl->Kl = 32; pl->direct = 0; pl->use_atomic = 0; }
        else if (K <= TR)          { pl->gpt = 1; pl->cpg = 1; pl->Kl = 32; pl->direct = 0; pl->use_atomic = 0; }
        else                       { pl->gpt = 1; pl->cpg = (K + TR - 1) / TR; pl->Kl = 32; pl->direct = 0; pl->use_atomic = 1; }
        // the cross-quarter pooling scratch pairs lane quarters (0,1) / (2,3) for two groups per tile: only right for 128 rows
        if (pl->gpt == 2 && TR != TC_ROWS) { pl->gpt = 1; pl->cpg = 1; }     // K == 32 is 'direct'; K == 64 with TR == 64 lands in K <= TR
        int pp[2] = {0, 0};
        for (int l = 0; l < L; ++l) {
            const int ab = t.ly[l].K / 8 * TR * 16;
            if (ab > pp[l & 1]) pp[l & 1] = ab;
        }
        if (pp[1] == 0) pp[1] = 16;
        auto eval = [&](int tmem_cols, int ctas, bool want_resident, bool inplace, TcsPlan *o, int U = 1) -> bool {
            *o = *pl;
            o->TR = TR; o->U = U;
            const int cols_team = tmem_cols / U;              // accumulator columns of one tile
            if (inplace && max_mid > cols_team) return false;
            if (max_ns > cols_team) return false;
            const int b0 = inplace ? (pp[0] > pp[1] ? pp[0] : pp[1]) : pp[0];
            const int b1 = inplace ? 0 : pp[1];
            o->a_bytes[0] = 2 * b0; o->a_bytes[1] = 2 * b1;
            o->a_lo[0] = b0; o->a_lo[1] = b1;
            const size_t fixed = (size_t)U * (o->a_bytes[0] + o->a_bytes[1]) + misc + (size_t)(U - 1) * 2 * TC_ROWS * 4;
            o->tmem_cols = tmem_cols; o->ctas_per_sm = ctas;
            o->pool_cols = o->direct ? 0 : (last_npad < cols_team ? last_npad : cols_team);
            const size_t base = fixed + (size_t)U * 4 * o->pool_cols * 4;
            const size_t budget = (size_t)TC_SMEM_LIMIT / ctas - (ctas > 1 ? 1024 : 0);
            if (want_resident) {
                if (base + t.w_bytes > budget) return false;
                o->resident = 1; o->nstage = 1; o->stage_bytes = 0; o->smem = base + t.w_bytes;
                return true;
            }
            if (base + 2 * (size_t)t.max_chunk > budget) return false;
            int ns = (int)((budget - base) / t.max_chunk);
            if (ns > TC_MAX_STAGES) ns = TC_MAX_STAGES;
            o->resident = 0; o->nstage = ns; o->stage_bytes = t.max_chunk; o->smem = base + (size_t)ns * t.max_chunk;
            return true;
        };
        TcsPlan c;
        const bool two_ok = max_ns <= 256;
        // two CTAs per SM first (one CTA's gather / epilogue under the other's MMAs: the schedule inside a CTA is lock-step),
        // with ping-pong activation buffers or, when those do not fit twice, one buffer overwritten in place
        if (two_ok && eval(256, 2, true, false, &c)) { *pl = c; return true; }
        if (two_ok && eval(256, 2, true, true, &c)) { *pl = c; return true; }
        if (eval(512, 1, true, false, &c)) { *pl = c; return true; }
        // streamed weights: two tiles per CTA sharing every chunk (one CTA per SM, a four-stage ring) is OPT-IN (PN2_F32_U2=1):
        // measured on sa2 (cls_ssg, B = 2048) it takes 13.4 ms against 10.7 ms for two independent CTAs per SM with two-stage
        // rings -- the lock-step barrier over 16 warps and the single issuing warp cost more than the shared chunks save
        static const bool u2 = getenv("PN2_F32_U2") && atoi(getenv("PN2_F32_U2")) == 1;
        if (u2 && two_ok && eval(512, 1, false, false, &c, 2) && c.nstage >= 4) { *pl = c; return true; }
        if (u2 && two_ok && eval(512, 1, false, true, &c, 2) && c.nstage >= 4) { *pl = c; return true; }
        if (two_ok && eval(256, 2, false, false, &c)) { *pl = c; return true; }
        if (two_ok && eval(256, 2, false, true, &c)) { *pl = c; return true; }
        if (eval(512, 1, false, false, &c)) { *pl = c; return true; }
        if (eval(512, 1, true, true, &c)) { *pl = c; return true; }
        if (eval(512, 1, false, true, &c)) { *pl = c; return true; }
    }
    return false;
}

}  // namespace

int64_t packed_bytes_f32_split(int L, const int *dims)
{
    const TcLayout t = tc_layout(L, dims, true);
    return t.replicas > 1 ? (int64_t)t.replicas * t.rep_stride : (int64_t)t.total_bytes;
}

// tile rows the split kernel would use for this stack (0: it does not cover it)
int sa_f32_split_tile_rows(int K, int L, const int *dims)
{
    static const bool off = getenv("PN2_F32_FFMA") != nullptr;
    if (off) return 0;
    TcsPlan pl;
    return tcs_plan(K, L, dims, &pl) ? pl.TR : 0;
}

// returns -100 when the tensor-core kernel does not cover the configuration (the caller falls back to the FFMA kernel)
int launch_sa_f32_split(const SaArgs &a, const void *packed_split, cudaStream_t st)
{
    static const bool off = getenv("PN2_F32_FFMA") != nullptr;
    if (off) return -100;
    {   // stacks without feature channels (sa1 of every model): the four-unit schedule of pn2_sa_tc2.cu with split operands
        const int rc = launch_sa_f32_split_units(a, packed_split, st);
        if (rc != -100) return rc;
    }
    TcsPlan pl;
    if (!tcs_plan(a.K, a.L, a.dims, &pl)) return -100;
    TcsParams p;
    p.points = a.points; p.sb = a.sb; p.sn = a.sn; p.sc = a.sc;
    p.feat = a.feat; p.fb = a.fb; p.fn = a.fn; p.fc = a.fc;
    p.centres = a.centres; p.idx32 = a.idx32;
    p.N = a.N; p.S = a.S; p.K = a.K; p.C = a.C; p.D = a.D; p.order = a.order;
    p.lay = pl.lay;
    p.packed = (const unsigned char *)packed_split;
    p.out = a.out; p.ob = a.ob; p.os = a.os; p.co_off = a.co_off;
    p.G = (int64_t)a.B * a.S;
    if (p.G >= (1ll << 31)) return -100;
    p.TR = pl.TR; p.a_lbo = pl.TR * 16;
    p.gpt = pl.gpt; p.cpg = pl.cpg; p.Kl = pl.Kl; p.direct = pl.direct; p.use_atomic = pl.use_atomic;
    p.tiles = pl.cpg > 1 ? p.G * pl.cpg : (p.G + pl.gpt - 1) / pl.gpt;
    p.feat_fast = (a.D == 0) || (a.fc == 1 && a.D % 8 == 0 && a.fn % 4 == 0 && a.fb % 4 == 0 && ((uintptr_t)a.feat & 15) == 0);
    p.tmem_cols = pl.tmem_cols; p.resident = pl.resident; p.nstage = pl.nstage; p.stage_bytes = pl.stage_bytes;
    p.a_bytes[0] = pl.a_bytes[0]; p.a_bytes[1] = pl.a_bytes[1]; p.a_lo[0] = pl.a_lo[0]; p.a_lo[1] = pl.a_lo[1];
    p.pool_cols = pl.pool_cols; p.U = pl.U;
    p.trace = get_debug_trace();
    // PN2_F32_GATHER: 4 (default) = register-staged loads, four 32-byte segments in flight (eight need 64 registers: under the
    // 96-register cap of two CTAs per SM ptxas serialises them again); 0 = cp.async into the operand slots + in-place conversion.
    // Traced on B200 (tools/ring_trace.py, sa2 of cls_ssg at B = 2048): the gather phase is 11-13 k of the 44 k cycles of a tile
    // either way -- it is bound by the index arithmetic and the hi / lo conversion (~40 instructions per segment on 16 warps per
    // SM) and by three dependent round trips (neighbour index -> row -> wait), not by bytes in flight: 10.9 ms against 11.1-11.3 ms
    static const int gd_env = getenv("PN2_F32_GATHER") ? atoi(getenv("PN2_F32_GATHER")) : 4;
    p.gather_depth = gd_env == 0 ? 0 : 4;
    static const int passes_env = getenv("PN2_F32_PASSES") ? atoi(getenv("PN2_F32_PASSES")) : 4;
    p.passes = passes_env == 3 ? 3 : 4;
    if (p.tiles == 0) return 0;
    auto kern = pl.U == 2 ? sa_tcs_kernel<2> : sa_tcs_kernel<1>;
    const unsigned nthreads = (unsigned)((pl.U * TC_CWARPS + 1) * 32);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) { set_error("sa_tcs: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e)); return (int)e; }
    int64_t grid = (p.tiles + pl.U - 1) / pl.U;
    const int64_t cap = (int64_t)148 * pl.ctas_per_sm;
    if (grid > cap) grid = cap;
    // streamed weights with one CTA per SM: the CTAs of a cluster share every chunk through one TMA multicast (the four-pass
    // MMAs of a chunk take long enough to hide the cluster round trip).  Measured (cls_ssg, B = 2048): sa3 2.80 ms with pairs,
    // 3.17 ms with clusters of four; with two CTAs per SM (sa2: a two-stage ring each) the lock-step handshake of a cluster
    // costs far more than the L2 traffic it saves: 10.7 ms alone, 20.3 / 22.9 ms in clusters of 2 / 4 (PN2_F32_CLUSTER forces it)
    static const int cs_env = getenv("PN2_F32_CLUSTER") ? atoi(getenv("PN2_F32_CLUSTER")) : 0;
    p.cs = 1;
    if (!pl.resident && (pl.ctas_per_sm == 1 || cs_env > 1)) {
        int cs = cs_env > 0 ? cs_env : 2;
        if (cs > 8) cs = 8;
        while (cs > 1 && (cs & (cs - 1))) --cs;
        while (cs > 1 && p.tiles < 2 * cs * pl.U) cs >>= 1;
        p.cs = cs;
        if (cs > 1) { const int64_t ncl = ((p.tiles + pl.U - 1) / pl.U + cs - 1) / cs, cap_cl = (int64_t)148 * pl.ctas_per_sm / cs; grid = (ncl < cap_cl ? ncl : cap_cl) * cs; }
    }
    if (p.cs == 1) {
        p.n_iter = (p.tiles + grid * pl.U - 1) / (grid * pl.U);
        kern<<<(unsigned)grid, nthreads, pl.smem, st>>>(p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(nthreads); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int max_clusters = 0;
        e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
        if (e != cudaSuccess || max_clusters < 1) { set_error("sa_tcs: cluster occupancy (cs=%d): %s", p.cs, cudaGetErrorString(e)); return e ? (int)e : -1; }
        if (grid / p.cs > max_clusters) grid = (int64_t)max_clusters * p.cs;
        cfg.gridDim = dim3((unsigned)grid);
        p.n_iter = (p.tiles + grid * pl.U - 1) / (grid * pl.U);
        e = cudaLaunchKernelEx(&cfg, kern, p);
        if (e != cudaSuccess) { set_error("sa_tcs: cluster launch (cs=%d): %s", p.cs, cudaGetErrorString(e)); return (int)e; }
    }
    return launch_status("pn2_sa_fused(f32, split tensor-core)");
}

}  // namespace pn2
