// K3 (bf16 mode) -- fused set abstraction on the 5th-generation tensor cores:
//   gather -> [tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> +bias -> ReLU -> bf16 back to smem] x L -> max over K
// (reference: model/pointnet2_utils.py:166-172 and :204-218; BatchNorm folded by pn2_pack_weights).
//
// Shape of the computation.  A tile is 128 grouped rows (= the 128 TMEM lanes = UMMA_M of a cta_group::1
// MMA): 128/K whole groups, or a 128-row slice of one large group.  Activations never leave the SM: the
// gathered rows are written as bf16 into a shared-memory A operand, every layer's accumulator lives in
// TMEM, the epilogue reads it back with tcgen05.ld (one row per thread), applies bias + ReLU and writes
// the next layer's A operand; after the last layer the K rows of a group are max-reduced with a
// warp-shuffle reduce-scatter (rows of a group are adjacent TMEM lanes) and only [groups x Cout] floats
// reach HBM.  Weights are pre-packed by pn2_pack_weights into exactly the shared-memory image the MMA wants
// and arrive through the TMA bulk-copy engine (cp.async.bulk + mbarrier), either once per CTA (all layers
// fit next to the activations) or through an mbarrier ring fed by a dedicated producer warp.
//
// Operand layout (both A and B): canonical K-major, no swizzle -- 8x8 bf16 "core matrices" of 128 contiguous
// bytes; rows at 16 B pitch inside a column of core matrices (SBO = 128 B), consecutive k-groups of 8 at
// LBO bytes.  For A, LBO = 128*16: a k-group column holds the 128 rows back to back, so a warp that writes
// 32 consecutive rows of one k-group (gather and epilogue both do) stores 512 contiguous bytes.
#include <stdlib.h>
#include <type_traits>
#include "pn2_sa_pipe.cuh"

namespace pn2 {

int64_t packed_bytes_bf16(int L, const int *dims)
{
    const TcLayout t = tc_layout(L, dims);
    return t.replicas > 1 ? (int64_t)t.replicas * t.rep_stride : (int64_t)t.total_bytes;
}

struct TcPackArgs {
    TcLayout lay;
    int dims[PN2_MAX_LAYERS + 1];
    const float *w[PN2_MAX_LAYERS], *b[PN2_MAX_LAYERS], *g[PN2_MAX_LAYERS], *be[PN2_MAX_LAYERS],
        *mu[PN2_MAX_LAYERS], *var[PN2_MAX_LAYERS];
    float eps;
    int C, D, order;   // layer-0 column permutation: internal order is [feat(D), xyz(C), 0-pad]
};

__global__ void pack_bf16_kernel(TcPackArgs a, unsigned char *__restrict__ packed)
{
    float *bias_out = reinterpret_cast<float *>(packed + a.lay.w_bytes);
    for (int l = 0; l < a.lay.L; ++l) {
        const TcLayer y = a.lay.ly[l];
        const int cin = a.dims[l];
        const int64_t nW = (int64_t)y.Npad * y.K;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nW + y.Npad; i += (int64_t)gridDim.x * blockDim.x) {
            const bool is_bias = i >= nW;
            const int n = is_bias ? (int)(i - nW) : (int)(i / y.K);
            const int k = is_bias ? 0 : (int)(i % y.K);
            float s = 0.f;
            if (n < y.Cout) s = __fdiv_rn(a.g[l][n], __fsqrt_rn(__fadd_rn(a.var[l][n], a.eps)));
            if (is_bias) {
                float v = 0.f;
                if (n < y.Cout) v = __fadd_rn(__fmul_rn(__fsub_rn(a.b[l][n], a.mu[l][n]), s), a.be[l][n]);
                bias_out[y.b_off + n] = v;
                if (l < a.lay.L - 1) {          // bias slice (see pn2_tc.cuh): three bf16 parts in k = 0, 1, 2
                    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                    const float r1 = __fsub_rn(v, __bfloat162float(hi));
                    const __nv_bfloat16 lo = __float2bfloat16_rn(r1);
                    const __nv_bfloat16 lo2 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(lo)));
                    __nv_bfloat16 *sl = reinterpret_cast<__nv_bfloat16 *>(packed + y.bs_off);
                    for (int k2 = 0; k2 < BIAS_SLICE_K; ++k2)
                        sl[(k2 / 8) * (y.Npad * 8) + n * 8 + (k2 % 8)] = k2 == 0 ? hi : (k2 == 1 ? lo : (k2 == 2 ? lo2 : __float2bfloat16_rn(0.f)));
                }
                continue;
            }
            int src = k;                       // column of the reference's weight matrix
            if (l == 0) {
                const int xo = (a.order == PN2_ORDER_MSG) ? a.D : 0;
                const int fo = (a.order == PN2_ORDER_MSG) ? 0 : a.C;
                if (k < a.D) src = fo + k;
                else if (k < a.D + a.C) src = xo + (k - a.D);
                else src = -1;
            } else if (k >= cin) src = -1;
            float v = 0.f;
            if (n < y.Cout && src >= 0) v = __fmul_rn(a.w[l][(int64_t)n * cin + src], s);
            // chunk (slab, kb) then canonical no-swizzle K-major inside the chunk
            const int kblk = a.lay.kblk;
            const int slab = n / y.NS, nl = n % y.NS, kb = k / kblk, kl = k % kblk;
            const int sp = a.lay.split ? 2 : 1;
            int64_t off = y.w_off + sp * ((int64_t)slab * y.NS * y.K * 2 + (int64_t)kb * y.NS * kblk * 2);
            if (a.lay.pair_halves) {      // [half][k-group][NS/2 rows][16 B]
                const int hr = y.NS / 2;
                off += (int64_t)(nl / hr) * (tc_chunk_bytes(y, kb, kblk) / 2) + (int64_t)(kl / 8) * (hr * 16) + (nl % hr) * 16 + (kl % 8) * 2;
            } else
            off += (int64_t)(kl / 8) * (y.NS * 16) + nl * 16 + (kl % 8) * 2;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            *reinterpret_cast<__nv_bfloat16 *>(packed + off) = hi;
            if (a.lay.split)        // fp32 mode: the bf16 remainder in the second image of the chunk (pn2_sa_tcs.cu)
                *reinterpret_cast<__nv_bfloat16 *>(packed + off + tc_chunk_bytes(y, kb, kblk)) = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
        }
    }
}

int launch_pack_bf16(int L, const int *dims, const void *const *ptrs, float eps, int C, int order, void *packed, cudaStream_t st, bool split)
{
    TcPackArgs a;
    a.lay = tc_layout(L, dims, split);
    a.eps = eps; a.C = C; a.D = dims[0] - C; a.order = order;
    for (int l = 0; l <= L; ++l) a.dims[l] = dims[l];
    for (int l = 0; l < L; ++l) {
        a.w[l] = (const float *)ptrs[6 * l + 0]; a.b[l] = (const float *)ptrs[6 * l + 1];
        a.g[l] = (const float *)ptrs[6 * l + 2]; a.be[l] = (const float *)ptrs[6 * l + 3];
        a.mu[l] = (const float *)ptrs[6 * l + 4]; a.var[l] = (const float *)ptrs[6 * l + 5];
    }
    pack_bf16_kernel<<<148, 256, 0, st>>>(a, (unsigned char *)packed);
    for (int r = 1; r < a.lay.replicas; ++r) {
        cudaError_t e = cudaMemcpyAsync((unsigned char *)packed + (size_t)r * a.lay.rep_stride, packed, (size_t)a.lay.total_bytes, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("pn2_pack_weights: replica copy: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return launch_status("pn2_pack_weights(bf16)");
}

// ------------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------------
struct TcParams {
    const float *points; int64_t sb, sn, sc;
    const float *feat;   int64_t fb, fn, fc;
    const __nv_bfloat16 *feat_bf16;      // optional bf16 copy of feat, [B,N,D] contiguous (D % 8 == 0): gathered instead of the fp32 rows
    const float *centres; const int32_t *idx32;
    int N, S, K, C, D, order;
    TcLayout lay;
    const unsigned char *packed;
    float *out; int64_t ob, os; int co_off;
    int64_t G;
    int gpt;           // groups per tile (1, 2 or 128/K)
    int cpg;           // 128-row chunks per group (K > 128) else 1
    int64_t tiles;
    int Kl;            // lanes of a warp that hold rows of the same group (min(K,32) for power-of-two K <= 32, else 32)
    int direct;        // K <= 32: pooled rows are stored straight from the warp
    int use_atomic;    // K > 128: combine tiles of a group with atomicMax
    int feat_fast;     // features can be gathered as 32-byte row segments
    int tmem_cols;
    int resident;      // all weight chunks stay in shared memory for the life of the CTA
    int nstage, stage_bytes;
    int cs;            // thread-block cluster size: the CTAs of a cluster share every streamed weight chunk (TMA multicast)
    int64_t n_iter;    // tile-loop trips, identical for every CTA (a CTA without a tile still takes part in the ring)
    int a_bytes[2];    // activation buffers (a_bytes[1] == 0: one buffer, each layer's epilogue overwrites its own input)
    int pool_cols;     // columns of the cross-warp pooling scratch (0 if direct)
    int pipelined_last;   // last layer transposed + pipelined over 128-column TMEM buffers (needs 128-channel slabs)
    int gather_depth;     // feature-row segments a thread keeps in flight during the gather (1 or 4)
    int pair;             // CTA-pair mode (tcgen05 cta_group::2, M = 256): the two CTAs of a cluster hold one tile each and HALF of
                          // every streamed weight chunk; CTA 0 issues the MMAs of both -- half the L2 -> SM weight bytes per tile
    unsigned long long *trace;   // pn2_debug_trace buffer (CTA 0: role 0 = compute warp 0, 1 = producer warp, 2 = pooling warp 4)
};

// PAIR: the CTA-pair form (tcgen05 cta_group::2).  A separate instantiation: a kernel that contains pair instructions can only be
// launched as a cluster of two (a plain launch fails with cudaErrorInvalidClusterSize), so the default kernel must not contain them.
template <bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, PAIR ? 1 : 2) sa_tc_kernel(const TcParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    // carve-up (all offsets multiples of 16)
    // the two activation buffers as offsets from `smem`: indexing an array of pointers with the layer number would demote
    // every activation store of the epilogue to a generic ST.E
    const uint32_t a1_off = p.a_bytes[1] ? (uint32_t)p.a_bytes[0] : 0u;
    unsigned char *Wsm = smem + p.a_bytes[0] + p.a_bytes[1];
    const int w_region = p.resident ? p.lay.w_bytes : p.nstage * p.stage_bytes;
    float *bias_s = reinterpret_cast<float *>(Wsm + w_region);
    float *pool_s = bias_s + p.lay.bias_floats;                 // [4][pool_cols]
    int *s_n = reinterpret_cast<int *>(pool_s + 4 * p.pool_cols);
    int *s_g = s_n + TC_ROWS;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(s_g + TC_ROWS);
    uint64_t *empty_bar = full_bar + TC_MAX_STAGES;
    uint64_t *acc_bar = empty_bar + TC_MAX_STAGES;
    uint64_t *slot_free = acc_bar + 1;      // in cluster rank 0: every CTA of the cluster has released ring slot s
    uint64_t *acc_full = slot_free + TC_MAX_STAGES;     // last layer, per 128-column TMEM buffer: slab's MMAs retired
    uint64_t *acc_empty = acc_full + 4;                  // ... and its pooling done (4 warp arrivals)
    uint64_t *peer_full = acc_empty + 4;                 // pair mode, in CTA 0: the peer's half of ring slot s has landed
    uint64_t *peer_ready = peer_full + TC_MAX_STAGES;    // pair mode, in CTA 0: the peer's operand is written / its accumulator drained
    uint32_t *vm_s = reinterpret_cast<uint32_t *>(peer_ready + 1);       // valid-row ballots of the tile
    uint32_t *tmem_slot = vm_s + 4;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler (MMA descriptors on the uniform datapath)
    const TcLayout &lay = p.lay;

    if (tid == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); mbar_init(slot_free + s, (uint32_t)p.cs); }
        mbar_init(acc_bar, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
        for (int s = 0; s < TC_MAX_STAGES; ++s) mbar_init(peer_full + s, 1);
        mbar_init(peer_ready, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    {
        const float *bsrc = reinterpret_cast<const float *>(p.packed + lay.w_bytes);
        for (int i = tid; i < lay.bias_floats; i += TC_THREADS) bias_s[i] = __ldg(bsrc + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint32_t cta_rank = 0;
    if (p.cs > 1 || PAIR) {
        // barriers of every CTA initialised before any remote arrive / multicast copy can reach them
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }

    // =====================================================================================================
    // producer warp: weights -> shared memory through the bulk-copy engine
    // =====================================================================================================
    if (warp == TC_CWARPS) {
        Tracer tr; tr.init(p.trace, 1, blockIdx.x == 0 && lane == 0);
        if (lane == 0) {
            if (p.resident) {
                mbar_expect_tx(full_bar, (uint32_t)lay.w_bytes);
                int off = 0;
                for (int l = 0; l < lay.L; ++l)
                    for (int sl = 0; sl < lay.ly[l].nslab; ++sl)
                        for (int kb = 0; kb < lay.ly[l].nkb; ++kb) {
                            const int cb = tc_chunk_bytes(lay.ly[l], kb);
                            bulk_g2s(Wsm + off, p.packed + off, (uint32_t)cb, full_bar);
                            off += cb;
                        }
            } else {
                // Ring of weight chunks.  In a cluster every CTA runs this loop in lock step (same chunk sequence); a
                // CTA frees slot s locally (its MMAs retired), posts the byte count on its own full barrier and
                // reports to rank 0; rank 0 issues ONE multicast copy per chunk once all CTAs have reported, which
                // lands in every CTA's slot and completes every CTA's full barrier: L2 -> SM weight traffic / cs.
                uint32_t q = 0;
                uint32_t free_remote = 0;
                if (p.cs > 1) asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(free_remote) : "r"(smem_u32(slot_free)));
                const uint16_t mask = (uint16_t)((1u << p.cs) - 1u);
                // this CTA's (cluster's) copy of the weight image: same bytes, different L2 lines than its neighbours read
                const unsigned char *wsrc = p.packed + (size_t)((blockIdx.x / (unsigned)p.cs) % (unsigned)lay.replicas) * lay.rep_stride;
                for (int64_t it = 0; it < p.n_iter; ++it) {
                    int off = 0;
                    for (int l = 0; l < lay.L; ++l)
                        for (int sl = 0; sl < lay.ly[l].nslab; ++sl)
                            for (int kb = 0; kb < lay.ly[l].nkb; ++kb, ++q) {
                                const int cb = tc_chunk_bytes(lay.ly[l], kb);
                                const int s = q % p.nstage;
                                const uint32_t ph = (q / p.nstage) & 1u;
                                mbar_wait(empty_bar + s, ph ^ 1u);
                                tr.mark(0x900);
                                if (PAIR) {
                                    // this CTA's half of the chunk: columns rank * NS/2 .. of every k-group (NS/2 rows of 16 B = one copy)
                                    const int half_rows = lay.ly[l].NS / 2, kgs = cb / (lay.ly[l].NS * 16);
                                    mbar_expect_tx(full_bar + s, (uint32_t)(cb / 2));
                                    if (lay.pair_halves)
                                        bulk_g2s(Wsm + (size_t)s * p.stage_bytes, wsrc + off + (size_t)cta_rank * (cb / 2), (uint32_t)(cb / 2), full_bar + s);
                                    else
                                    for (int kg = 0; kg < kgs; ++kg)
                                        bulk_g2s(Wsm + (size_t)s * p.stage_bytes + (size_t)kg * half_rows * 16,
                                                 wsrc + off + (size_t)kg * lay.ly[l].NS * 16 + (size_t)cta_rank * half_rows * 16,
                                                 (uint32_t)(half_rows * 16), full_bar + s);
                                    off += cb;
                                    continue;
                                }
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
        // the producer warp takes no further part (it must not fall through to the TMEM dealloc barrier
        // before its copies are consumed: the compute warps wait on every full barrier they were promised)
    } else {
        // =================================================================================================
        // compute warps
        // =================================================================================================
        const int q4 = warp & 3;            // TMEM lane quarter of this warp
        const int half = warp >> 2;         // which 32-column blocks of a round this warp drains
        const int row = q4 * 32 + lane;     // the tile row (= TMEM lane) this thread owns in the epilogue
        const int K = p.K, C = p.C, D = p.D;
        uint32_t acc_phase = 0, wq = 0, gsl = 0;      // gsl: last-layer slabs issued / pooled so far (TMEM buffer ring position)
        uint32_t pair_phase = 0, peer_ready_remote = 0, peer_full_remote = 0;
        if (PAIR) {
            asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(peer_ready_remote) : "r"(smem_u32(peer_ready)));
            asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(peer_full_remote) : "r"(smem_u32(peer_full)));
        }
        if (p.resident && warp == 0) mbar_wait(full_bar, 0);
        Tracer tr; tr.init(p.trace, warp == 4 ? 2 : 0, blockIdx.x == 0 && (warp == 0 || warp == 4) && lane == 0);

        for (int64_t it = 0; it < p.n_iter; ++it) {
            const int64_t tile = it * gridDim.x + blockIdx.x;     // beyond p.tiles: an all-padding tile (keeps the cluster's ring in step)
            cbar_sync();
            tr.mark(0x100);
            // ---- row bookkeeping ------------------------------------------------------------------------
            if (tid < TC_ROWS) {
                const int r = tid;
                int64_t g; int j;
                if (p.cpg == 1) { g = tile * p.gpt + r / K; j = r % K; if (r >= p.gpt * K) g = p.G; }
                else            { g = tile / p.cpg; j = (int)(tile % p.cpg) * TC_ROWS + r; }
                if (tile >= p.tiles) g = p.G;
                int n = -1;
                if (g < p.G && j < K) n = (p.order == PN2_ORDER_ALL) ? j : __ldg(p.idx32 + g * K + j);
                if (n >= p.N) n = -1;            // 'no neighbour' marker of an empty ball (pointnet2_utils.py:85)
                s_n[r] = n; s_g[r] = (int)g;
            }
            cbar_sync();
            // ---- gather: layer-0 A operand, internal channel order [feat(D), xyz(C), zero pad] ------------
            {
                unsigned char *A0 = smem;
                const int K0 = lay.ly[0].K;
                if (p.feat_fast) {
                    const int cpr = D / 8;                               // 16-byte chunks per row
                    if (p.feat_bf16) {
                        // the previous layer left a bf16 copy of its rows: 16 bytes = the 8 channels of one core-matrix row, copied by
                        // cp.async straight into the operand (no conversion, no register held while in flight, every segment of the
                        // tile on its way at once; half the bytes of the fp32 rows)
                        // lane pairs take the two 16-byte halves of one 32-byte sector (k-groups 2j, 2j+1 of a row) in the same instruction
                        for (int i = tid; i < TC_ROWS * ((cpr + 1) & ~1); i += TC_CTHREADS) {
                            const int kg = ((i >> 1) / TC_ROWS) * 2 + (i & 1), r = (i >> 1) % TC_ROWS;
                            const int n = s_n[r];
                            const __nv_bfloat16 *f = p.feat_bf16 + ((int64_t)(s_g[r] / p.S) * p.N + (n < 0 ? 0 : n)) * D + kg * 8;
                            if (kg < cpr) cp_async16_zfill(A0 + (size_t)kg * A_LBO + r * 16, n < 0 ? p.feat_bf16 : f, n >= 0);
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    } else if (p.gather_depth > 1) {
                        // DEPTH 32-byte segments per thread in flight before the first store: the loads are L2 / HBM latency and
                        // nothing else runs in this CTA during the gather (one CTA per SM for the wide stacks)
                        auto gather_feat = [&](auto depth_tag) {
                            constexpr int DEPTH = decltype(depth_tag)::value;
                            for (int i0 = tid; i0 < TC_ROWS * cpr; i0 += DEPTH * TC_CTHREADS) {
                                float4 x[DEPTH], y[DEPTH];
#pragma unroll
                                for (int u = 0; u < DEPTH; ++u) {
                                    const int i = i0 + u * TC_CTHREADS;
                                    x[u] = y[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (i < TC_ROWS * cpr) {
                                        const int r = i % TC_ROWS, kg = i / TC_ROWS;
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
                                    if (i < TC_ROWS * cpr) {
                                        const int r = i % TC_ROWS, kg = i / TC_ROWS;
                                        *reinterpret_cast<uint4 *>(A0 + (size_t)kg * A_LBO + r * 16) =
                                            make_uint4(pack_bf16x2(x[u].x, x[u].y), pack_bf16x2(x[u].z, x[u].w), pack_bf16x2(y[u].x, y[u].y), pack_bf16x2(y[u].z, y[u].w));
                                    }
                                }
                            }
                        };
                        gather_feat(std::integral_constant<int, 4>{});
                    } else
                    for (int it = tid; it < TC_ROWS * cpr; it += TC_CTHREADS) {
                        const int r = it % TC_ROWS, kg = it / TC_ROWS;
                        const int n = s_n[r];
                        uint4 o = make_uint4(0u, 0u, 0u, 0u);
                        if (n >= 0) {
                            const float *f = p.feat + (int64_t)(s_g[r] / p.S) * p.fb + (int64_t)n * p.fn + kg * 8;
                            const float4 x = __ldg(reinterpret_cast<const float4 *>(f));
                            const float4 y = __ldg(reinterpret_cast<const float4 *>(f) + 1);
                            o = make_uint4(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w), pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
                        }
                        *reinterpret_cast<uint4 *>(A0 + (size_t)kg * A_LBO + r * 16) = o;
                    }
                    if (tid < TC_ROWS) {                                  // xyz + zero padding: k in [D, K0)
                        const int r = tid, n = s_n[r];
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
                            *reinterpret_cast<uint4 *>(A0 + (size_t)kg * A_LBO + r * 16) =
                                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                        }
                    }
                    if (p.feat_bf16) asm volatile("cp.async.wait_group 0;" ::: "memory");
                } else {
                    // generic strides / D not a multiple of 8: one element at a time
                    for (int it = tid; it < TC_ROWS * K0; it += TC_CTHREADS) {
                        const int r = it % TC_ROWS, k = it / TC_ROWS;
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
                        *reinterpret_cast<__nv_bfloat16 *>(A0 + (size_t)(k / 8) * A_LBO + r * 16 + (k % 8) * 2) = __float2bfloat16_rn(x);
                    }
                }
            }
            fence_proxy_async();     // generic-proxy stores -> visible to the tensor core's async proxy
            tr.mark(0x200);

            const bool row_valid = s_n[row] >= 0;   // written before the last cbar_sync
            // ---- layers ---------------------------------------------------------------------------------------
            for (int l = 0; l < lay.L; ++l) {
                const TcLayer ly = lay.ly[l];
                const bool last = (l == lay.L - 1);
                const uint32_t a_in = smem_u32(smem + ((l & 1) ? a1_off : 0u));
                unsigned char *a_out = smem + ((l & 1) ? 0u : a1_off);
                const int slabs_per_round = p.tmem_cols / ly.NS;
                const uint32_t idesc = instr_desc_bf16(ly.NS);
                if (last && p.pipelined_last) {
                    // ---- last layer, TRANSPOSED and PIPELINED (128-channel slabs): D^T[channel][row] = W slab (M operand)
                    //      x X^T (the tile's 128 rows as N).  Warp 0 issues slab after slab into a ring of 128-column
                    //      TMEM buffers as the weight chunks arrive; warps 4-7 (one per lane quarter = 32 channels) pool
                    //      a finished slab in registers (max over the rows = TMEM columns, FMNMX3 tree, no shuffles) while
                    //      the next slabs' MMAs run; warps 1-3 rest.  (The non-pipelined form below issues a round of
                    //      slabs, waits, and reduces over rows with warp shuffles, all warps in lock step.) ------------
                    const int nbuf = p.tmem_cols / TC_ROWS;
                    tc_fence_before();
                    cbar_sync();                                   // the layer's A operand is complete
                    tr.mark(0x300 + l);
                    if (warp == 0) {
                        tc_fence_after();
                        for (int sl = 0; sl < ly.nslab; ++sl, ++gsl) {
                            const int buf = (int)(gsl % (uint32_t)nbuf);
                            const uint32_t use = gsl / (uint32_t)nbuf;
                            if (use > 0) { mbar_wait(acc_empty + buf, (use - 1u) & 1u); tc_fence_after(); }
                            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * TC_ROWS);
                            for (int kb = 0; kb < ly.nkb; ++kb) {
                                uint32_t w_addr;
                                int s = 0;
                                if (p.resident) {
                                    w_addr = smem_u32(Wsm) + (uint32_t)(ly.w_off + sl * ly.NS * ly.K * 2 + kb * ly.NS * KBLK * 2);
                                } else {
                                    s = wq % p.nstage;
                                    mbar_wait(full_bar + s, (wq / p.nstage) & 1u);
                                    tc_fence_after();
                                    tr.mark(0x800);
                                    w_addr = smem_u32(Wsm) + (uint32_t)(s * p.stage_bytes);
                                    ++wq;
                                }
                                const int kw = min(KBLK, ly.K - kb * KBLK);
                                const uint32_t a_lo = desc_lo_ns(a_in + (uint32_t)(kb * 8 * A_LBO), A_LBO);
                                const uint32_t b_lo = desc_lo_ns(w_addr, (uint32_t)(ly.NS * 16));
                                if (elect_one()) {
                                    for (int j = 0; j < kw / 16; ++j)      // weights as the M operand, activations as N
                                        umma_bf16(d_tmem, mk_desc64(b_lo + (uint32_t)(j * 2 * ly.NS), DESC_HI_NS),
                                                  mk_desc64(a_lo + (uint32_t)(j * 2 * A_LBO / 16), DESC_HI_NS), idesc, (kb | j) ? 1u : 0u);
                                    if (!p.resident) umma_commit(empty_bar + s);
                                }
                                __syncwarp();
                            }
                            if (elect_one()) umma_commit(acc_full + buf);
                            __syncwarp();
                            tr.mark(0x700 + sl);
                        }
                    } else if (warp >= 4) {
                        // valid-row ballots of the tile (rows without a neighbour pool as -inf), shared by the four pooling warps
                        const unsigned vmw = __ballot_sync(0xffffffffu, s_n[q4 * 32 + lane] >= 0);
                        if (lane == 0) vm_s[q4] = vmw;
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        PoolOut o;
                        o.all_valid = (vm_s[0] & vm_s[1] & vm_s[2] & vm_s[3]) == 0xffffffffu;
                        o.has_bf = false; o.pb = nullptr;
                        o.use_atomic = p.use_atomic != 0;
                        o.os = p.os; o.ob = p.ob; o.contig = p.ob == (int64_t)p.S * p.os; o.S = p.S; o.G = (int)p.G;
                        const int logKe = 31 - __clz(TC_ROWS / p.gpt);      // rows per group inside a tile
                        const int g_first = (int)(p.cpg == 1 ? tile * p.gpt : tile / p.cpg);
                        const bool tile_ok = tile < p.tiles;
                        for (int sl = 0; sl < ly.nslab; ++sl, ++gsl) {
                            const int buf = (int)(gsl % (uint32_t)nbuf);
                            const uint32_t use = gsl / (uint32_t)nbuf;
                            mbar_wait(acc_full + buf, use & 1u);
                            tc_fence_after();
                            const int c = sl * TC_ROWS + row;
                            o.bv = bias_s[ly.b_off + c];
                            o.c_ok = tile_ok && c < ly.Cout;
                            o.po = p.out + p.co_off + c;
                            o.gg = g_first;
                            pool_slice_any<false>(logKe, tmem_base + (uint32_t)(buf * TC_ROWS) + ((uint32_t)(q4 * 32) << 16), 0, 4, vm_s, o);
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(acc_empty + buf);
                            tr.mark(0xA00 + sl);
                        }
                    } else {
                        gsl += (uint32_t)ly.nslab;                 // warps 1-3: keep the slab counter in step
                    }
                    continue;
                }
                for (int sl0 = 0; sl0 < ly.nslab; sl0 += slabs_per_round) {
                    const int sl1 = min(ly.nslab, sl0 + slabs_per_round);
                    tc_fence_before();
                    cbar_sync();       // A operand complete / previous round's TMEM reads and pool scratch drained
                    tr.mark(0x300 + l);
                    if (PAIR && warp == 0) {
                        // ---- CTA pair: one instruction stream for two tiles.  The peer reports (a) that its operand is written
                        //      and its accumulator drained (this round's barrier above), (b) every half chunk that has landed
                        //      in its ring; CTA 0 issues M = 256 MMAs (rows 0..127 = its own tile, 128..255 = the peer's, the
                        //      N = NS weight columns half from each CTA's ring slot) and commits to the barriers of both ----
                        if (cta_rank != 0) {
                            if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer_ready_remote) : "memory");
                            for (int sl = sl0; sl < sl1; ++sl)
                                for (int kb = 0; kb < ly.nkb; ++kb, ++wq) {
                                    const int s = wq % p.nstage;
                                    mbar_wait(full_bar + s, (wq / p.nstage) & 1u);
                                    // relaxed: the chunk was written by the bulk-copy engine and is read by the tensor core, both through the
                                    // async proxy; a release at cluster scope here is a full fence per chunk (~1000 cycles: it WAS the chunk period)
                                    if (lane == 0) asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer_full_remote + (uint32_t)(s * 8)) : "memory");
                                    __syncwarp();
                                }
                        } else {
                            mbar_wait_cluster(peer_ready, pair_phase);
                            pair_phase ^= 1u;
                            tc_fence_after();
                            const uint32_t idesc2 = (idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
                            for (int sl = sl0; sl < sl1; ++sl) {
                                const uint32_t d_tmem = tmem_base + (uint32_t)((sl - sl0) * ly.NS);
                                for (int kb = 0; kb < ly.nkb; ++kb, ++wq) {
                                    const int s = wq % p.nstage;
                                    // CTA-scope waits: what the barriers order is the issue of MMAs that read the peer's shared memory through
                                    // the async proxy, not generic loads of this thread (a cluster-scope acquire per chunk cost ~100 cycles more)
                                    mbar_wait(peer_full + s, (wq / p.nstage) & 1u);
                                    mbar_wait(full_bar + s, (wq / p.nstage) & 1u);
                                    tc_fence_after();
                                    tr.mark(0x800);
                                    const uint32_t w_addr = smem_u32(Wsm) + (uint32_t)(s * p.stage_bytes);
                                    const int kw = min(KBLK, ly.K - kb * KBLK);
                                    const uint32_t a_lo = desc_lo_ns(a_in + (uint32_t)(kb * 8 * A_LBO), A_LBO);
                                    const uint32_t b_lo = desc_lo_ns(w_addr, (uint32_t)(ly.NS * 8));       // k-groups of NS/2 rows
                                    if (elect_one()) {
                                        for (int j = 0; j < kw / 16; ++j)
                                            umma2_bf16(d_tmem, mk_desc64(a_lo + (uint32_t)(j * 2 * A_LBO / 16), DESC_HI_NS),
                                                       mk_desc64(b_lo + (uint32_t)(j * ly.NS), DESC_HI_NS), idesc2, (kb | j) ? 1u : 0u);
                                        umma2_commit(empty_bar + s);      // frees the slot in BOTH rings when these MMAs retire
                                    }
                                    __syncwarp();
                                }
                            }
                            if (elect_one()) umma2_commit(acc_bar);
                        }
                    } else if (warp == 0) {
                        // the whole warp walks the chunk sequence (warp-uniform addresses -> uniform datapath), one
                        // elected lane issues: ~65 cycles per MMA instead of ~120 from a divergent single thread
                        tc_fence_after();
                        for (int sl = sl0; sl < sl1; ++sl) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)((sl - sl0) * ly.NS);
                            for (int kb = 0; kb < ly.nkb; ++kb) {
                                uint32_t w_addr;
                                int s = 0;
                                if (p.resident) {
                                    w_addr = smem_u32(Wsm) + (uint32_t)(ly.w_off + sl * ly.NS * ly.K * 2 + kb * ly.NS * KBLK * 2);
                                } else {
                                    s = wq % p.nstage;
                                    mbar_wait(full_bar + s, (wq / p.nstage) & 1u);
                                    tc_fence_after();
                                    tr.mark(0x800);
                                    w_addr = smem_u32(Wsm) + (uint32_t)(s * p.stage_bytes);
                                    ++wq;
                                }
                                const int kw = min(KBLK, ly.K - kb * KBLK);
                                const uint32_t a_lo = desc_lo_ns(a_in + (uint32_t)(kb * 8 * A_LBO), A_LBO);
                                const uint32_t b_lo = desc_lo_ns(w_addr, (uint32_t)(ly.NS * 16));
                                if (elect_one()) {
                                    for (int j = 0; j < kw / 16; ++j)
                                        umma_bf16(d_tmem, mk_desc64(a_lo + (uint32_t)(j * 2 * A_LBO / 16), DESC_HI_NS),
                                                  mk_desc64(b_lo + (uint32_t)(j * 2 * ly.NS), DESC_HI_NS), idesc, (kb | j) ? 1u : 0u);
                                    if (!p.resident) umma_commit(empty_bar + s);   // frees the ring slot when these MMAs retire
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

                    // ---- epilogue of this round: columns [sl0*NS, sl1*NS) -----------------------------------------
                    const int ncols = (sl1 - sl0) * ly.NS;
                    const int col_base = sl0 * ly.NS;               // column inside the layer
                    for (int blk = half; blk < ncols / 32; blk += 2) {
                        float v[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(blk * 32), v);
                        const int c0 = col_base + blk * 32;
                        if (!last) {
                            const float *bs = bias_s + ly.b_off + c0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bs[i], 0.f);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                *reinterpret_cast<uint4 *>(a_out + (size_t)(c0 / 8 + i) * A_LBO + row * 16) =
                                    make_uint4(pack_bf16x2(v[8 * i + 0], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                               pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
                        } else {
                            // max over the rows of each group; bias + ReLU commute with the max and are applied after
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
                            const int nf = 32 / p.Kl;               // columns left in this lane
                            const int cl = (lane & (p.Kl - 1)) * nf;
                            if (p.direct) {
                                const int64_t g = s_g[row];
                                if (g < p.G) {          // an empty ball (-inf in every row) pools to relu(-inf) = 0
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
                        cbar_sync();
                        // combine the 4 lane quarters: one group per tile (gpt == 1) or two (K == 64)
                        for (int it = tid; it < ncols * p.gpt; it += TC_CTHREADS) {
                            const int c = it % ncols, gl = it / ncols;
                            float m;
                            if (p.gpt == 2) m = fmaxf(pool_s[(2 * gl) * p.pool_cols + c], pool_s[(2 * gl + 1) * p.pool_cols + c]);
                            else m = fmaxf(fmaxf(pool_s[c], pool_s[p.pool_cols + c]), fmaxf(pool_s[2 * p.pool_cols + c], pool_s[3 * p.pool_cols + c]));
                            const int r0 = gl * (TC_ROWS / p.gpt);
                            const int64_t g = s_g[r0];
                            if (g >= p.G || col_base + c >= ly.Cout) continue;          // empty ball: m = -inf -> 0 below
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
        cbar_sync();
        if (warp == 0 && !PAIR) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
    // no CTA of a cluster may exit while a peer can still write its shared memory or signal its barriers
    if (p.cs > 1 || PAIR) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (PAIR && warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
}

// ------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------
struct TcPlan {
    TcLayout lay;
    int a_bytes[2], pool_cols, tmem_cols, resident, nstage, stage_bytes, ctas_per_sm;
    size_t smem;
    int gpt, cpg, Kl, direct, use_atomic;
};

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static bool tc_plan(int K, int L, const int *dims, TcPlan *pl)
{
    pl->lay = tc_layout(L, dims);
    const TcLayout &t = pl->lay;
    int pp[2] = {0, 0}, max_ns = 32, max_mid = 0;
    for (int l = 0; l < L; ++l) {
        const int ab = t.ly[l].K / 8 * A_LBO;
        if (ab > pp[l & 1]) pp[l & 1] = ab;
        if (t.ly[l].NS > max_ns) max_ns = t.ly[l].NS;
        if (l < L - 1 && t.ly[l].Npad > max_mid) max_mid = t.ly[l].Npad;
    }
    if (pp[1] == 0) pp[1] = 16;
    // grouping mode
    if (K <= 32 && is_pow2(K)) { pl->gpt = TC_ROWS / K; pl->cpg = 1; pl->Kl = K; pl->direct = 1; pl->use_atomic = 0; }
    else if (K == 64)          { pl->gpt = 2; pl->cpg = 1; pl->Kl = 32; pl->direct = 0; pl->use_atomic = 0; }
    else if (K <= TC_ROWS)     { pl->gpt = 1; pl->cpg = 1; pl->Kl = 32; pl->direct = 0; pl->use_atomic = 0; }
    else                       { pl->gpt = 1; pl->cpg = (K + TC_ROWS - 1) / TC_ROWS; pl->Kl = 32; pl->direct = 0; pl->use_atomic = 1; }
    const int last_npad = t.ly[L - 1].Npad;
    const size_t misc = (size_t)t.bias_floats * 4 + 2 * TC_ROWS * 4 + (4 * TC_MAX_STAGES + 2 + 8) * 8 + 16 + 16;
    // Candidates: two CTAs per SM with 256 TMEM columns each (one CTA's epilogue overlaps the other's MMAs),
    // or one CTA per SM with all 512 columns.  Weights stay resident in shared memory when they fit, else
    // they stream through an mbarrier ring.  Preference: resident x2, resident x1, ring x2, ring x1.
    // Activations ping-pong between two buffers, or -- when that does not fit -- live in ONE buffer that each
    // layer's epilogue overwrites in place.  In place is legal when every non-last layer finishes all its MMAs
    // before its epilogue starts, i.e. when its whole output fits the TMEM allocation (one round).
    auto eval = [&](int tmem_cols, int ctas, bool want_resident, bool inplace, TcPlan *o) -> bool {
        *o = *pl;
        if (inplace && max_mid > tmem_cols) return false;
        o->a_bytes[0] = inplace ? (pp[0] > pp[1] ? pp[0] : pp[1]) : pp[0];
        o->a_bytes[1] = inplace ? 0 : pp[1];
        const size_t fixed = (size_t)o->a_bytes[0] + o->a_bytes[1] + misc;
        o->tmem_cols = tmem_cols; o->ctas_per_sm = ctas;
        o->pool_cols = o->direct ? 0 : (last_npad < tmem_cols ? last_npad : tmem_cols);
        const size_t base = fixed + (size_t)4 * o->pool_cols * 4;
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
    TcPlan c;
    const bool two_ok = max_ns <= 256;
    if (two_ok && eval(256, 2, true, false, &c)) { *pl = c; return true; }
    if (eval(512, 1, true, false, &c)) { *pl = c; return true; }
    if (two_ok && eval(256, 2, false, false, &c)) { *pl = c; return true; }
    if (eval(512, 1, false, false, &c)) { *pl = c; return true; }
    if (eval(512, 1, true, true, &c)) { *pl = c; return true; }
    if (eval(512, 1, false, true, &c)) { *pl = c; return true; }
    return false;
}

int sa_bf16_needs_zero(int S, int K, int L, const int *dims)
{
    (void)S; (void)L; (void)dims;
    return K > TC_ROWS;
}

int launch_sa_bf16(const SaArgs &a, cudaStream_t st)
{
    // the pipelined kernel (pn2_sa_tc2.cu) covers every stack whose weights fit in shared memory; this file's
    // kernel streams weights through a ring and handles the wide group_all stacks and unusual group sizes
    static const bool force_ring = getenv("PN2_TC_RING_ONLY") != nullptr;
    if (!force_ring) {
        const int rc = launch_sa_bf16_pipelined(a, st);
        if (rc != -100) return rc;
    }
    static const bool twin_off = getenv("PN2_TC_RING_TWIN") && atoi(getenv("PN2_TC_RING_TWIN")) == 0;
    const bool use_twin = !twin_off && a.feat_bf16 && a.D > 0 && a.D % 8 == 0 && ((uintptr_t)a.feat_bf16 & 15) == 0;
    PN2_REQUIRE(a.D == 0 || a.feat || use_twin, "pn2_sa_fused(bf16): this configuration needs the fp32 feature rows (or bf16 rows with D %% 8 == 0)");
    // the streaming kernel has no bf16 side output: a caller that was promised one by pn2_sa_bf16_side_io (which cannot see the
    // batch size or PN2_TC_RING_ONLY) must not receive an unwritten tensor
    PN2_REQUIRE(a.out_bf16 == nullptr || !sa_bf16_pipelined_covers(a.K, a.C, a.D, a.order, a.L, a.dims),
                "pn2_sa_fused(bf16): B*S*K = %lld rows (or PN2_TC_RING_ONLY) exceed the pipelined kernel that writes out_bf16; pass out_bf16 = NULL or split the batch",
                (long long)a.B * a.S * a.K);
    TcPlan pl;
    PN2_REQUIRE(tc_plan(a.K, a.L, a.dims, &pl), "pn2_sa_fused(bf16): layer widths do not fit shared memory");
    TcParams p;
    p.points = a.points; p.sb = a.sb; p.sn = a.sn; p.sc = a.sc;
    p.feat = a.feat; p.fb = a.fb; p.fn = a.fn; p.fc = a.fc;
    p.centres = a.centres; p.idx32 = a.idx32;
    p.N = a.N; p.S = a.S; p.K = a.K; p.C = a.C; p.D = a.D; p.order = a.order;
    p.lay = pl.lay;
    p.packed = (const unsigned char *)a.packed;
    p.out = a.out; p.ob = a.ob; p.os = a.os; p.co_off = a.co_off;
    p.G = (int64_t)a.B * a.S;
    PN2_REQUIRE(p.G < (1ll << 31), "pn2_sa_fused(bf16): too many groups");
    p.gpt = pl.gpt; p.cpg = pl.cpg; p.Kl = pl.Kl; p.direct = pl.direct; p.use_atomic = pl.use_atomic;
    p.tiles = pl.cpg > 1 ? p.G * pl.cpg : (p.G + pl.gpt - 1) / pl.gpt;
    p.feat_fast = (a.D == 0) || (a.fc == 1 && a.D % 8 == 0 && a.fn % 4 == 0 && a.fb % 4 == 0 && ((uintptr_t)a.feat & 15) == 0);
    p.feat_bf16 = use_twin ? (const __nv_bfloat16 *)a.feat_bf16 : nullptr;     // PN2_TC_RING_TWIN=0: always the fp32 rows
    if (p.feat_bf16) p.feat_fast = 1;         // the twin is contiguous [B,N,D] whatever the strides of the fp32 rows
    p.tmem_cols = pl.tmem_cols; p.resident = pl.resident; p.nstage = pl.nstage; p.stage_bytes = pl.stage_bytes;
    p.a_bytes[0] = pl.a_bytes[0]; p.a_bytes[1] = pl.a_bytes[1]; p.pool_cols = pl.pool_cols;
    static const int pipe_env = getenv("PN2_RING_PIPELINED") ? atoi(getenv("PN2_RING_PIPELINED")) : 1;
    p.pipelined_last = (pipe_env && pl.lay.ly[a.L - 1].NS == TC_ROWS && pl.tmem_cols >= 2 * TC_ROWS && TC_ROWS / pl.gpt >= 8) ? 1 : 0;
    static const int gather_env = getenv("PN2_TC_GATHER") ? atoi(getenv("PN2_TC_GATHER")) : 4;
    p.gather_depth = gather_env == 1 ? 1 : 4;       // fp32 rows (no bf16 copy from the previous layer): segments in flight per thread
    p.trace = get_debug_trace();
    if (p.tiles == 0) return 0;
    cudaError_t e = cudaFuncSetAttribute(sa_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sa_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) { set_error("sa_tc: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e)); return (int)e; }
    int64_t grid = p.tiles;
    const int64_t cap = (int64_t)148 * pl.ctas_per_sm;
    if (grid > cap) grid = cap;
    // streamed weights with one CTA per SM: clusters of up to 8 CTAs share each chunk through TMA multicast
    static const int cs_env = getenv("PN2_TC_CLUSTER") ? atoi(getenv("PN2_TC_CLUSTER")) : 1;   // measured slower so far (profiles/README.md): off by default
    p.cs = 1;
    // CTA-pair mode (PN2_TC_PAIR=1, opt-in): streamed stacks with one CTA per SM; every chunk is fetched half by each CTA of a pair
    static const int pair_env = getenv("PN2_TC_PAIR") ? atoi(getenv("PN2_TC_PAIR")) : 0;
    p.pair = 0;
    PN2_REQUIRE(!pl.lay.pair_halves || !pl.resident, "pn2_sa_fused(bf16): pair-packed weights cannot be resident");
    if (pl.lay.pair_halves || (pair_env == 1 && !pl.resident && pl.ctas_per_sm == 1 && p.tiles >= 2)) {
        p.pair = 1;
        p.pipelined_last = 0;                       // the pair issues the plain (rows x channels) form of every layer
        p.stage_bytes = pl.stage_bytes / 2;         // half chunks: the same ring bytes hold twice the stages
        p.nstage = pl.nstage * 2 > TC_MAX_STAGES ? TC_MAX_STAGES : pl.nstage * 2;
        const int64_t npairs = (p.tiles + 1) / 2;
        grid = (npairs < 74 ? npairs : 74) * 2;
    }
    if (!p.pair && !pl.resident && pl.ctas_per_sm == 1) {
        int cs = cs_env < 1 ? 1 : (cs_env > 8 ? 8 : cs_env);
        while (cs > 1 && (cs & (cs - 1))) --cs;
        while (cs > 1 && p.tiles < 2 * cs) cs >>= 1;
        p.cs = cs;
        if (cs > 1) { const int64_t ncl = (p.tiles + cs - 1) / cs; grid = (ncl < 148 / cs ? ncl : 148 / cs) * cs; }
    }
    const int cluster = p.pair ? 2 : p.cs;
    if (cluster == 1) {
        p.n_iter = (p.tiles + grid - 1) / grid;
        sa_tc_kernel<false><<<(unsigned)grid, TC_THREADS, pl.smem, st>>>(p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = pl.smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        // persistent: no more clusters than can be resident at once (a cluster must fit inside one GPC)
        int max_clusters = 0;
        e = p.pair ? cudaOccupancyMaxActiveClusters(&max_clusters, sa_tc_kernel<true>, &cfg) : cudaOccupancyMaxActiveClusters(&max_clusters, sa_tc_kernel<false>, &cfg);
        if (e != cudaSuccess || max_clusters < 1) { set_error("sa_tc: cluster occupancy (cs=%d): %s", cluster, cudaGetErrorString(e)); return e ? (int)e : -1; }
        if (grid / cluster > max_clusters) grid = (int64_t)max_clusters * cluster;
        cfg.gridDim = dim3((unsigned)grid);
        p.n_iter = (p.tiles + grid - 1) / grid;
        e = p.pair ? cudaLaunchKernelEx(&cfg, sa_tc_kernel<true>, p) : cudaLaunchKernelEx(&cfg, sa_tc_kernel<false>, p);
        if (e != cudaSuccess) { set_error("sa_tc: cluster launch (cs=%d): %s", cluster, cudaGetErrorString(e)); return (int)e; }
    }
    {
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            set_error("pn2_sa_fused(bf16, streaming): %s (pair=%d cluster=%d grid=%lld tiles=%lld smem=%zu resident=%d ctas=%d nstage=%d)", cudaGetErrorString(le),
                      p.pair, cluster, (long long)grid, (long long)p.tiles, pl.smem, pl.resident, pl.ctas_per_sm, p.nstage);
            return (int)le;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// Self-test: one 128 x n_cols x k_dim product through exactly the descriptor / MMA / TMEM-load helpers
// above, operands made of small integers (exact in bf16), compared with a scalar loop in the same kernel.
// variant 0 = the no-swizzle descriptor convention of the product kernels; 1 = LBO and SBO swapped (diagnostic);
// 2 = A operand 128-byte swizzled (activations of the pipelined kernel); 3 = B operand 128-byte swizzled (its
// transposed last layer: weights as A, activations as B).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(int n_cols, int k_dim, int variant, float *max_err, float *dump)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_err[4];
    unsigned char *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);    // swizzled blocks need 1024-byte alignment
    const size_t a_bytes = variant == 2 ? (size_t)((k_dim + 63) / 64) * SW128_BLOCK : (size_t)(k_dim / 8) * A_LBO;
    unsigned char *As = base;                               // [k_dim/8][A_LBO], or swizzled blocks
    unsigned char *Bs = base + a_bytes;                     // [k_dim/8][n_cols*16], or swizzled blocks (variant 3, n_cols = 128)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto a_val = [](int r, int k) { return (float)(((r * 7 + k * 3) % 13) - 6); };
    auto b_val = [](int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4); };
    for (int i = tid; i < TC_ROWS * k_dim; i += 128) {
        const int r = i % TC_ROWS, k = i / TC_ROWS;
        const size_t off = variant == 2 ? (size_t)sw128_chunk_off(r, k / 8) + (k % 8) * 2 : (size_t)(k / 8) * A_LBO + r * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(As + off) = __float2bfloat16_rn(a_val(r, k));
    }
    for (int i = tid; i < n_cols * k_dim; i += 128) {
        const int n = i % n_cols, k = i / n_cols;
        const size_t off = variant == 3 ? (size_t)sw128_chunk_off(n, k / 8) + (k % 8) * 2 : (size_t)(k / 8) * (n_cols * 16) + n * 16 + (k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(Bs + off) = __float2bfloat16_rn(b_val(n, k));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = instr_desc_bf16(n_cols);
        for (int j = 0; j < k_dim / 16; ++j) {
            uint64_t ad, bd;
            if (variant == 0 || variant == 2 || variant == 3) {
                ad = smem_desc(smem_u32(As) + (uint32_t)(j * 2 * A_LBO), A_LBO, 128);
                bd = smem_desc(smem_u32(Bs) + (uint32_t)(j * 2 * n_cols * 16), (uint32_t)(n_cols * 16), 128);
                if (variant == 2) ad = smem_desc_sw128(smem_u32(As) + (uint32_t)((j / 4) * SW128_BLOCK + (j % 4) * 32));
                if (variant == 3) bd = smem_desc_sw128(smem_u32(Bs) + (uint32_t)((j / 4) * SW128_BLOCK + (j % 4) * 32));
            } else {
                ad = smem_desc(smem_u32(As) + (uint32_t)(j * 2 * A_LBO), 128, A_LBO);
                bd = smem_desc(smem_u32(Bs) + (uint32_t)(j * 2 * n_cols * 16), 128, (uint32_t)(n_cols * 16));
            }
            umma_bf16(tmem_base, ad, bd, idesc, j ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    float err = 0.f;
    const int row = warp * 32 + lane;
    for (int blk = 0; blk < n_cols / 32; ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int n = blk * 32 + i;
            float ref = 0.f;
            for (int k = 0; k < k_dim; ++k) ref += a_val(row, k) * b_val(n, k);
            err = fmaxf(err, fabsf(ref - v[i]));
            if (dump) dump[(size_t)row * n_cols + n] = v[i];
        }
    }
    for (int o = 16; o; o >>= 1) err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
    if (lane == 0) s_err[warp] = err;
    tc_fence_before();
    __syncthreads();
    if (tid == 0) *max_err = fmaxf(fmaxf(s_err[0], s_err[1]), fmaxf(s_err[2], s_err[3]));
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// variant 4: the CTA-PAIR form (tcgen05.mma.cta_group::2, M = 256).  A cluster of two CTAs; CTA r holds rows 128 r .. 128 r + 127
// of A and columns r n/2 .. of B (n/2 rows of the K-major B image) at the SAME shared-memory offsets; CTA 0 alone issues the
// MMAs and commits to the barrier of both CTAs; each CTA finds its 128 x n block of D in its own tensor memory.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma2_selftest_kernel(int n_cols, int k_dim, float *max_err, float *dump)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_err[4];
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    unsigned char *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const int nh = n_cols / 2;                               // this CTA's share of the B columns
    unsigned char *As = base;                                // [k_dim/8][128 rows][16 B]
    unsigned char *Bs = base + (size_t)(k_dim / 8) * A_LBO;  // [k_dim/8][nh rows][16 B]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto a_val = [](int r, int k) { return (float)(((r * 7 + k * 3) % 13) - 6); };
    auto b_val = [](int n, int k) { return (float)(((n * 5 + k * 11) % 9) - 4); };
    for (int i = tid; i < TC_ROWS * k_dim; i += 128) {
        const int r = i % TC_ROWS, k = i / TC_ROWS;
        *reinterpret_cast<__nv_bfloat16 *>(As + (size_t)(k / 8) * A_LBO + r * 16 + (k % 8) * 2) = __float2bfloat16_rn(a_val((int)rank * TC_ROWS + r, k));
    }
    for (int i = tid; i < nh * k_dim; i += 128) {
        const int n = i % nh, k = i / nh;
        *reinterpret_cast<__nv_bfloat16 *>(Bs + (size_t)(k / 8) * (nh * 16) + n * 16 + (k % 8) * 2) = __float2bfloat16_rn(b_val((int)rank * nh + n, k));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // both CTAs' operands written and fenced, both barriers initialised, before the leader issues
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (rank == 0 && tid == 0) {
        // instruction descriptor: D = f32, A = B = bf16, K-major, N = n_cols, M = 256
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int j = 0; j < k_dim / 16; ++j) {
            const uint64_t ad = smem_desc(smem_u32(As) + (uint32_t)(j * 2 * A_LBO), A_LBO, 128);
            const uint64_t bd = smem_desc(smem_u32(Bs) + (uint32_t)(j * 2 * nh * 16), (uint32_t)(nh * 16), 128);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(j ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    float err = 0.f;
    const int row = warp * 32 + lane;
    for (int blk = 0; blk < n_cols / 32; ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int n = blk * 32 + i;
            float ref = 0.f;
            for (int k = 0; k < k_dim; ++k) ref += a_val((int)rank * TC_ROWS + row, k) * b_val(n, k);
            err = fmaxf(err, fabsf(ref - v[i]));
            if (dump) dump[(size_t)((int)rank * TC_ROWS + row) * n_cols + n] = v[i];
        }
    }
    for (int o = 16; o; o >>= 1) err = fmaxf(err, __shfl_xor_sync(0xffffffffu, err, o));
    if (lane == 0) s_err[warp] = err;
    tc_fence_before();
    __syncthreads();
    if (tid == 0) atomicMax(reinterpret_cast<int *>(max_err), __float_as_int(fmaxf(fmaxf(s_err[0], s_err[1]), fmaxf(s_err[2], s_err[3]))));
    // neither CTA may leave (and release its tensor memory / shared memory) while the pair's MMAs or the peer's commit are in flight
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// TMEM -> register read bandwidth probe (variant >= 10 of pn2_selftest_umma): `warps` warps of one CTA each issue
// `reps` x tcgen05.ld.32x32b.x32 back to back (wait every `batch` loads); *max_err receives bytes per clock.
__global__ void __launch_bounds__(512, 1) tmem_bw_probe_kernel(int reps, int batch, float *result)
{
    __shared__ uint32_t tmem_slot;
    __shared__ long long t_begin[16], t_end[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += batch) {
        for (int j = 0; j < batch; ++j) {
            uint32_t r[32];
            tmem_ld32_nowait(base + (uint32_t)(((i + j) * 32) & 511), r);
            if (j == batch - 1) tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc ^= r[k];
        }
    }
    tmem_ld_wait();
    const long long t1 = clock64();
    if (lane == 0) { t_begin[warp] = t0; t_end[warp] = t1; }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        const int nw = blockDim.x >> 5;
        long long lo = t_begin[0], hi = t_end[0];
        for (int w = 1; w < nw; ++w) { lo = min(lo, t_begin[w]); hi = max(hi, t_end[w]); }
        result[0] = (float)((double)nw * reps * 4096.0 / (double)(hi - lo));
        result[1] = (float)(hi - lo);
    }
    if (acc == 0x12345678u) result[2] = 1.f;
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// tcgen05.mma throughput probe (variant 100 + 10*a_swizzled + b_swizzled [+ 50: two accumulators interleaved] of
// pn2_selftest_umma): one thread issues k_dim MMAs (M = 128, N = n_cols, K = 16) back to back into the same
// accumulator, then commits; dump[0] = cycles per MMA (issue to completion), dump[1] = cycles spent issuing.
__global__ void __launch_bounds__(128, 1) umma_rate_probe_kernel(int n_cols, int count, int a_sw, int b_sw, int two_acc, float *result)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    unsigned char *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char *As = base, *Bs = base + 4 * SW128_BLOCK;          // 256 channels of 128 rows each, zero-filled
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8 * SW128_BLOCK / 16; i += 128) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = instr_desc_bf16(n_cols);
        if (two_acc == 2) {
            // issue-rate floor: the same two descriptors every time, loop unrolled by 8, nothing else in the loop
            const uint64_t ad = smem_desc_sw128(smem_u32(As)), bd = smem_desc(smem_u32(Bs), (uint32_t)(n_cols * 16), 128);
            const long long t0 = clock64();
            for (int j = 0; j < count; j += 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) umma_bf16(tmem_slot, ad, bd, idesc, 1u);
            }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            result[0] = (float)((double)(t2 - t0) / count);
            result[1] = (float)((double)(t1 - t0) / count);
        } else {
        const long long t0 = clock64();
        for (int j = 0; j < count; ++j) {
            const int ks = j & 15;                                  // 16 K-steps = 256 channels, then wrap
            const uint64_t ad = a_sw ? smem_desc_sw128(smem_u32(As) + (uint32_t)((ks >> 2) * SW128_BLOCK + (ks & 3) * 32))
                                     : smem_desc(smem_u32(As) + (uint32_t)(ks * 2 * A_LBO), A_LBO, 128);
            const uint64_t bd = b_sw ? smem_desc_sw128(smem_u32(Bs) + (uint32_t)((ks >> 2) * SW128_BLOCK + (ks & 3) * 32))
                                     : smem_desc(smem_u32(Bs) + (uint32_t)(ks * 2 * n_cols * 16), (uint32_t)(n_cols * 16), 128);
            umma_bf16(tmem_slot + (two_acc ? (uint32_t)((j & 1) * 256) : 0u), ad, bd, idesc, 1u);
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        result[0] = (float)((double)(t2 - t0) / count);
        result[1] = (float)((double)(t1 - t0) / count);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

int launch_selftest_umma(int n_cols, int k_dim, int variant, float *max_err, float *dump, cudaStream_t st)
{
    if (variant >= 100) {  // MMA rate probe
        const int v = variant - 100, two = v >= 200 ? 2 : (v >= 50 ? 1 : 0), a_sw = ((v % 50) / 10) != 0, b_sw = (v % 10) != 0;
        PN2_REQUIRE(dump && n_cols >= 16 && n_cols <= 256 && n_cols % 16 == 0 && (!b_sw || n_cols == TC_ROWS), "pn2_selftest_umma(rate probe): bad arguments");
        const size_t smem = 1024 + 8 * (size_t)SW128_BLOCK;
        cudaError_t e = cudaFuncSetAttribute(umma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("rate probe: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        umma_rate_probe_kernel<<<1, 128, smem, st>>>(n_cols, k_dim, a_sw, b_sw, two, dump);
        return launch_status("pn2_selftest_umma(rate probe)");
    }
    if (variant >= 10) {   // TMEM read-bandwidth probe: n_cols = warps, k_dim = loads per warp, variant - 10 = loads per wait
        PN2_REQUIRE(n_cols >= 1 && n_cols <= 16 && k_dim >= 1 && dump, "pn2_selftest_umma(probe): bad arguments");
        tmem_bw_probe_kernel<<<1, n_cols * 32, 0, st>>>(k_dim, variant - 10 < 1 ? 1 : variant - 10, dump);
        return launch_status("pn2_selftest_umma(probe)");
    }
    PN2_REQUIRE(n_cols >= 32 && n_cols <= 256 && n_cols % 32 == 0, "pn2_selftest_umma: n_cols must be a multiple of 32 in [32,256]");
    PN2_REQUIRE(k_dim >= 16 && k_dim <= 256 && k_dim % 16 == 0, "pn2_selftest_umma: k_dim must be a multiple of 16 in [16,256]");
    if (variant == 4) {    // CTA pair, M = 256: max_err must be initialised to a negative value (both CTAs atomicMax into it); dump is [256][n_cols]
        const size_t smem2 = 1024 + (size_t)(k_dim / 8) * A_LBO + (size_t)(k_dim / 8) * (n_cols / 2) * 16;
        cudaError_t e2 = cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e2 != cudaSuccess) { set_error("selftest(pair): smem attribute: %s", cudaGetErrorString(e2)); return (int)e2; }
        umma2_selftest_kernel<<<2, 128, smem2, st>>>(n_cols, k_dim, max_err, dump);
        return launch_status("pn2_selftest_umma(pair)");
    }
    PN2_REQUIRE(variant != 3 || n_cols == TC_ROWS, "pn2_selftest_umma: variant 3 needs n_cols = 128");
    const size_t sw = (size_t)((k_dim + 63) / 64) * SW128_BLOCK;
    const size_t smem = 1024 + (variant == 2 ? sw : (size_t)(k_dim / 8) * A_LBO) + (variant == 3 ? sw : (size_t)(k_dim / 8) * n_cols * 16);
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("selftest: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    umma_selftest_kernel<<<1, 128, smem, st>>>(n_cols, k_dim, variant, max_err, dump);
    return launch_status("pn2_selftest_umma");
}

}  // namespace pn2
