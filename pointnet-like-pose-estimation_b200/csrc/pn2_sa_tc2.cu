// K3 (bf16 mode), pipelined variant -- the fused set-abstraction kernel for stacks whose weights fit in
// shared memory next to the activations (every sa1 / sa2 configuration of the reference's SSG models and most
// MSG branches).  Reference: model/pointnet2_utils.py:166-172 (SSG) and :204-218 (one MSG branch).
//
// A tile is 128 grouped rows (128 / K whole groups) = the 128 TMEM lanes of a cta_group::1, M = 128 tcgen05.mma.
// One persistent CTA per SM keeps U tiles in flight ("units": an activation buffer in shared memory, written in
// place by every layer's epilogue, and an accumulator region in TMEM) and walks each through
//     gather -> [MMA l -> epilogue l] x (L-1) -> MMA L-1 -> max-pool -> store
// All weights (BN folded, bf16, canonical K-major core-matrix layout, written by pn2_pack_weights) are
// bulk-copied (cp.async.bulk + mbarrier) into shared memory once per CTA.
//
//   hidden layer            : the folded bias is added INSIDE the MMA: one extra K = 16 step multiplies a constant
//                             activation slice (channels 0, 1 = 1.0) with the layer's "bias slice" (bf16 hi and lo
//                             parts of the fp32 bias, pn2_tc.cuh), so the epilogue is
//                             tcgen05.ld 32x32b.x32 (two in flight) -> cvt.rn.relu.bf16x2 -> st.shared.v4
//                             with compile-time swizzle offsets (one XOR per 16-byte chunk).
//   last layer + pooling    : the product is issued TRANSPOSED, D^T[channel][row] = W (M operand) x X^T (N operand;
//                             both are K-major in shared memory, so swapping roles costs nothing).  A thread then
//                             owns one output channel (TMEM lane) and the rows of a group are consecutive TMEM
//                             columns: the max over the K neighbours is a tree of FMNMX3 on the registers that
//                             tcgen05.ld returns, bias + ReLU (which commute with the max) are applied once per
//                             group, and a warp stores 32 consecutive channels of a group (128 B coalesced).
//   gather                  : features of the previous set-abstraction layer arrive in bf16 (side output of this
//                             kernel) by TMA tile::gather4 straight into the 128-byte-swizzled operand; fp32
//                             features / raw coordinates go through registers.
//
// Two schedules share those pieces:
//   sa_tc2_kernel  U (<= 4) units of four warps, thread = row; each unit's first warp issues the unit's MMAs (one
//                  elected lane, descriptors on the uniform datapath) after a unit-wide named barrier.
//   sa_tc3_kernel  for stacks that leave room for only TWO tiles per SM (sa2-sized): 16 worker warps serve both tiles
//                  in a fixed (layer, unit) schedule, each epilogue / pooling phase split 16 ways, one extra warp
//                  issues every MMA; the MMAs of one tile always run underneath the CUDA-core phase of the other.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "pn2_sa_pipe.cuh"

namespace pn2 {

// FMODE: how layer 0's feature channels arrive -- 0: none (D = 0), 1: bf16 rows by TMA gather4, 2: anything else
// (bf16 rows by cp.async, fp32 rows through registers, strided fp32).  PRE_XYZ: C <= 4, coordinates prefetched one
// tile ahead.  Separate instantiations keep the register allocation of the common cases free of spills (a spilled
// prefetch register turns the prefetch into a blocking load).
template <int FMODE, bool PRE_XYZ, bool SPLIT = false>
__global__ void __launch_bounds__(T2_MAX_THREADS, 1) sa_tc2_kernel(const __grid_constant__ Tc2Params p)
{
    static_assert(!SPLIT || FMODE == 0, "split operands (fp32 mode): stacks without feature channels only");
    extern __shared__ __align__(128) unsigned char smem[];
    const Tc2Smem sm = carve(smem, p);
    const int U = p.U;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (uniform datapath)
    const int L = p.L;

    cta_prologue(sm, p, tid, 4);
    if (warp == 0) tmem_alloc(sm.tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    const int tiles = (int)p.tiles;
    const int stride = (int)gridDim.x * U;

    {
        const int u = warp >> 2, q = warp & 3, row = q * 32 + lane;
        unsigned char *A = sm.A_base + (size_t)u * p.a_unit_bytes;
        const uint32_t a_addr = smem_u32(A);
        const uint32_t a_rowx = a_addr + (uint32_t)(row * 128 + ((row & 7) << 4));
        unsigned char *X_row = A + p.xyz_off + row * 16;            // this row in the coordinate block [2][128][16 B]
        uint32_t *vmask_u = sm.vmask + 8 * u;
        uint32_t parity = 0;
        const uint32_t tmem_row = tmem_base + (uint32_t)(u * p.cols_per_unit) + ((uint32_t)(q * 32) << 16);
        const int K = p.K, C = p.C, D = p.D, S = p.S, N = p.N;
        const int logK = 31 - __clz(K);
        const int gpt = TC_ROWS >> logK;
        const int rows = (int)p.rows, G = (int)p.G;
        uint32_t acc_phase = 0;
        const int gu = (int)blockIdx.x * U + u;
        uint64_t *my_acc_ready = sm.acc_ready + u;
        const bool twin = FMODE == 1 || (FMODE == 2 && p.feat_fast && p.feat_bf16 != nullptr);    // bf16 features: TMA / cp.async straight into the operand
        constexpr bool pre_xyz = PRE_XYZ;                           // coordinates prefetched one tile ahead

        // raw neighbour index of this thread's row in `tile` (-1 outside the problem); no use of the value here,
        // so the load stays in flight
        auto load_index = [&](int tile) -> int {
            const int r = tile * TC_ROWS + row;
            return (tile < tiles && r < rows) ? __ldg(p.idx32 + r) : -1;
        };
        // n == N is the 'no neighbour' marker of an empty ball (:85)
        auto is_valid = [&](int n) -> bool { return n >= 0 && n < N; };
        // coordinates of point n and of the group's centre (first 4 channels), raw
        auto load_xyz4 = [&](int tile, int n, float (&px)[4], float (&pc)[4]) {
            const int g = (tile * TC_ROWS + row) >> logK;
            const int b = (int)((unsigned)g / (unsigned)S);
            const float *pp = p.points + (int64_t)b * p.sb + (int64_t)n * p.sn;
            const float *cc = p.centres + (int64_t)g * C;
            const bool ok = is_valid(n);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                px[i] = pc[i] = 0.f;
                if (ok && i < C) { px[i] = __ldg(pp); pc[i] = __ldg(cc + i); }
                pp += p.sc;
            }
        };
        // Feature rows -> operand buffer (bf16 features, asynchronous; the buffer must be free when this is called).
        uint64_t *my_feat_ready = sm.feat_ready + u;
        uint32_t feat_phase = 0;
        auto issue_twin = [&](int tile, int n_mine) {
            if constexpr (FMODE == 0) return;
            if constexpr (FMODE == 1) {
                // TMA gather4: lane l < 8 moves rows 4l..4l+3 of this warp's 32 (one instruction per 64-channel
                // block); invalid rows use an out-of-range row index and arrive as zeros
                const int g = (tile * TC_ROWS + row) >> logK;
                const int b = (int)((unsigned)g / (unsigned)S);
                const int ri = is_valid(n_mine) ? b * N + n_mine : p.feat_rows;
                const int l4 = (lane & 7) * 4;
                const int r0 = __shfl_sync(0xffffffffu, ri, l4), r1 = __shfl_sync(0xffffffffu, ri, l4 + 1);
                const int r2 = __shfl_sync(0xffffffffu, ri, l4 + 2), r3 = __shfl_sync(0xffffffffu, ri, l4 + 3);
                if (q == 0 && lane == 0) mbar_expect_tx(my_feat_ready, (uint32_t)(TC_ROWS * D * 2));
                if (lane < 8)
                    for (int c0 = 0; c0 < D; c0 += 64)
                        tma_gather4(A + (c0 >> 6) * SW128_BLOCK + (q * 32 + l4) * 128, &p.tmap, c0, r0, r1, r2, r3, my_feat_ready);
                return;
            }
            // cp.async: lanes are mapped (4 rows) x (8 chunks of 16 B): eight lanes read 128 contiguous bytes of one
            // source row and write one 128-byte swizzled line (coalesced, conflict free)
            const int nv = is_valid(n_mine) ? n_mine : -1;
            const int ch = lane & 7;
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int rl = i * 4 + (lane >> 3);
                const int r = q * 32 + rl;
                const int nr = __shfl_sync(0xffffffffu, nv, rl);
                const int g = (tile * TC_ROWS + r) >> logK;
                const int b = (int)((unsigned)g / (unsigned)S);
                const __nv_bfloat16 *src = p.feat_bf16 + ((int64_t)b * N + (nr < 0 ? 0 : nr)) * D + ch * 8;
                unsigned char *dst = A + r * 128 + ((ch ^ (r & 7)) << 4);
                for (int c0 = 0; c0 < D; c0 += 64)
                    if (c0 + ch * 8 < D) cp_async16_zfill(dst + (c0 >> 6) * SW128_BLOCK, src + c0, nr >= 0);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // fp32 features: through registers (convert to bf16), same lane mapping
        auto gather_f32 = [&](int tile, int n_mine) {
            const int nv = is_valid(n_mine) ? n_mine : -1;
            const int ch = lane & 7;
            for (int c0 = 0; c0 < D; c0 += 64) {
                float4 v[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = i * 4 + (lane >> 3);
                    const int r = q * 32 + rl;
                    const int nr = __shfl_sync(0xffffffffu, nv, rl);
                    const int g = (tile * TC_ROWS + r) >> logK;
                    const int b = (int)((unsigned)g / (unsigned)S);
                    v[2 * i] = v[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (nr >= 0 && c0 + ch * 8 < D) {
                        const float *src = p.feat + (int64_t)b * p.fb + (int64_t)nr * p.fn + c0 + ch * 8;
                        v[2 * i] = __ldg(reinterpret_cast<const float4 *>(src));
                        v[2 * i + 1] = __ldg(reinterpret_cast<const float4 *>(src) + 1);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = q * 32 + i * 4 + (lane >> 3);
                    if (c0 + ch * 8 < D)
                        *reinterpret_cast<uint4 *>(A + (c0 >> 6) * SW128_BLOCK + r * 128 + ((ch ^ (r & 7)) << 4)) =
                            make_uint4(pack_bf16x2(v[2 * i].x, v[2 * i].y), pack_bf16x2(v[2 * i].z, v[2 * i].w),
                                       pack_bf16x2(v[2 * i + 1].x, v[2 * i + 1].y), pack_bf16x2(v[2 * i + 1].z, v[2 * i + 1].w));
                }
            }
        };

        Tracer tr; tr.init(p.trace, u, blockIdx.x == 0 && q == 0 && lane == 0);
        const uint32_t wsm16 = smem_u32(sm.Wsm) >> 4;
        const uint32_t a_sw = desc_lo_sw128(a_addr);
        const uint32_t a_xyz = desc_lo_ns(a_addr + (uint32_t)p.xyz_off, A_LBO);
        const uint32_t ones_lo = desc_lo_ns(smem_u32(sm.ones), A_LBO);
        const uint32_t d_tm = tmem_base + (uint32_t)(u * p.cols_per_unit);
        bool w_ready = false;
        // operand written (and proxy-fenced) / TMEM drained by this warp: meet the unit's other warps at a named
        // barrier, then the unit's first warp issues the layer
        auto issue_layer = [&](int l) {
            asm volatile("bar.sync %0, 128;" ::"r"(u + 1) : "memory");
            if (q == 0) {
                if (!w_ready) { mbar_wait(sm.w_bar, 0); w_ready = true; }
                tc_fence_after();
                if (elect_one()) {
                    if constexpr (SPLIT) issue_layer_mmas_split(p, l, wsm16, a_sw, a_xyz, ones_lo, d_tm, p.a_lo_off >> 4);
                    else issue_layer_mmas(p, l, wsm16, a_sw, a_xyz, ones_lo, d_tm);
                    umma_commit(my_acc_ready);
                }
                __syncwarp();
            }
        };
        // Prefetch discipline: fence.proxy.async is a MEMBAR + proxy fence and stalls on every global load this
        // thread still has in flight, so loads are only issued at the start of the pooling phase (the longest
        // fence-free stretch): the neighbour index of the tile after next and the coordinates of the next tile.
        int n_cur = load_index(gu);
        int n_next = load_index(gu + stride);
        float px[4], pc[4];
        if (gu < tiles) {
            if (twin) issue_twin(gu, n_cur);
            if (pre_xyz) load_xyz4(gu, n_cur, px, pc);
        }

        for (int tile = gu; tile < tiles; tile += stride) {
            uint32_t *vmask_s = vmask_u + 4 * parity;    // a warp can run at most one tile ahead of its unit
            parity ^= 1u;
            const int n = n_cur;
            int n_next2 = -1;
            const bool valid = is_valid(n);
            const int g = (tile * TC_ROWS + row) >> logK;      // group of this row
            const int b = (int)((unsigned)g / (unsigned)S);

            // ---- gather: features (swizzled blocks) were started earlier (bf16 twin) or are converted here (fp32);
            //      coordinates minus centre go to the unit's coordinate block: channels [D, D+16) of layer 0 ----------
            if constexpr (FMODE == 2) {
                if (p.feat_fast) { if (!twin) gather_f32(tile, n); }
                else {
                    for (int k = 0; k < D; ++k) {
                        const float t = valid ? __ldg(p.feat + (int64_t)b * p.fb + (int64_t)n * p.fn + (int64_t)k * p.fc) : 0.f;
                        *reinterpret_cast<__nv_bfloat16 *>(A + sw128_chunk_off(row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(t);
                    }
                }
            }
            if (pre_xyz) {
                const float d0 = __fsub_rn(px[0], pc[0]), d1 = __fsub_rn(px[1], pc[1]), d2 = __fsub_rn(px[2], pc[2]), d3 = __fsub_rn(px[3], pc[3]);
                const uint32_t h0 = pack_bf16x2(d0, d1), h1 = pack_bf16x2(d2, d3);
                *reinterpret_cast<uint4 *>(X_row) = make_uint4(h0, h1, 0u, 0u);
                *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(0u, 0u, 0u, 0u);
                if constexpr (SPLIT) {      // the bf16 remainders: second operand image
                    *reinterpret_cast<uint4 *>(X_row + p.a_lo_off) =
                        make_uint4(pack_bf16x2(d0 - __uint_as_float(h0 << 16), d1 - __uint_as_float(h0 & 0xffff0000u)),
                                   pack_bf16x2(d2 - __uint_as_float(h1 << 16), d3 - __uint_as_float(h1 & 0xffff0000u)), 0u, 0u);
                    *reinterpret_cast<uint4 *>(X_row + p.a_lo_off + A_LBO) = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                const float *pp = p.points + (int64_t)b * p.sb + (int64_t)n * p.sn;
                const float *cc = p.centres + (int64_t)g * C;
                float x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a0 = 0.f, a1 = 0.f;
                    if (valid && i < C) { a0 = __ldg(pp); a1 = __ldg(cc + i); }
                    pp += p.sc;
                    x[i] = __fsub_rn(a0, a1);
                }
                uint32_t h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
                *reinterpret_cast<uint4 *>(X_row) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(h[4], h[5], h[6], h[7]);
                if constexpr (SPLIT) {
                    uint32_t lw[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        lw[i] = pack_bf16x2(x[2 * i] - __uint_as_float(h[i] << 16), x[2 * i + 1] - __uint_as_float(h[i] & 0xffff0000u));
                    *reinterpret_cast<uint4 *>(X_row + p.a_lo_off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    *reinterpret_cast<uint4 *>(X_row + p.a_lo_off + A_LBO) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
                }
            }
            if constexpr (FMODE == 1) { mbar_wait(my_feat_ready, feat_phase); feat_phase ^= 1u; }
            if constexpr (FMODE == 2) { if (twin) asm volatile("cp.async.wait_group 0;" ::: "memory"); }
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) vmask_s[q] = vm;      // read by the whole unit in the pooling step (ordered by the barriers)
            fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core (async proxy)
            tr.mark(0x300);
            issue_layer(0);
            tr.mark(0x700);
            const bool more = tile + stride < tiles;

#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                mbar_wait(my_acc_ready, acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
                tr.mark(0x400 + l);
                if (l < L - 1) {
                    // ---- hidden layer: accumulator (bias included) -> ReLU -> bf16 -> the next layer's A operand (in place) --
                    if constexpr (SPLIT) epi_hidden_split(p.ly[l].nb32, tmem_row, a_rowx, p.a_lo_off);
                    else epi_hidden_any(p.ly[l].nb32, tmem_row, a_rowx);
                    tc_fence_before();           // TMEM reads ordered before the next MMA's write to this region
                    fence_proxy_async();
                    tr.mark(0x500 + l);
                    issue_layer(l + 1);
                    tr.mark(0x700 + l + 1);
                } else {
                    // ---- last layer (transposed): this thread = channel mh*128 + row; columns = the tile's rows ------
                    // every MMA that reads the operand buffer has completed: the next tile's feature rows can start
                    // flowing in underneath the pooling
                    if (more) {
                        if (twin) issue_twin(tile + stride, n_next);
                        if (pre_xyz) load_xyz4(tile + stride, n_next, px, pc);
                    }
                    n_next2 = load_index(tile + 2 * stride);
                    tr.mark(0x610);
                    const Tc2Layer &ly = p.ly[l];
                    const float *bs = sm.bias_s + ly.b_off;
                    if (!w_ready) { mbar_wait(sm.w_bar, 0); w_ready = true; }     // the fp32 biases arrive with the weight image
                    PoolOut o;
                    o.all_valid = (vmask_s[0] & vmask_s[1] & vmask_s[2] & vmask_s[3]) == 0xffffffffu;
                    o.has_bf = p.out_bf16 != nullptr;
                    o.os = p.os; o.ob = p.ob; o.contig = p.ob == (int64_t)S * p.os; o.S = S; o.G = G;
                    const int g0 = tile * gpt;
                    for (int mh = 0; mh < ly.n_mh; ++mh) {
                        const int c = mh * TC_ROWS + row;
                        o.bv = bs[c];
                        o.c_ok = c < ly.Cout;
                        o.po = p.out + p.co_off + c;
                        o.pb = p.out_bf16 + p.co_off + c;
                        o.gg = g0;
                        pool_slice_any<false>(logK, tmem_row + (uint32_t)(mh * TC_ROWS), 0, 4, vmask_s, o);
                    }
                    tc_fence_before();
                    tr.mark(0x600);
                }
            }
            n_cur = n_next; n_next = n_next2;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ======================================================================================================
// Cooperative variant, for stacks whose weights leave room for only TWO tiles per SM (sa2-sized).
//
// With two units of four warps each unit's chain (epilogue, pooling) is far longer than its MMAs, and the two
// chains tend to run in phase, so the tensor pipe idles.  Here all 16 worker warps serve BOTH tiles: the CTA
// walks a fixed schedule
//     for layer l: for unit u: wait acc(u,l) -> epilogue / pooling of (u,l) split 16 ways -> signal MMA warp
// so every CUDA-core phase is four times shorter and the MMAs of one unit always run underneath the phase of
// the other.  Warp w covers TMEM lane quarter w & 3 and column share w >> 2; threads of share 0 also own "their"
// row for the gather bookkeeping (neighbour index, coordinates, valid-row ballot).
// ======================================================================================================
constexpr int T3_WORKERS = 16;
constexpr int T3_THREADS = (T3_WORKERS + 1) * 32;      // + the MMA-issue warp

__global__ void __launch_bounds__(T3_THREADS, 1) sa_tc3_kernel(const __grid_constant__ Tc2Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const Tc2Smem sm = carve(smem, p);
    constexpr int U = 2;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = p.L;

    cta_prologue(sm, p, tid, T3_WORKERS);      // a_ready: 16 warp arrivals (+ the TMA bytes of a gather)
    if (warp == T3_WORKERS) tmem_alloc(sm.tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    const int tiles = (int)p.tiles;
    const int stride = (int)gridDim.x * U;
    const int first_tile = (int)blockIdx.x * U;
    const int n_iter = first_tile < tiles ? (tiles - first_tile + stride - 1) / stride : 0;   // trips of unit 0 (>= unit 1's)
    uint64_t *acc_ready = sm.acc_ready, *a_ready = sm.a_ready;

    if (warp == T3_WORKERS) {
        // ---- MMA warp: same schedule as the workers, one elected lane issues ---------------------------------
        const uint32_t wsm16 = smem_u32(sm.Wsm) >> 4;
        const uint32_t ones_lo = desc_lo_ns(smem_u32(sm.ones), A_LBO);
        mbar_wait(sm.w_bar, 0);
        uint32_t phases = 0;
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (first_tile + u + it * stride >= tiles) continue;
                    mbar_wait(a_ready + u, (phases >> u) & 1u);
                    phases ^= 1u << u;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sm.A_base) + (uint32_t)(u * p.a_unit_bytes);
                    if (elect_one()) {
                        issue_layer_mmas(p, l, wsm16, desc_lo_sw128(a_addr), desc_lo_ns(a_addr + (uint32_t)p.xyz_off, A_LBO), ones_lo,
                                         tmem_base + (uint32_t)(u * p.cols_per_unit));
                        umma_commit(acc_ready + u);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- worker warps -----------------------------------------------------------------------------------
        const int q = warp & 3, hb = warp >> 2, row = q * 32 + lane;
        const int K = p.K, C = p.C, D = p.D, S = p.S, N = p.N;
        const int logK = 31 - __clz(K);
        const int gpt = TC_ROWS >> logK;
        const int rows = (int)p.rows, G = (int)p.G;
        const bool pre_xyz = C <= 4;
        const int n_blk_feat = D / 64;                  // the coop variant is only launched with TMA-gathered features (D % 64 == 0) or D == 0
        uint32_t acc_phase = 0;
        bool w_ready = false;
        Tracer tr; tr.init(p.trace, 0, blockIdx.x == 0 && warp == 0 && lane == 0);

        auto load_index = [&](int tile) -> int {
            const int r = tile * TC_ROWS + row;
            return (tile < tiles && r < rows) ? __ldg(p.idx32 + r) : -1;
        };
        auto is_valid = [&](int n) -> bool { return n >= 0 && n < N; };
        auto load_xyz4 = [&](int tile, int n, float (&px)[4], float (&pc)[4]) {
            const int g = (tile * TC_ROWS + row) >> logK;
            const int b = (int)((unsigned)g / (unsigned)S);
            const float *pp = p.points + (int64_t)b * p.sb + (int64_t)n * p.sn;
            const float *cc = p.centres + (int64_t)g * C;
            const bool ok = is_valid(n);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                px[i] = pc[i] = 0.f;
                if (ok && i < C) { px[i] = __ldg(pp); pc[i] = __ldg(cc + i); }
                pp += p.sc;
            }
        };
        // feature rows of `tile` -> unit u's swizzled blocks by TMA gather4, spread evenly over the 16 warps: warp
        // (q, hb) moves rows q*32 + hb*8 .. +8 (lanes 0 and 1: 4 rows each), all 64-channel blocks; one thread posts
        // the byte count on the unit's barrier
        auto issue_gather = [&](int u, int tile, int n_mine) {
            if (n_blk_feat == 0) return;
            unsigned char *A = sm.A_base + (size_t)u * p.a_unit_bytes;
            const int g = (tile * TC_ROWS + row) >> logK;
            const int b = (int)((unsigned)g / (unsigned)S);
            const int ri = is_valid(n_mine) ? b * N + n_mine : p.feat_rows;
            const int l4 = hb * 8 + (lane & 1) * 4;
            const int r0 = __shfl_sync(0xffffffffu, ri, l4), r1 = __shfl_sync(0xffffffffu, ri, l4 + 1);
            const int r2 = __shfl_sync(0xffffffffu, ri, l4 + 2), r3 = __shfl_sync(0xffffffffu, ri, l4 + 3);
            if (warp == 0 && lane == 0) mbar_expect_tx_only(a_ready + u, (uint32_t)(TC_ROWS * D * 2));
            if (lane < 2)
                for (int blk = 0; blk < n_blk_feat; ++blk)
                    tma_gather4(A + blk * SW128_BLOCK + (q * 32 + l4) * 128, &p.tmap, blk * 64, r0, r1, r2, r3, a_ready + u);
        };
        // coordinates minus centre (+ zero padding) of this thread's row -> unit u's coordinate block; valid-row ballot
        auto finish_gather = [&](int u, int tile, int n, const float (&px)[4], const float (&pc)[4], uint32_t par) {
            if (hb != 0) return;
            unsigned char *X_row = sm.A_base + (size_t)u * p.a_unit_bytes + p.xyz_off + row * 16;
            const bool valid = is_valid(n);
            if (pre_xyz) {
                *reinterpret_cast<uint4 *>(X_row) = make_uint4(pack_bf16x2(__fsub_rn(px[0], pc[0]), __fsub_rn(px[1], pc[1])),
                                                               pack_bf16x2(__fsub_rn(px[2], pc[2]), __fsub_rn(px[3], pc[3])), 0u, 0u);
                *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                const int g = (tile * TC_ROWS + row) >> logK;
                const int b = (int)((unsigned)g / (unsigned)S);
                const float *pp = p.points + (int64_t)b * p.sb + (int64_t)n * p.sn;
                const float *cc = p.centres + (int64_t)g * C;
                float x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a0 = 0.f, a1 = 0.f;
                    if (valid && i < C) { a0 = __ldg(pp); a1 = __ldg(cc + i); }
                    pp += p.sc;
                    x[i] = __fsub_rn(a0, a1);
                }
                *reinterpret_cast<uint4 *>(X_row) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15]));
            }
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) sm.vmask[8 * u + 4 * par + q] = vm;
        };
        auto arrive = [&](int u) {
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready + u);
        };

        // ---- prologue: first tile of each unit ----------------------------------------------------------------
        int n_cur[U], n_nxt[U];
        float px[U][4], pc[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tile = first_tile + u;
            n_cur[u] = load_index(tile);
            n_nxt[u] = -1;
            if (tile < tiles) {
                issue_gather(u, tile, n_cur[u]);
                if (pre_xyz && hb == 0) load_xyz4(tile, n_cur[u], px[u], pc[u]);
                finish_gather(u, tile, n_cur[u], px[u], pc[u], 0u);
                fence_proxy_async();
                arrive(u);
            }
        }

        for (int it = 0; it < n_iter; ++it) {
            const uint32_t par = (uint32_t)it & 1u;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const Tc2Layer &ly = p.ly[l];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int tile = first_tile + u + it * stride;
                    if (tile >= tiles) continue;
                    const uint32_t a_addr = smem_u32(sm.A_base) + (uint32_t)(u * p.a_unit_bytes);
                    const uint32_t tmem_row = tmem_base + (uint32_t)(u * p.cols_per_unit) + ((uint32_t)(q * 32) << 16);
                    const bool more = tile + stride < tiles;
                    if (l == 0) n_nxt[u] = load_index(tile + stride);      // needed after this tile's first epilogue at the earliest
                    mbar_wait(acc_ready + u, (acc_phase >> u) & 1u);
                    acc_phase ^= 1u << u;
                    tc_fence_after();
                    tr.mark(0x400 + l * 2 + u);
                    if (l < L - 1) {
                        // hidden layer: this warp's 32-column blocks hb, hb+4, ...
                        const uint32_t a_rowx = a_addr + (uint32_t)(row * 128 + ((row & 7) << 4));
                        const uint32_t a_rowx_hb = (a_rowx ^ (uint32_t)((hb & 1) << 6)) + (uint32_t)((hb >> 1) * SW128_BLOCK);
                        if (ly.nb32 == 4) {
                            uint32_t r[32];
                            tmem_ld32_nowait(tmem_row + (uint32_t)(32 * hb), r);
                            tmem_ld_wait();
                            epi_store32<0>(r, a_rowx_hb, 0u);
                        } else if (ly.nb32 == 8) {
                            uint32_t r0[32], r1[32];
                            tmem_ld32_nowait(tmem_row + (uint32_t)(32 * hb), r0);
                            tmem_ld32_nowait(tmem_row + (uint32_t)(32 * hb + 128), r1);
                            tmem_ld_wait();
                            epi_store32<0>(r0, a_rowx_hb, 0u);
                            epi_store32<0>(r1, a_rowx_hb, (uint32_t)(2 * SW128_BLOCK));
                        } else epi_hidden_rt(tmem_row, a_rowx, hb, 4, ly.nb32);
                        tc_fence_before();
                        fence_proxy_async();
                        arrive(u);
                        tr.mark(0x500 + l * 2 + u);
                        if (l == 0 && more && pre_xyz && hb == 0) load_xyz4(tile + stride, n_nxt[u], px[u], pc[u]);
                    } else {
                        // last layer (transposed): the operand buffer is free -> next tile's feature rows start flowing
                        if (more) {
                            issue_gather(u, tile + stride, n_nxt[u]);
                            if (L == 1 && pre_xyz && hb == 0) load_xyz4(tile + stride, n_nxt[u], px[u], pc[u]);
                        }
                        const float *bs = sm.bias_s + ly.b_off;
                        const uint32_t *vm = sm.vmask + 8 * u + 4 * par;
                        if (!w_ready) { mbar_wait(sm.w_bar, 0); w_ready = true; }
                        PoolOut o;
                        o.all_valid = (vm[0] & vm[1] & vm[2] & vm[3]) == 0xffffffffu;
                        o.has_bf = p.out_bf16 != nullptr;
                        o.os = p.os; o.ob = p.ob; o.contig = p.ob == (int64_t)S * p.os; o.S = S; o.G = G;
                        const int g0 = tile * gpt;
                        // 16-way split of the (128-channel slice, 32-row block) items of this lane quarter: whole groups
                        // stay in one warp
                        const int n_mh = ly.n_mh;
                        const int bpg = K >= 32 ? K / 32 : 1;                   // 32-row blocks per group
                        int ipw = n_mh; if (ipw < bpg) ipw = bpg;               // items per warp
                        const int item0 = hb * ipw;
                        if (item0 < n_mh * 4) {
                            const int mh = item0 >> 2, blk_lo = item0 & 3, blk_hi = min(4, blk_lo + ipw);
                            for (int mhh = mh; mhh < n_mh && mhh * 4 < item0 + ipw; ++mhh) {
                                const int c = mhh * TC_ROWS + row;
                                o.bv = bs[c];
                                o.c_ok = c < ly.Cout;
                                const int b_lo = (mhh == mh) ? blk_lo : 0;
                                const int b_hi = (mhh == mh) ? blk_hi : min(4, item0 + ipw - mhh * 4);
                                o.gg = g0 + ((b_lo * 32) >> logK);
                                o.po = p.out + p.co_off + c;
                                o.pb = p.out_bf16 + p.co_off + c;
                                pool_slice_any<true>(logK, tmem_row + (uint32_t)(mhh * TC_ROWS), b_lo, b_hi, vm, o);
                            }
                        }
                        tc_fence_before();
                        if (more) {
                            finish_gather(u, tile + stride, n_nxt[u], px[u], pc[u], par ^ 1u);
                            fence_proxy_async();
                            arrive(u);
                        }
                        tr.mark(0x600 + u);
                        n_cur[u] = n_nxt[u];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == T3_WORKERS) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------
static unsigned long long *g_trace = nullptr;
void set_debug_trace(void *buf) { g_trace = (unsigned long long *)buf; }
unsigned long long *get_debug_trace() { return g_trace; }

static int pow2_at_least(int x) { int p = 32; while (p < x) p <<= 1; return p; }

struct Tc2Plan { int U, cols_per_unit, tmem_cols, a_unit_bytes, xyz_off; size_t smem; };

static bool tc2_plan(int K, int C, int D, int order, const TcLayout &t, Tc2Plan *pl, bool split = false)
{
    // features fill whole K = 16 slices of the swizzled blocks, the (<= 16) coordinate channels their own slice
    if (D % 16 != 0 || C < 1 || C > 16 || t.ly[0].K != D + 16) return false;
    if (order == PN2_ORDER_ALL) return false;                       // group_all stacks are wide: pn2_sa_tc.cu
    if (K < 8 || K > TC_ROWS || (K & (K - 1))) return false;
    if (t.L > T2_MAX_L) return false;
    if (t.L < 1 || t.ly[t.L - 1].Npad > 512 || t.ly[t.L - 1].nslab != 1) return false;
    int kmax = D, nmax = 0;                                         // widest feature part of any layer's operand
    for (int l = 0; l < t.L; ++l) {
        if (l < t.L - 1 && (t.ly[l].nslab != 1 || t.ly[l].NS > 256)) return false;
        if (l > 0 && t.ly[l].K > kmax) kmax = t.ly[l].K;
        if (t.ly[l].Npad > nmax) nmax = t.ly[l].Npad;
    }
    pl->cols_per_unit = pow2_at_least(nmax);
    pl->xyz_off = (kmax + 63) / 64 * SW128_BLOCK;                   // swizzled 64-channel blocks, then [2][128][16 B] coordinates
    pl->a_unit_bytes = (pl->xyz_off + 2 * A_LBO) * (split ? 2 : 1);      // split operands: hi image, then lo image
    for (int U = T2_MAX_UNITS; U >= 1; U >>= 1) {
        if (U * pl->cols_per_unit > 512) continue;
        const size_t smem = tc2_smem_bytes(U, pl->a_unit_bytes, t.total_bytes);
        if (smem > (size_t)T2_SMEM_LIMIT) continue;
        pl->U = U; pl->smem = smem; pl->tmem_cols = pow2_at_least(U * pl->cols_per_unit);
        return true;
    }
    return false;
}

int sa_bf16_pipelined_covers(int K, int C, int D, int order, int L, const int *dims)
{
    const TcLayout lay = tc_layout(L, dims);
    Tc2Plan pl;
    if (getenv("PN2_TC_RING_ONLY") != nullptr) return 0;          // every launch then goes to the streaming kernel (no bf16 side io)
    return tc2_plan(K, C, D, order, lay, &pl) ? 1 : 0;
}

// returns -100 when this variant does not cover the configuration (the caller falls back to pn2_sa_tc.cu)
int launch_sa_bf16_pipelined(const SaArgs &a, cudaStream_t st)
{
    const TcLayout lay = tc_layout(a.L, a.dims);
    Tc2Plan pl;
    if (!tc2_plan(a.K, a.C, a.D, a.order, lay, &pl)) return -100;
    Tc2Params p;
    p.points = a.points; p.sb = a.sb; p.sn = a.sn; p.sc = a.sc;
    p.feat = a.feat; p.fb = a.fb; p.fn = a.fn; p.fc = a.fc;
    p.centres = a.centres; p.idx32 = a.idx32;
    p.N = a.N; p.S = a.S; p.K = a.K; p.C = a.C; p.D = a.D; p.order = a.order;
    p.L = lay.L;
    for (int l = 0; l < lay.L; ++l) {
        const TcLayer &y = lay.ly[l];
        const bool last = (l == lay.L - 1);
        Tc2Layer &t = p.ly[l];
        const uint32_t lbo = (uint32_t)(last ? y.Npad : y.NS) * 16u;
        t.w_lo_rel = ((uint32_t)y.w_off >> 4) | ((lbo >> 4) << 16);
        t.w_step = (2u * lbo) >> 4;
        t.bs_lo_rel = ((uint32_t)y.bs_off >> 4) | ((lbo >> 4) << 16);
        t.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((last ? TC_ROWS : y.NS) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
        t.nks = (l == 0 ? a.D : y.K) / 16;
        t.nb32 = y.Npad / 32;
        t.n_mh = y.Npad / TC_ROWS;
        t.Cout = y.Cout; t.b_off = y.b_off;
    }
    p.img_bytes = lay.total_bytes; p.bias_off = lay.w_bytes;
    p.packed = (const unsigned char *)a.packed;
    p.out = a.out; p.ob = a.ob; p.os = a.os; p.co_off = a.co_off;
    p.G = (int64_t)a.B * a.S;
    p.rows = p.G * a.K;
    p.tiles = (p.rows + TC_ROWS - 1) / TC_ROWS;
    if (p.rows >= (1ll << 31) - TC_ROWS) return -100;             // the kernel indexes rows with 32-bit integers
    p.U = pl.U; p.cols_per_unit = pl.cols_per_unit; p.tmem_cols = pl.tmem_cols;
    p.a_unit_bytes = pl.a_unit_bytes; p.xyz_off = pl.xyz_off;
    p.feat_fast = (a.D == 0) || (a.fc == 1 && a.fn % 4 == 0 && a.fb % 4 == 0 && ((uintptr_t)a.feat & 15) == 0);
    // bf16 features of the previous layer: [B,N,D] contiguous, 16-byte rows chunks
    p.feat_bf16 = (a.feat_bf16 && p.feat_fast && a.D > 0 && ((uintptr_t)a.feat_bf16 & 15) == 0) ? (const __nv_bfloat16 *)a.feat_bf16 : nullptr;
    p.out_bf16 = (__nv_bfloat16 *)a.out_bf16;
    p.trace = g_trace;
    p.use_tma = 0; p.feat_rows = a.B * a.N; p.static_sched = 1;
    p.split = 0; p.a_lo_off = 0;
    static const bool no_tma = getenv("PN2_NO_TMA_GATHER") != nullptr;
    if (p.feat_bf16 && a.D % 64 == 0 && (int64_t)a.B * a.N < (1ll << 31) - 1 && !no_tma) {
        if (make_row_gather_map(p.feat_bf16, (int64_t)a.B * a.N, a.D, &p.tmap) == 0) p.use_tma = 1;
    }
    if (p.tiles == 0) return 0;
    // warp-specialised schedule (pn2_sa_tc4.cu): configurations whose layer-0 features are absent or bf16 rows
    // PN2_TC4: -1 (default) = two-unit stacks only (sa2-sized: 2.68 ms against 2.79 ms for the cooperative kernel below),
    // 0 = never, 1 = every stack it covers (four-unit stacks are slower on it: 3.6 ms against 1.46 ms for sa1)
    static const int tc4_env = getenv("PN2_TC4") ? atoi(getenv("PN2_TC4")) : -1;
    if ((tc4_env > 0 || (tc4_env < 0 && pl.U == 2)) && (a.D == 0 || (p.feat_bf16 && a.D % 64 == 0))) return launch_sa_tc4(p, pl.smem, st);
    if (a.D > 0 && !a.feat && !p.use_tma) return -100;           // bf16-only features: the remaining kernels gather them by TMA
    static const int coop_env = getenv("PN2_TC_COOP") ? atoi(getenv("PN2_TC_COOP")) : 1;
    if (coop_env && pl.U == 2 && (a.D == 0 || p.use_tma)) {
        cudaError_t e3 = cudaFuncSetAttribute(sa_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e3 != cudaSuccess) { set_error("sa_tc3: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e3)); return (int)e3; }
        int64_t grid3 = (p.tiles + 1) / 2;
        if (grid3 > 148) grid3 = 148;
        sa_tc3_kernel<<<(unsigned)grid3, T3_THREADS, pl.smem, st>>>(p);
        return launch_status("pn2_sa_fused(bf16, cooperative)");
    }
    int64_t grid = (p.tiles + pl.U - 1) / pl.U;
    if (grid > 148) grid = 148;
    auto run = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) { set_error("sa_tc2: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e)); return (int)e; }
        kern<<<(unsigned)grid, 4 * pl.U * 32, pl.smem, st>>>(p);
        return launch_status("pn2_sa_fused(bf16, pipelined)");
    };
    const int fmode = a.D == 0 ? 0 : (p.use_tma ? 1 : 2);
    const bool pre = a.C <= 4;
    if (fmode == 0) return pre ? run(sa_tc2_kernel<0, true>) : run(sa_tc2_kernel<0, false>);
    if (fmode == 1) return pre ? run(sa_tc2_kernel<1, true>) : run(sa_tc2_kernel<1, false>);
    return pre ? run(sa_tc2_kernel<2, true>) : run(sa_tc2_kernel<2, false>);
}

// fp32 mode (split operands, see pn2_sa_tcs.cu) on the four-unit schedule above: stacks WITHOUT feature channels (sa1 of every
// model: D = 0) whose split weights fit shared memory next to four units.  `packed_split` = the split image of
// pn2_pack_weights(PN2_PREC_F32).  Returns -100 when the configuration is not covered (the caller uses pn2_sa_tcs.cu).
int launch_sa_f32_split_units(const SaArgs &a, const void *packed_split, cudaStream_t st)
{
    static const bool off = getenv("PN2_F32_UNITS") != nullptr && atoi(getenv("PN2_F32_UNITS")) == 0;
    if (off || a.D != 0) return -100;
    const TcLayout lay = tc_layout(a.L, a.dims, true);
    Tc2Plan pl;
    if (!tc2_plan(a.K, a.C, a.D, a.order, lay, &pl, true) || pl.U < 2) return -100;
    Tc2Params p;
    p.points = a.points; p.sb = a.sb; p.sn = a.sn; p.sc = a.sc;
    p.feat = nullptr; p.fb = p.fn = p.fc = 0;
    p.centres = a.centres; p.idx32 = a.idx32;
    p.N = a.N; p.S = a.S; p.K = a.K; p.C = a.C; p.D = 0; p.order = a.order;
    p.L = lay.L;
    const int spc = lay.kblk / 16;
    for (int l = 0; l < lay.L; ++l) {
        const TcLayer &y = lay.ly[l];
        const bool last = (l == lay.L - 1);
        Tc2Layer &t = p.ly[l];
        const uint32_t lbo = (uint32_t)(last ? y.Npad : y.NS) * 16u;
        t.w_lo_rel = ((uint32_t)y.w_off >> 4) | ((lbo >> 4) << 16);
        t.w_step = (2u * lbo) >> 4;
        t.bs_lo_rel = ((uint32_t)y.bs_off >> 4) | ((lbo >> 4) << 16);
        t.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((last ? TC_ROWS : y.NS) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
        t.nks = (l == 0 ? 0 : y.K) / 16;
        t.nb32 = y.Npad / 32;
        t.n_mh = y.Npad / TC_ROWS;
        t.Cout = y.Cout; t.b_off = y.b_off;
        t.s_spc = spc;
        t.s_cbf16 = (uint32_t)tc_chunk_bytes(y, 0, lay.kblk) >> 4;
        t.s_cbl16 = (uint32_t)tc_chunk_bytes(y, y.nkb - 1, lay.kblk) >> 4;
    }
    p.img_bytes = lay.total_bytes; p.bias_off = lay.w_bytes;
    p.packed = (const unsigned char *)packed_split;
    p.out = a.out; p.ob = a.ob; p.os = a.os; p.co_off = a.co_off;
    p.G = (int64_t)a.B * a.S;
    p.rows = p.G * a.K;
    p.tiles = (p.rows + TC_ROWS - 1) / TC_ROWS;
    if (p.rows >= (1ll << 31) - TC_ROWS) return -100;
    p.U = pl.U; p.cols_per_unit = pl.cols_per_unit; p.tmem_cols = pl.tmem_cols;
    p.a_unit_bytes = pl.a_unit_bytes; p.xyz_off = pl.xyz_off;
    p.feat_fast = 1; p.feat_bf16 = nullptr; p.out_bf16 = nullptr;
    p.trace = nullptr; p.use_tma = 0; p.feat_rows = a.B * a.N; p.static_sched = 1;
    p.split = 1; p.a_lo_off = (uint32_t)(pl.a_unit_bytes / 2);
    memset(&p.tmap, 0, sizeof(p.tmap));
    if (p.tiles == 0) return 0;
    int64_t grid = (p.tiles + pl.U - 1) / pl.U;
    if (grid > 148) grid = 148;
    auto run = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess) { set_error("sa_tc2(split): smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e)); return (int)e; }
        kern<<<(unsigned)grid, 4 * pl.U * 32, pl.smem, st>>>(p);
        return launch_status("pn2_sa_fused(f32, split units)");
    };
    return a.C <= 4 ? run(sa_tc2_kernel<0, true, true>) : run(sa_tc2_kernel<0, false, true>);
}

}  // namespace pn2
