// K3 (bf16 mode), pipelined variant -- the fused set-abstraction kernel for stacks whose weights fit in
// shared memory next to the activations (every sa1 / sa2 configuration of the reference's SSG models and most
// MSG branches).  Reference: model/pointnet2_utils.py:166-172 (SSG) and :204-218 (one MSG branch).
//
// One persistent CTA per SM runs U independent "units".  A unit is a warpgroup of 4 warps (128 threads, one
// thread per row of a 128-row tile = one TMEM lane) that walks its tiles through the chain
//     gather -> [MMA l -> epilogue l] x (L-1) -> MMA L-1 -> max-pool -> store
// and owns one activation buffer in shared memory (each layer's epilogue overwrites its own, fully consumed,
// input) and one accumulator region in TMEM.  One extra warp issues every tcgen05.mma of the CTA (one elected
// lane, descriptors on the uniform datapath: ~65 cycles per instruction, the tensor-pipe time of an N = 128
// MMA): it serves the units round-robin, waiting on a unit's "operand ready" mbarrier (4 warp arrivals) and
// committing to its "accumulator ready" mbarrier, so a unit's MMAs run back to back at pipe speed while the
// other units drain TMEM or gather.
// All weights (BN folded, bf16, canonical K-major core-matrix layout, written by pn2_pack_weights) are
// bulk-copied (cp.async.bulk + mbarrier) into shared memory once per CTA.
//
//   epilogue (hidden layer): tcgen05.ld 32x32b.x32 -> +bias (FADD2) -> cvt.rn.relu.bf16x2 -> st.shared.v4
//   last layer + pooling   : the product is issued TRANSPOSED, D^T[channel][row] = W (M operand) x X^T (N operand;
//                            both are K-major in shared memory, so swapping roles costs nothing).  A thread then
//                            owns one output channel (TMEM lane) and the rows of a group are consecutive TMEM
//                            columns: the max over the K neighbours is a chain of FMNMX3 on the registers that
//                            tcgen05.ld returns, bias + ReLU (which commute with the max) are applied once per
//                            group, and a warp stores 32 consecutive channels of a group (128 B coalesced).
//   gather                 : features of the previous set-abstraction layer arrive in bf16 (side output of this
//                            kernel) and go global -> shared with cp.async, 16 B = one k-group of one row,
//                            no registers; fp32 features / raw coordinates go through registers.
#include <cuda.h>
#include <stdlib.h>
#include "pn2_tc.cuh"

namespace pn2 {

constexpr int T2_MAX_UNITS = 4;
constexpr int T2_MAX_THREADS = (4 * T2_MAX_UNITS + 1) * 32;   // 544: unit warps + the MMA-issue warp
constexpr int T2_SMEM_LIMIT = 227 * 1024;

struct Tc2Params {
    const float *points; int64_t sb, sn, sc;
    const float *feat;   int64_t fb, fn, fc;
    const float *centres; const int32_t *idx32;
    int N, S, K, C, D, order;
    TcLayout lay;
    const unsigned char *packed;
    float *out; int64_t ob, os; int co_off;
    int64_t G, rows, tiles;
    int U, cols_per_unit, tmem_cols, a_unit_bytes, xyz_off, feat_fast;   // xyz_off: the no-swizzle coordinate block inside a unit's buffer
    const __nv_bfloat16 *feat_bf16; __nv_bfloat16 *out_bf16;
    unsigned long long *trace;   // debug timeline (pn2_debug_trace), CTA 0 only; nullptr = off
    int unit_issue;              // 1: each unit's first warp issues its own MMAs (small stacks, U = 4); 0: the MMA warp does
    int use_tma, feat_rows;      // bf16 features gathered by TMA (tile::gather4) through `tmap` = [feat_rows = B*N][D]
    alignas(64) CUtensorMap tmap;
};

__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long x, y, z;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(z));
    return r;
}

// hidden-layer epilogue of 32 accumulator columns [c0, c0+32) of this thread's row
__device__ __forceinline__ void epi_block(const uint32_t (&r)[32], const float *bs, unsigned char *A, int c0, int row)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 b0 = *reinterpret_cast<const float4 *>(bs + c0 + 8 * i);
        const float4 b1 = *reinterpret_cast<const float4 *>(bs + c0 + 8 * i + 4);
        const float2 v0 = add2(make_float2(__uint_as_float(r[8 * i + 0]), __uint_as_float(r[8 * i + 1])), make_float2(b0.x, b0.y));
        const float2 v1 = add2(make_float2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])), make_float2(b0.z, b0.w));
        const float2 v2 = add2(make_float2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])), make_float2(b1.x, b1.y));
        const float2 v3 = add2(make_float2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])), make_float2(b1.z, b1.w));
        *reinterpret_cast<uint4 *>(A + sw128_chunk_off(row, c0 / 8 + i)) =
            make_uint4(relu_pack_bf16x2(v0.x, v0.y), relu_pack_bf16x2(v1.x, v1.y), relu_pack_bf16x2(v2.x, v2.y), relu_pack_bf16x2(v3.x, v3.y));
    }
}

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// max of r[OFF .. OFF+CNT) (CNT = 8, 16 or 32) as a tree of 3-input maxima
template <int OFF, int CNT>
__device__ __forceinline__ float max_range(const uint32_t (&r)[32])
{
    float m = __uint_as_float(r[OFF]);
#pragma unroll
    for (int j = 1; j + 1 < CNT; j += 2) m = max3(m, __uint_as_float(r[OFF + j]), __uint_as_float(r[OFF + j + 1]));
    return fmaxf(m, __uint_as_float(r[OFF + CNT - 1]));
}
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool pred)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(pred ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

constexpr int T2_MAX_L = 4;
constexpr int T2_TRACE_SLOTS = 1024;     // per role (units 0..3, MMA warp = 4)

// debug timeline: (event << 48) | clock, appended to this role's slice of the trace buffer
struct Tracer {
    unsigned long long *buf; int n;
    __device__ __forceinline__ void init(unsigned long long *base, int role, bool on) { buf = (on && base) ? base + (size_t)role * T2_TRACE_SLOTS : nullptr; n = 0; }
    __device__ __forceinline__ void mark(int ev)
    {
        if (buf && n < T2_TRACE_SLOTS) buf[n++] = ((unsigned long long)ev << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    }
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t lo, uint32_t hi)
{
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// low / high words of a K-major no-swizzle descriptor (see smem_desc): stepping the start address by X bytes
// is `lo += X >> 4`
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1

__global__ void __launch_bounds__(T2_MAX_THREADS, 1) sa_tc2_kernel(const __grid_constant__ Tc2Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const TcLayout &lay = p.lay;
    const int U = p.U;
    // activation buffers first: their 128-byte-swizzled blocks need 1024-byte alignment
    unsigned char *A_base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char *Wsm = A_base + (size_t)U * p.a_unit_bytes;
    float *bias_s = reinterpret_cast<float *>(Wsm + lay.w_bytes);
    uint32_t *vmask_base = reinterpret_cast<uint32_t *>(bias_s + lay.bias_floats);   // [U][2][4] valid-row ballots, by tile parity
    uint64_t *acc_ready = reinterpret_cast<uint64_t *>(vmask_base + 8 * T2_MAX_UNITS);
    uint64_t *a_ready = acc_ready + T2_MAX_UNITS;        // per unit: operand written / TMEM drained (one arrival per warp)
    uint64_t *feat_ready = a_ready + T2_MAX_UNITS;     // per unit: the TMA gather of the next tile's feature rows has landed
    uint64_t *w_bar = feat_ready + T2_MAX_UNITS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_bar + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (uniform datapath)
    const int L = lay.L;

    if (tid == 0) {
        for (int u = 0; u < T2_MAX_UNITS; ++u) { mbar_init(acc_ready + u, 1); mbar_init(a_ready + u, 4); mbar_init(feat_ready + u, 1); }
        mbar_init(w_bar, 1);
        fence_mbar_init();
        // all weights, once per CTA, through the bulk-copy engine
        mbar_expect_tx(w_bar, (uint32_t)lay.w_bytes);
        for (int off = 0; off < lay.w_bytes; off += 32768) {
            const int nb = min(32768, lay.w_bytes - off);
            bulk_g2s(Wsm + off, p.packed + off, (uint32_t)nb, w_bar);
        }
    }
    const int mma_warp = 4 * U;
    if (warp == mma_warp) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    {
        const float *bsrc = reinterpret_cast<const float *>(p.packed + lay.w_bytes);
        for (int i = tid; i < lay.bias_floats; i += blockDim.x) bias_s[i] = __ldg(bsrc + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles = (int)p.tiles;
    const int stride = (int)gridDim.x * U;

    if (warp == mma_warp) {
      if (!p.unit_issue) {
        // =================================================================================================
        // MMA warp: every tcgen05.mma of the CTA, units served round-robin.  The whole warp runs the loop
        // (warp-uniform values -> uniform datapath), one elected lane issues.
        // =================================================================================================
        int n_max = 0;
        for (int u = 0; u < U; ++u) {
            const int gu = (int)blockIdx.x * U + u;
            n_max = max(n_max, gu < tiles ? (tiles - gu + stride - 1) / stride : 0);
        }
        const int ks_xyz = p.D / 16;
        mbar_wait(w_bar, 0);
        uint32_t phases = 0;
        for (int it = 0; it < n_max; ++it) {
#pragma unroll
            for (int l = 0; l < T2_MAX_L; ++l) {
                if (l >= L) break;
                const TcLayer &ly = lay.ly[l];
                const bool last = (l == L - 1);
                const uint32_t lbo = (uint32_t)(last ? ly.Npad : ly.NS) * 16u;
                const uint32_t w_lo = desc_lo(smem_u32(Wsm) + (uint32_t)ly.w_off, lbo);
                const uint32_t w_step = (2u * lbo) >> 4;
                const uint32_t idesc = instr_desc_bf16(last ? TC_ROWS : ly.NS);
                const int nks = ly.K / 16;
                const int n_mh = last ? ly.Npad / TC_ROWS : 1;
                for (int u = 0; u < U; ++u) {
                    if ((int)blockIdx.x * U + u + it * stride >= tiles) continue;
                    mbar_wait(a_ready + u, (phases >> u) & 1u);
                    phases ^= 1u << u;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(A_base) + (uint32_t)(u * p.a_unit_bytes);
                    const uint32_t a_sw = desc_lo_sw128(a_addr);
                    const uint32_t a_xyz = desc_lo(a_addr + (uint32_t)p.xyz_off, A_LBO);
                    const uint32_t d_tm = tmem_base + (uint32_t)(u * p.cols_per_unit);
                    if (elect_one()) {
                        for (int mh = 0; mh < n_mh; ++mh)
                            for (int ks = 0; ks < nks; ++ks) {
                                const bool xyz = (l == 0 && ks == ks_xyz);
                                const uint64_t act = xyz ? mk_desc(a_xyz, DESC_HI)
                                                         : mk_desc(a_sw + (uint32_t)((ks >> 2) * (SW128_BLOCK / 16) + (ks & 3) * 2), DESC_HI_SW128);
                                const uint64_t wgt = mk_desc(w_lo + (uint32_t)(mh * TC_ROWS) + (uint32_t)ks * w_step, DESC_HI);
                                // hidden layers: D[row][channel] = X (M operand) x W^T; last layer, transposed:
                                // D[channel][row] = W (128 output channels as M) x X^T (the tile's 128 rows as N)
                                if (!last) umma_bf16(d_tm, act, wgt, idesc, ks ? 1u : 0u);
                                else umma_bf16(d_tm + (uint32_t)(mh * TC_ROWS), wgt, act, idesc, ks ? 1u : 0u);
                            }
                        umma_commit(acc_ready + u);
                    }
                    __syncwarp();
                }
            }
        }
      }
    } else {
        const int u = warp >> 2, q = warp & 3, row = q * 32 + lane;
        unsigned char *A = A_base + (size_t)u * p.a_unit_bytes;
        unsigned char *X_row = A + p.xyz_off + row * 16;            // this row in the coordinate block [2][128][16 B]
        uint32_t *vmask_u = vmask_base + 8 * u;
        uint32_t parity = 0;
        const uint32_t tmem_row = tmem_base + (uint32_t)(u * p.cols_per_unit) + ((uint32_t)(q * 32) << 16);
        const int K = p.K, C = p.C, D = p.D, S = p.S, N = p.N;
        const int logK = 31 - __clz(K);
        const int gpt = TC_ROWS >> logK;
        const int rows = (int)p.rows, G = (int)p.G;
        uint32_t acc_phase = 0;
        const int gu = (int)blockIdx.x * U + u;
        uint64_t *my_acc_ready = acc_ready + u;
        const bool twin = p.feat_fast && p.feat_bf16 != nullptr;    // bf16 features: cp.async straight into the operand
        const bool pre_xyz = C <= 4;                                // coordinates prefetched one tile ahead

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
        // Feature rows -> operand buffer.  Lanes are mapped (4 rows) x (8 chunks of 16 B): eight lanes read 128
        // contiguous bytes of one source row (bf16) and write one 128-byte swizzled line, so the copy is coalesced
        // in global memory and conflict free in shared memory.  The row's neighbour index lives in the lane that
        // owns the row (thread = row everywhere else) and is fetched with a shuffle.
        // bf16 features: asynchronous (cp.async), the buffer must be free when this is called.
        uint64_t *my_feat_ready = feat_ready + u;
        uint32_t feat_phase = 0;
        auto issue_twin = [&](int tile, int n_mine) {
            if (p.use_tma) {
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
        uint64_t *my_a_ready = a_ready + u;
        const uint32_t a_sw = desc_lo_sw128(smem_u32(A));
        const uint32_t a_xyz = desc_lo(smem_u32(A) + (uint32_t)p.xyz_off, A_LBO);
        const uint32_t d_tm = tmem_base + (uint32_t)(u * p.cols_per_unit);
        const int ks_xyz = D / 16;
        bool w_ready = false;
        // operand written (and proxy-fenced) / TMEM drained by this warp.  Either tell the MMA warp, or (small
        // stacks) meet the unit's other warps at a named barrier and let this unit's first warp issue the layer.
        auto issue_layer = [&](int l) {
            if (!p.unit_issue) {
                __syncwarp();
                if (lane == 0) mbar_arrive(my_a_ready);
                return;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(u + 1) : "memory");
            if (q == 0) {
                if (!w_ready) { mbar_wait(w_bar, 0); w_ready = true; }
                tc_fence_after();
                const TcLayer &ly = lay.ly[l];
                const bool last = (l == L - 1);
                const uint32_t lbo = (uint32_t)(last ? ly.Npad : ly.NS) * 16u;
                const uint32_t w_lo = desc_lo(smem_u32(Wsm) + (uint32_t)ly.w_off, lbo);
                const uint32_t w_step = (2u * lbo) >> 4;
                const uint32_t idesc = instr_desc_bf16(last ? TC_ROWS : ly.NS);
                const int nks = ly.K / 16;
                if (elect_one()) {
                    const int n_mh = last ? ly.Npad / TC_ROWS : 1;
                    for (int mh = 0; mh < n_mh; ++mh)
                        for (int ks = 0; ks < nks; ++ks) {
                            const bool xyz = (l == 0 && ks == ks_xyz);
                            const uint64_t act = xyz ? mk_desc(a_xyz, DESC_HI)
                                                     : mk_desc(a_sw + (uint32_t)((ks >> 2) * (SW128_BLOCK / 16) + (ks & 3) * 2), DESC_HI_SW128);
                            const uint64_t wgt = mk_desc(w_lo + (uint32_t)(mh * TC_ROWS) + (uint32_t)ks * w_step, DESC_HI);
                            if (!last) umma_bf16(d_tm, act, wgt, idesc, ks ? 1u : 0u);
                            else umma_bf16(d_tm + (uint32_t)(mh * TC_ROWS), wgt, act, idesc, ks ? 1u : 0u);
                        }
                    umma_commit(my_acc_ready);
                }
                __syncwarp();
            }
        };
        int n_cur = load_index(gu);
        float px[4], pc[4];
        if (gu < tiles) {
            if (twin) issue_twin(gu, n_cur);
            if (pre_xyz) load_xyz4(gu, n_cur, px, pc);
        }

        for (int tile = gu; tile < tiles; tile += stride) {
            uint32_t *vmask_s = vmask_u + 4 * parity;    // a warp can run at most one tile ahead of its unit
            parity ^= 1u;
            const int n = n_cur;
            const int n_next = load_index(tile + stride);          // used after the first epilogue at the earliest
            const bool valid = is_valid(n);
            const int g = (tile * TC_ROWS + row) >> logK;      // group of this row
            const int b = (int)((unsigned)g / (unsigned)S);

            // ---- gather: features (swizzled blocks) were started earlier (bf16 twin) or are converted here (fp32);
            //      coordinates minus centre go to the unit's coordinate block: channels [D, D+16) of layer 0 ----------
            if (D) {
                if (p.feat_fast) { if (!twin) gather_f32(tile, n); }
                else {
                    for (int k = 0; k < D; ++k) {
                        const float t = valid ? __ldg(p.feat + (int64_t)b * p.fb + (int64_t)n * p.fn + (int64_t)k * p.fc) : 0.f;
                        *reinterpret_cast<__nv_bfloat16 *>(A + sw128_chunk_off(row, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(t);
                    }
                }
            }
            if (pre_xyz) {
                *reinterpret_cast<uint4 *>(X_row) = make_uint4(pack_bf16x2(__fsub_rn(px[0], pc[0]), __fsub_rn(px[1], pc[1])),
                                                               pack_bf16x2(__fsub_rn(px[2], pc[2]), __fsub_rn(px[3], pc[3])), 0u, 0u);
                *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(0u, 0u, 0u, 0u);
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
                *reinterpret_cast<uint4 *>(X_row) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15]));
            }
            if (twin) {
                if (p.use_tma) { mbar_wait(my_feat_ready, feat_phase); feat_phase ^= 1u; }
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (lane == 0) vmask_s[q] = vm;      // read by the whole unit in the pooling step (ordered by the mbarriers)
            fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core (async proxy)
            tr.mark(0x300);
            issue_layer(0);
            tr.mark(0x700);
            const bool more = tile + stride < tiles;

#pragma unroll
            for (int l = 0; l < T2_MAX_L; ++l) {
                if (l >= L) break;
                const TcLayer &ly = lay.ly[l];
                mbar_wait(my_acc_ready, acc_phase);
                acc_phase ^= 1u;
                tc_fence_after();
                tr.mark(0x400 + l);
                const float *bs = bias_s + ly.b_off;
                if (l < L - 1) {
                    // ---- hidden layer: accumulator -> +bias -> ReLU -> bf16 -> the next layer's A operand (in place) --
#pragma unroll 1
                    for (int c0 = 0; c0 < ly.Npad; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32_nowait(tmem_row + (uint32_t)c0, r);
                        tmem_ld_wait();
                        epi_block(r, bs, A, c0, row);
                    }
                    tc_fence_before();           // TMEM reads ordered before the next MMA's write to this region
                    fence_proxy_async();
                    tr.mark(0x500 + l);
                    issue_layer(l + 1);
                    tr.mark(0x700 + l + 1);
                    // the next tile's index has arrived by now: start its coordinate loads (they land during the
                    // remaining layers)
                    if (l == 0 && L > 1 && more && pre_xyz) load_xyz4(tile + stride, n_next, px, pc);
                } else {
                    // ---- last layer (transposed): this thread = channel mh*128 + row; columns = the tile's rows ------
                    // every MMA that reads the operand buffer has completed: the next tile's feature rows can start
                    // flowing in underneath the pooling
                    if (more) {
                        if (twin) issue_twin(tile + stride, n_next);
                        if (L == 1 && pre_xyz) load_xyz4(tile + stride, n_next, px, pc);
                    }
                    tr.mark(0x610);
                    const uint32_t m0 = vmask_s[0], m1 = vmask_s[1], m2 = vmask_s[2], m3 = vmask_s[3];
                    const bool all_valid = (m0 & m1 & m2 & m3) == 0xffffffffu;
                    const int g0 = tile * gpt;
                    const int b0 = (int)((unsigned)g0 / (unsigned)S);
                    const int s0 = g0 - b0 * S;
                    const int64_t wrap = p.ob - (int64_t)S * p.os;      // offset correction when a group index rolls into the next cloud
                    for (int mh = 0; mh < ly.Npad / TC_ROWS; ++mh) {
                        const int c = mh * TC_ROWS + row;
                        const float bv = bs[c];
                        const bool c_ok = c < ly.Cout;
                        int64_t o = (int64_t)b0 * p.ob + (int64_t)s0 * p.os + p.co_off + c;   // output element of (group, channel)
                        int ss = s0, gg = g0;
                        // group finished: max over its rows -> + bias -> ReLU -> [B,S,Ctot]; then step to the next group
                        auto emit = [&](float m) {
                            if (c_ok && gg < G && (all_valid || m > -INFINITY)) {
                                const float y = fmaxf(m + bv, 0.f);
                                p.out[o] = y;
                                if (p.out_bf16) p.out_bf16[o] = __float2bfloat16_rn(y);
                            }
                            ++gg; o += p.os;
                            if (++ss == S) { ss = 0; o += wrap; }
                        };
                        float run = -INFINITY;
#pragma unroll 1
                        for (int blk = 0; blk < 4; ++blk) {
                            uint32_t r[32];
                            tmem_ld32_nowait(tmem_row + (uint32_t)(mh * TC_ROWS + blk * 32), r);
                            tmem_ld_wait();
                            tr.mark(0x620 + mh * 4 + blk);
                            if (!all_valid) {        // padding rows / empty balls (never on the reference's configurations)
                                const uint32_t mk = blk == 0 ? m0 : blk == 1 ? m1 : blk == 2 ? m2 : m3;
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (!((mk >> j) & 1u)) r[j] = 0xff800000u;   // -inf
                            }
                            if (K >= 32) {
                                run = fmaxf(run, max_range<0, 32>(r));
                                if ((((blk + 1) * 32) & (K - 1)) == 0) { emit(run); run = -INFINITY; }
                            } else if (K == 16) {
                                emit(max_range<0, 16>(r));
                                emit(max_range<16, 16>(r));
                            } else {
                                emit(max_range<0, 8>(r));
                                emit(max_range<8, 8>(r));
                                emit(max_range<16, 8>(r));
                                emit(max_range<24, 8>(r));
                            }
                            tr.mark(0x630 + mh * 4 + blk);
                        }
                    }
                    tc_fence_before();
                    tr.mark(0x600);
                }
            }
            n_cur = n_next;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == mma_warp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ======================================================================================================
// Cooperative variant, for stacks whose weights leave room for only TWO tiles per SM (sa2-sized).
//
// With two units of four warps each unit's chain (epilogue, pooling: a few hundred dependent instructions per
// warp, ~8 cycles each) is far longer than its MMAs, and the two chains tend to run in phase, so the tensor
// pipe idles.  Here all 16 worker warps serve BOTH tiles: the CTA walks a fixed schedule
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
    const TcLayout &lay = p.lay;
    constexpr int U = 2;
    unsigned char *A_base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    unsigned char *Wsm = A_base + (size_t)U * p.a_unit_bytes;
    float *bias_s = reinterpret_cast<float *>(Wsm + lay.w_bytes);
    uint32_t *vmask_base = reinterpret_cast<uint32_t *>(bias_s + lay.bias_floats);   // [U][2][4] valid-row ballots, by tile parity
    uint64_t *acc_ready = reinterpret_cast<uint64_t *>(vmask_base + 8 * T2_MAX_UNITS);
    uint64_t *a_ready = acc_ready + T2_MAX_UNITS;        // per unit: 16 warp arrivals (+ the TMA bytes of a gather)
    uint64_t *w_bar = a_ready + 2 * T2_MAX_UNITS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(w_bar + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = lay.L;

    if (tid == 0) {
        for (int u = 0; u < U; ++u) { mbar_init(acc_ready + u, 1); mbar_init(a_ready + u, T3_WORKERS); }
        mbar_init(w_bar, 1);
        fence_mbar_init();
        mbar_expect_tx(w_bar, (uint32_t)lay.w_bytes);
        for (int off = 0; off < lay.w_bytes; off += 32768) {
            const int nb = min(32768, lay.w_bytes - off);
            bulk_g2s(Wsm + off, p.packed + off, (uint32_t)nb, w_bar);
        }
    }
    if (warp == T3_WORKERS) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    {
        const float *bsrc = reinterpret_cast<const float *>(p.packed + lay.w_bytes);
        for (int i = tid; i < lay.bias_floats; i += blockDim.x) bias_s[i] = __ldg(bsrc + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles = (int)p.tiles;
    const int stride = (int)gridDim.x * U;
    const int first_tile = (int)blockIdx.x * U;
    const int n_iter = first_tile < tiles ? (tiles - first_tile + stride - 1) / stride : 0;   // trips of unit 0 (>= unit 1's)

    if (warp == T3_WORKERS) {
        // ---- MMA warp: same schedule as the workers, one elected lane issues ---------------------------------
        const int ks_xyz = p.D / 16;
        mbar_wait(w_bar, 0);
        uint32_t phases = 0;
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll
            for (int l = 0; l < T2_MAX_L; ++l) {
                if (l >= L) break;
                const TcLayer &ly = lay.ly[l];
                const bool last = (l == L - 1);
                const uint32_t lbo = (uint32_t)(last ? ly.Npad : ly.NS) * 16u;
                const uint32_t w_lo = desc_lo(smem_u32(Wsm) + (uint32_t)ly.w_off, lbo);
                const uint32_t w_step = (2u * lbo) >> 4;
                const uint32_t idesc = instr_desc_bf16(last ? TC_ROWS : ly.NS);
                const int nks = ly.K / 16;
                const int n_mh = last ? ly.Npad / TC_ROWS : 1;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (first_tile + u + it * stride >= tiles) continue;
                    mbar_wait(a_ready + u, (phases >> u) & 1u);
                    phases ^= 1u << u;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(A_base) + (uint32_t)(u * p.a_unit_bytes);
                    const uint32_t a_sw = desc_lo_sw128(a_addr);
                    const uint32_t a_xyz = desc_lo(a_addr + (uint32_t)p.xyz_off, A_LBO);
                    const uint32_t d_tm = tmem_base + (uint32_t)(u * p.cols_per_unit);
                    if (elect_one()) {
                        for (int mh = 0; mh < n_mh; ++mh)
                            for (int ks = 0; ks < nks; ++ks) {
                                const bool xyz = (l == 0 && ks == ks_xyz);
                                const uint64_t act = xyz ? mk_desc(a_xyz, DESC_HI)
                                                         : mk_desc(a_sw + (uint32_t)((ks >> 2) * (SW128_BLOCK / 16) + (ks & 3) * 2), DESC_HI_SW128);
                                const uint64_t wgt = mk_desc(w_lo + (uint32_t)(mh * TC_ROWS) + (uint32_t)ks * w_step, DESC_HI);
                                if (!last) umma_bf16(d_tm, act, wgt, idesc, ks ? 1u : 0u);
                                else umma_bf16(d_tm + (uint32_t)(mh * TC_ROWS), wgt, act, idesc, ks ? 1u : 0u);
                            }
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
            unsigned char *A = A_base + (size_t)u * p.a_unit_bytes;
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
            unsigned char *X_row = A_base + (size_t)u * p.a_unit_bytes + p.xyz_off + row * 16;
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
            if (lane == 0) vmask_base[8 * u + 4 * par + q] = vm;
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
#pragma unroll
            for (int l = 0; l < T2_MAX_L; ++l) {
                if (l >= L) break;
                const TcLayer &ly = lay.ly[l];
                const float *bs = bias_s + ly.b_off;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int tile = first_tile + u + it * stride;
                    if (tile >= tiles) continue;
                    unsigned char *A = A_base + (size_t)u * p.a_unit_bytes;
                    const uint32_t tmem_row = tmem_base + (uint32_t)(u * p.cols_per_unit) + ((uint32_t)(q * 32) << 16);
                    const bool more = tile + stride < tiles;
                    if (l == 0) n_nxt[u] = load_index(tile + stride);      // needed after this tile's first epilogue at the earliest
                    mbar_wait(acc_ready + u, (acc_phase >> u) & 1u);
                    acc_phase ^= 1u << u;
                    tc_fence_after();
                    tr.mark(0x400 + l * 2 + u);
                    if (l < L - 1) {
                        // hidden layer: this warp's 32-column blocks hb, hb+4, ...
#pragma unroll 1
                        for (int c0 = 32 * hb; c0 < ly.Npad; c0 += 128) {
                            uint32_t r[32];
                            tmem_ld32_nowait(tmem_row + (uint32_t)c0, r);
                            tmem_ld_wait();
                            epi_block(r, bs, A, c0, row);
                        }
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
                        const uint32_t *vm = vmask_base + 8 * u + 4 * par;
                        const bool all_valid = (vm[0] & vm[1] & vm[2] & vm[3]) == 0xffffffffu;
                        const int g0 = tile * gpt;
                        const int64_t wrap = p.ob - (int64_t)S * p.os;
                        // 16-way split of the (128-channel slice, 32-row block) items of this lane quarter: whole groups
                        // stay in one warp
                        const int n_mh = ly.Npad / TC_ROWS;
                        const int bpg = K >= 32 ? K / 32 : 1;                   // 32-row blocks per group
                        int ipw = (n_mh * 4) / 4; if (ipw < bpg) ipw = bpg;     // items per warp
                        const int item0 = hb * ipw;
                        if (item0 < n_mh * 4) {
                            const int mh = item0 >> 2, blk_lo = item0 & 3, blk_hi = min(4, blk_lo + ipw);
                            for (int mhh = mh; mhh < n_mh && mhh * 4 < item0 + ipw; ++mhh) {
                                const int c = mhh * TC_ROWS + row;
                                const float bv = bs[c];
                                const bool c_ok = c < ly.Cout;
                                const int b_lo = (mhh == mh) ? blk_lo : 0;
                                const int b_hi = (mhh == mh) ? blk_hi : min(4, item0 + ipw - mhh * 4);
                                int gg = g0 + ((b_lo * 32) >> logK);
                                const int bb = (int)((unsigned)gg / (unsigned)S);
                                int ss = gg - bb * S;
                                int64_t o = (int64_t)bb * p.ob + (int64_t)ss * p.os + p.co_off + c;
                                auto emit = [&](float m) {
                                    if (c_ok && gg < G && (all_valid || m > -INFINITY)) {
                                        const float y = fmaxf(m + bv, 0.f);
                                        p.out[o] = y;
                                        if (p.out_bf16) p.out_bf16[o] = __float2bfloat16_rn(y);
                                    }
                                    ++gg; o += p.os;
                                    if (++ss == S) { ss = 0; o += wrap; }
                                };
                                float run = -INFINITY;
#pragma unroll 1
                                for (int blk = b_lo; blk < b_hi; ++blk) {
                                    uint32_t r[32];
                                    tmem_ld32_nowait(tmem_row + (uint32_t)(mhh * TC_ROWS + blk * 32), r);
                                    tmem_ld_wait();
                                    if (!all_valid) {
                                        const uint32_t mk = vm[blk];
#pragma unroll
                                        for (int j = 0; j < 32; ++j) if (!((mk >> j) & 1u)) r[j] = 0xff800000u;
                                    }
                                    if (K >= 32) {
                                        run = fmaxf(run, max_range<0, 32>(r));
                                        if ((((blk + 1) * 32) & (K - 1)) == 0) { emit(run); run = -INFINITY; }
                                    } else if (K == 16) {
                                        emit(max_range<0, 16>(r));
                                        emit(max_range<16, 16>(r));
                                    } else {
                                        emit(max_range<0, 8>(r));
                                        emit(max_range<8, 8>(r));
                                        emit(max_range<16, 8>(r));
                                        emit(max_range<24, 8>(r));
                                    }
                                }
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

static int pow2_at_least(int x) { int p = 32; while (p < x) p <<= 1; return p; }

struct Tc2Plan { int U, cols_per_unit, tmem_cols, a_unit_bytes, xyz_off; size_t smem; };

static bool tc2_plan(int K, int C, int D, int order, const TcLayout &t, Tc2Plan *pl)
{
    // features fill whole K = 16 slices of the swizzled blocks, the (<= 16) coordinate channels their own slice
    if (D % 16 != 0 || C < 1 || C > 16 || t.ly[0].K != D + 16) return false;
    if (order == PN2_ORDER_ALL) return false;                       // group_all stacks are wide: pn2_sa_tc.cu
    if (K < 8 || K > TC_ROWS || (K & (K - 1))) return false;
    if (t.L > 4) return false;
    if (t.L < 1 || t.ly[t.L - 1].Npad > 512) return false;
    int kmax = D, nmax = 0;                                         // widest feature part of any layer's operand
    for (int l = 0; l < t.L; ++l) {
        if (l < t.L - 1 && (t.ly[l].nslab != 1 || t.ly[l].NS > 256)) return false;
        if (l > 0 && t.ly[l].K > kmax) kmax = t.ly[l].K;
        if (t.ly[l].Npad > nmax) nmax = t.ly[l].Npad;
    }
    pl->cols_per_unit = pow2_at_least(nmax);
    pl->xyz_off = (kmax + 63) / 64 * SW128_BLOCK;                   // swizzled 64-channel blocks, then [2][128][16 B] coordinates
    pl->a_unit_bytes = pl->xyz_off + 2 * A_LBO;
    for (int U = T2_MAX_UNITS; U >= 1; U >>= 1) {
        if (U * pl->cols_per_unit > 512) continue;
        const size_t smem = (size_t)t.w_bytes + (size_t)t.bias_floats * 4 + (size_t)U * pl->a_unit_bytes +
                            8 * T2_MAX_UNITS * 4 + (3 * T2_MAX_UNITS + 1) * 8 + 16 + 1024;   // + alignment slack
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
    p.lay = lay;
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
    static const int issue_env = getenv("PN2_UNIT_ISSUE") ? atoi(getenv("PN2_UNIT_ISSUE")) : -1;
    p.unit_issue = issue_env >= 0 ? issue_env : (pl.U == T2_MAX_UNITS);   // measured: profiles/README.md
    p.use_tma = 0; p.feat_rows = a.B * a.N;
    static const bool no_tma = getenv("PN2_NO_TMA_GATHER") != nullptr;
    if (p.feat_bf16 && a.D % 64 == 0 && (int64_t)a.B * a.N < (1ll << 31) - 1 && !no_tma) {
        if (make_row_gather_map(p.feat_bf16, (int64_t)a.B * a.N, a.D, &p.tmap) == 0) p.use_tma = 1;
    }
    if (p.tiles == 0) return 0;
    static const int coop_env = getenv("PN2_TC_COOP") ? atoi(getenv("PN2_TC_COOP")) : 1;
    if (coop_env && pl.U == 2 && (a.D == 0 || p.use_tma)) {
        cudaError_t e3 = cudaFuncSetAttribute(sa_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e3 != cudaSuccess) { set_error("sa_tc3: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e3)); return (int)e3; }
        int64_t grid3 = (p.tiles + 1) / 2;
        if (grid3 > 148) grid3 = 148;
        sa_tc3_kernel<<<(unsigned)grid3, T3_THREADS, pl.smem, st>>>(p);
        return launch_status("pn2_sa_fused(bf16, cooperative)");
    }
    cudaError_t e = cudaFuncSetAttribute(sa_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) { set_error("sa_tc2: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e)); return (int)e; }
    int64_t grid = (p.tiles + pl.U - 1) / pl.U;
    if (grid > 148) grid = 148;
    sa_tc2_kernel<<<(unsigned)grid, (4 * pl.U + 1) * 32, pl.smem, st>>>(p);
    return launch_status("pn2_sa_fused(bf16, pipelined)");
}

}  // namespace pn2
