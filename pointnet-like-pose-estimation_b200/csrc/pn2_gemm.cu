// Row GEMM on the 5th-generation tensor cores:  C[M,N] = A[M,K] . W[N,K]^T  for TALL operands (M in the hundreds of
// thousands to millions of point / grouped rows, K and N a few hundred channels) -- the 1x1 convolutions of the path when
// they are NOT fused into the set-abstraction kernel:
//   * the per-point product P = feat (s Wf)^T of the factorised first layer of an MSG branch (pointnet2_utils.py:208-217
//     evaluated once per point instead of once per grouped row): bf16 operands (fp32 rows converted on load), bf16 rows out;
//   * the training-mode convolutions on the grouped rows and their input gradients (pointnet2_utils.py:168-170 / :214-217
//     under train_rotation.py:121-124): fp32 operands, kind::tf32 (the reference's cuDNN policy), fp32 out.
// These products are HBM-bound (2 K N / (K + N) flops per byte moved is far below the tensor roof), so the kernel is built
// around the streams, not the MMA: a persistent CTA keeps its slice of W (<= 256 output columns) resident in shared memory as
// 128-byte-swizzled K-major blocks; 128-row blocks of A arrive through a 4-stage ring by TMA (cp.async.bulk.tensor, 128-byte
// swizzle = the MMA operand image, rows / columns outside the tensor zero-filled by the copy engine; fp32 rows that must
// become bf16 go through 8 loader warps instead: coalesced 16-byte loads -> convert -> swizzled st.shared -> fence); one
// warp issues tcgen05.mma into one of two TMEM accumulators, and four epilogue warps drain the other one: tcgen05.ld ->
// convert -> a swizzled 16 KB staging tile -> TMA store (whole 128-byte lines, clipped at the tensor edge).  The loads of
// tile t+1, the MMAs of tile t and the stores of tile t-1 overlap.
#include <cuda.h>
#include <string.h>
#include "pn2_tc.cuh"

namespace pn2 {

namespace {

constexpr int GM_EPI_WARPS = 4;
constexpr int GM_LOAD_WARPS = 8;
constexpr int GM_MMA_WARP = GM_EPI_WARPS;
constexpr int GM_LOAD_WARP0 = GM_EPI_WARPS + 1;
constexpr int GM_THREADS = (GM_EPI_WARPS + 1 + GM_LOAD_WARPS) * 32;      // 416
constexpr int GM_LOADERS = GM_LOAD_WARPS * 32;                           // 256
constexpr int GM_STAGES = 4;
constexpr int GM_BLOCK = TC_ROWS * 128;                                  // one A block: 128 rows x 128 bytes

struct GemmParams {
    const void *A; int64_t lda;
    const void *W; int64_t ldw;
    void *C; int64_t ldc;
    int64_t M; int N, K;
    int a_f32;          // A elements are fp32 (else bf16)
    int tf32;           // kind::tf32 on fp32 operands (else kind::f16 on bf16 operands, fp32 A converted on load)
    int c_bf16;         // C elements are bf16 (else fp32)
    int Nt;             // output columns per CTA (multiple of 16, <= 256)
    int nkb;            // K blocks of 128 bytes (64 bf16 / 32 tf32 elements)
    int64_t tiles_m;
    int tmem_cols;
    alignas(64) CUtensorMap tmA;      // A as [M][K], box 128 bytes x 128 rows (unused by the converting loaders)
    alignas(64) CUtensorMap tmC;      // C as [M][N], box 128 bytes x 128 rows
};

__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, int c0, int c1, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 3, 128;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor: D = f32, A = B = tf32 (format 2), both K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ uint4 ldg_nc16(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// 16 bytes of the operand image for elements [k0, k0 + cpe) of one source row (row == nullptr or k0 >= K: zeros)
template <bool SRC_F32, bool TO_BF16>
__device__ __forceinline__ uint4 load_chunk(const void *row, int k0, int K)
{
    uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (row == nullptr || k0 >= K) return z;
    if constexpr (SRC_F32 && TO_BF16) {            // 8 fp32 -> 8 bf16
        const uint4 a = ldg_nc16(reinterpret_cast<const float *>(row) + k0);
        const uint4 b = ldg_nc16(reinterpret_cast<const float *>(row) + k0 + 4);
        return make_uint4(pack_bf16x2(__uint_as_float(a.x), __uint_as_float(a.y)), pack_bf16x2(__uint_as_float(a.z), __uint_as_float(a.w)),
                          pack_bf16x2(__uint_as_float(b.x), __uint_as_float(b.y)), pack_bf16x2(__uint_as_float(b.z), __uint_as_float(b.w)));
    } else if constexpr (SRC_F32) {                // 4 fp32 (tf32 operand: the tensor core ignores the low 13 mantissa bits)
        return ldg_nc16(reinterpret_cast<const float *>(row) + k0);
    } else {                                       // 8 bf16
        return ldg_nc16(reinterpret_cast<const __nv_bfloat16 *>(row) + k0);
    }
}

template <bool A_F32, bool TF32>
__global__ void __launch_bounds__(GM_THREADS, 1) rows_gemm_kernel(const __grid_constant__ GemmParams p)
{
    constexpr bool TMA_A = !(A_F32 && !TF32);       // plain copies go through the copy engine; fp32 -> bf16 through registers
    extern __shared__ __align__(1024) unsigned char gsm_raw[];
    unsigned char *gsm = gsm_raw + ((1024u - (smem_u32(gsm_raw) & 1023u)) & 1023u);
    const int w_block = p.Nt * 128;
    unsigned char *Asm = gsm;                                           // [GM_STAGES][128 rows x 128 B]
    unsigned char *Csm = Asm + GM_STAGES * GM_BLOCK;                    // [2][128 rows x 128 B] output staging
    unsigned char *Wsm = Csm + 2 * GM_BLOCK;                            // [nkb][Nt rows x 128 B], 128-byte swizzle
    uint64_t *full = reinterpret_cast<uint64_t *>(Wsm + (size_t)p.nkb * w_block);
    uint64_t *empty = full + GM_STAGES;
    uint64_t *acc_full = empty + GM_STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    constexpr int CPE = TF32 ? 4 : 8;               // elements per 16-byte operand chunk
    constexpr int KPB = TF32 ? 32 : 64;             // elements per 128-byte block row
    const int n0 = blockIdx.y * p.Nt;
    const int ncols = min(p.Nt, p.N - n0);

    if (tid == 0) {
        for (int s = 0; s < GM_STAGES; ++s) { mbar_init(full + s, TMA_A ? 1 : GM_LOAD_WARPS); mbar_init(empty + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, GM_EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    // this CTA's slice of W -> shared memory, once
    for (int i = tid; i < p.nkb * p.Nt * 8; i += GM_THREADS) {
        const int kb = i / (p.Nt * 8), rem = i - kb * (p.Nt * 8);
        const int r = rem >> 3, j = rem & 7;
        const void *row = nullptr;
        if (r < ncols) row = TF32 ? (const void *)(reinterpret_cast<const float *>(p.W) + (int64_t)(n0 + r) * p.ldw)
                                  : (const void *)(reinterpret_cast<const __nv_bfloat16 *>(p.W) + (int64_t)(n0 + r) * p.ldw);
        const uint4 v = load_chunk<TF32, false>(row, kb * KPB + j * CPE, p.K);
        *reinterpret_cast<uint4 *>(Wsm + (size_t)kb * w_block + r * 128 + ((j ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= GM_LOAD_WARP0) {
        // ================================ loaders: A blocks -> ring ==========================================================
        const int lt = tid - GM_LOAD_WARP0 * 32;            // 0 .. 255
        uint32_t q = 0;
        if constexpr (TMA_A) {
            if (lt == 0) {
                for (int64_t t = blockIdx.x; t < p.tiles_m; t += gridDim.x)
                    for (int kb = 0; kb < p.nkb; ++kb, ++q) {
                        const int s = q % GM_STAGES;
                        mbar_wait(empty + s, ((q / GM_STAGES) & 1u) ^ 1u);
                        mbar_expect_tx(full + s, (uint32_t)GM_BLOCK);
                        tma_load_2d(Asm + (size_t)s * GM_BLOCK, &p.tmA, kb * KPB, (int)(t * TC_ROWS), full + s);
                    }
            }
        } else
        for (int64_t t = blockIdx.x; t < p.tiles_m; t += gridDim.x) {
            const int64_t m0 = t * TC_ROWS;
            for (int kb = 0; kb < p.nkb; ++kb, ++q) {
                uint4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = lt + GM_LOADERS * i;
                    const int r = c >> 3, j = c & 7;
                    const void *row = nullptr;
                    if (m0 + r < p.M) row = A_F32 ? (const void *)(reinterpret_cast<const float *>(p.A) + (m0 + r) * p.lda)
                                                  : (const void *)(reinterpret_cast<const __nv_bfloat16 *>(p.A) + (m0 + r) * p.lda);
                    v[i] = load_chunk<A_F32, A_F32 && !TF32>(row, kb * KPB + j * CPE, p.K);
                }
                const int s = q % GM_STAGES;
                mbar_wait(empty + s, ((q / GM_STAGES) & 1u) ^ 1u);
                unsigned char *dst = Asm + (size_t)s * GM_BLOCK;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = lt + GM_LOADERS * i;
                    const int r = c >> 3, j = c & 7;
                    *reinterpret_cast<uint4 *>(dst + r * 128 + ((j ^ (r & 7)) << 4)) = v[i];
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full + s);
            }
        }
    } else if (warp == GM_MMA_WARP) {
        // ================================ MMA issue =========================================================================
        const uint32_t idesc = TF32 ? instr_desc_tf32(p.Nt) : instr_desc_bf16(p.Nt);
        uint32_t q = 0, tl = 0;
        for (int64_t t = blockIdx.x; t < p.tiles_m; t += gridDim.x, ++tl) {
            const uint32_t buf = tl & 1u;
            mbar_wait(acc_empty + buf, ((tl >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.Nt;
            for (int kb = 0; kb < p.nkb; ++kb, ++q) {
                const int s = q % GM_STAGES;
                mbar_wait(full + s, (q / GM_STAGES) & 1u);
                tc_fence_after();
                const uint32_t a_lo = desc_lo_sw128(smem_u32(Asm) + (uint32_t)(s * GM_BLOCK));
                const uint32_t b_lo = desc_lo_sw128(smem_u32(Wsm) + (uint32_t)(kb * w_block));
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                   // four K steps of 32 bytes per block
                        const uint64_t ad = mk_desc64(a_lo + (uint32_t)(2 * j), DESC_HI_SW128);
                        const uint64_t bd = mk_desc64(b_lo + (uint32_t)(2 * j), DESC_HI_SW128);
                        if (TF32) umma_tf32(d_tmem, ad, bd, idesc, (kb | j) ? 1u : 0u);
                        else umma_bf16(d_tmem, ad, bd, idesc, (kb | j) ? 1u : 0u);
                    }
                    umma_commit(empty + s);                         // frees the ring slot when these MMAs retire
                    if (kb == p.nkb - 1) umma_commit(acc_full + buf);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ epilogue: TMEM -> staging tile -> TMA store ========================================
        const int qd = warp;                                        // TMEM lane quarter
        const int row = qd * 32 + lane;
        const uint32_t row_off = (uint32_t)(row * 128);
        const int cps = p.c_bf16 ? 64 : 32;                         // output columns per 128-byte staging row
        uint32_t tl = 0, sq = 0;                                    // sq: staging tiles written so far (buffer = sq & 1)
        for (int64_t t = blockIdx.x; t < p.tiles_m; t += gridDim.x, ++tl) {
            const uint32_t buf = tl & 1u;
            mbar_wait(acc_full + buf, (tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + buf * (uint32_t)p.Nt + ((uint32_t)(qd * 32) << 16);
            for (int c0 = 0; c0 < ncols; c0 += cps, ++sq) {
                unsigned char *stg = Csm + (size_t)(sq & 1u) * GM_BLOCK;
                if (tid == 0) tma_store_wait_read<1>();             // the store that read this buffer two tiles ago is done with it
                epi_bar();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && !p.c_bf16) break;
                    const int cc = c0 + 32 * h;                     // 32 accumulator columns -> 8 (fp32) or 4 (bf16) chunks
                    uint32_t r[32];
                    if (cc < p.Nt) {
                        if (cc + 32 <= p.Nt) tmem_ld32_nowait(trow + (uint32_t)cc, r);
                        else {
                            uint32_t r16[16];
                            tmem_ld16(trow + (uint32_t)cc, r16);
#pragma unroll
                            for (int i = 0; i < 16; ++i) { r[i] = r16[i]; r[16 + i] = 0u; }
                        }
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = 0u;
                    }
                    if (p.c_bf16) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int j = 4 * h + i;
                            *reinterpret_cast<uint4 *>(stg + row_off + ((j ^ (row & 7)) << 4)) =
                                make_uint4(pack_bf16x2(__uint_as_float(r[8 * i + 0]), __uint_as_float(r[8 * i + 1])),
                                           pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                                           pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                                           pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4 *>(stg + row_off + ((j ^ (row & 7)) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    }
                }
                fence_proxy_async();
                epi_bar();
                if (tid == 0) {
                    tma_store_2d(&p.tmC, n0 + c0, (int)(t * TC_ROWS), stg);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);
        }
        if (tid == 0) tma_store_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}


// ------------------------------------------------------------------------------------------------------------------------
// Weight gradient of a 1x1 convolution on rows:  dW[Cout,Cin] = dy[M,Cout]^T . x[M,Cin]  (M = grouped rows, in the millions;
// the output a few thousand elements).  The reduction runs over the ROW index, so both operands are "MN-major" for the
// tensor core: a block of 32 rows x 32 channels (128 bytes per row) copied as it lies in HBM IS the canonical MN-major
// operand (channels contiguous; SWIZZLE_128B_BASE32B, the 32-byte-atom swizzle TMA applies as CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B),
// delivered without touching a register.  A CTA
// owns a contiguous range of rows and one (128 output channels) x (<= 256 input channels) tile of dW, accumulates it in TMEM
// over its whole range (one tf32 MMA per 8 rows), and writes an fp32 partial; a second small kernel adds the partials in a
// fixed order (deterministic).
constexpr int WG_KT = 32;                         // rows per ring stage
constexpr int WG_STAGES = 4;
constexpr int WG_THREADS = 6 * 32;                // 4 epilogue warps, MMA warp, TMA warp
constexpr int WG_MNBLOCK = WG_KT * 128;           // 32 rows x 128 bytes: one 32-channel block of a stage

struct WgradParams {
    int64_t M, rows_per_cta;
    int Cout, Cin, nt, nbb;       // nt: input channels per tile (multiple of 16, <= 256); nbb = ceil(nt / 32)
    int n_ci_tiles, tmem_cols;
    float *ws;                    // [gridDim.x][Cout][Cin] partials
    alignas(64) CUtensorMap tmY;  // dy as [M][Cout], box 32 channels x 32 rows
    alignas(64) CUtensorMap tmX;  // x as [M][Cin],  box 32 channels x 32 rows
};

// MN-major descriptor for 32-bit operands: SWIZZLE_128B_BASE32B (layout type 1; the only layout the tensor core takes for
// MN-major tf32): rows of 128 bytes = 32 channels, atoms of 4 rows with 32-byte chunk c of row r at position c ^ (r & 3);
// LBO = bytes between 32-channel blocks, SBO = bytes between 4-row atoms (512: rows are consecutive)
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((512u >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (1ull << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1) rows_wgrad_kernel(const __grid_constant__ WgradParams p)
{
    extern __shared__ __align__(1024) unsigned char wsm_raw[];
    unsigned char *sm = wsm_raw + ((1024u - (smem_u32(wsm_raw) & 1023u)) & 1023u);
    const int stage_bytes = (4 + p.nbb) * WG_MNBLOCK;
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)WG_STAGES * stage_bytes);
    uint64_t *empty = full + WG_STAGES;
    uint64_t *acc_full = empty + WG_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int co0 = (blockIdx.y / p.n_ci_tiles) * TC_ROWS, ci0 = (blockIdx.y % p.n_ci_tiles) * p.nt;
    const int64_t r_begin = (int64_t)blockIdx.x * p.rows_per_cta;
    int64_t r_end = r_begin + p.rows_per_cta;
    if (r_end > p.M) r_end = p.M;
    const int n_st = r_end > r_begin ? (int)((r_end - r_begin + WG_KT - 1) / WG_KT) : 0;

    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 5) {
        if (lane == 0) {
            for (int i = 0; i < n_st; ++i) {
                const int s = i % WG_STAGES;
                mbar_wait(empty + s, ((i / WG_STAGES) & 1u) ^ 1u);
                mbar_expect_tx(full + s, (uint32_t)stage_bytes);
                unsigned char *dst = sm + (size_t)s * stage_bytes;
                // rows past the CTA's range would double-count: the last stage of a range is clipped by r_end through the
                // tensor map only at the end of the tensor, so ranges are multiples of WG_KT rows (host) except the last one
                const int row = (int)(r_begin + (int64_t)i * WG_KT);
                for (int b = 0; b < 4; ++b) tma_load_2d(dst + b * WG_MNBLOCK, &p.tmY, co0 + 32 * b, row, full + s);
                for (int b = 0; b < p.nbb; ++b) tma_load_2d(dst + (4 + b) * WG_MNBLOCK, &p.tmX, ci0 + 32 * b, row, full + s);
            }
        }
    } else if (warp == 4) {
        // D[co][ci] += dy^T (A: M = 128 output channels, MN-major) x (B: N = nt input channels, MN-major), K = 8 rows per MMA
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.nt >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
        for (int i = 0; i < n_st; ++i) {
            const int s = i % WG_STAGES;
            mbar_wait(full + s, (i / WG_STAGES) & 1u);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sm) + (uint32_t)(s * stage_bytes);
            const uint32_t b0 = a0 + 4u * WG_MNBLOCK;
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < WG_KT / 8; ++j)
                    umma_tf32(tmem_base, smem_desc_mn_sw128(a0 + (uint32_t)(j * 1024), WG_MNBLOCK), smem_desc_mn_sw128(b0 + (uint32_t)(j * 1024), WG_MNBLOCK),
                              idesc, (i | j) ? 1u : 0u);
                umma_commit(empty + s);
                if (i == n_st - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // partial tile -> workspace (zeros when this CTA's row range is empty)
        if (n_st > 0) { mbar_wait(acc_full, 0); tc_fence_after(); }
        const int co = co0 + warp * 32 + lane;
        float *dst = p.ws + ((int64_t)blockIdx.x * p.Cout + co) * p.Cin + ci0;
        const int nci = min(p.nt, p.Cin - ci0);
        for (int c0 = 0; c0 < p.nt; c0 += 16) {
            uint32_t r[16];
            if (n_st > 0) { tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = 0u;
            }
            if (co < p.Cout) {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (c0 + i < nci) dst[c0 + i] = __uint_as_float(r[i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

__global__ void rows_wgrad_reduce_kernel(const float *__restrict__ ws, int n_part, int64_t n, float *__restrict__ dW)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int q = 0; q < n_part; ++q) a += ws[(int64_t)q * n + i];
    dW[i] = a;
}

static void wgrad_plan(int64_t M, int Cout, int Cin, int *gx, int *nt, int *n_ci, int *n_co, int64_t *rows_per_cta)
{
    const int cin_pad = rup(Cin, 16);
    *n_ci = (cin_pad + 255) / 256;
    *nt = rup((cin_pad + *n_ci - 1) / *n_ci, 16);
    *n_co = (Cout + TC_ROWS - 1) / TC_ROWS;
    int g = 148 / (*n_ci * *n_co);
    if (g < 1) g = 1;
    const int64_t units = (M + WG_KT - 1) / WG_KT;
    if (g > units) g = (int)units;
    if (g < 1) g = 1;
    *rows_per_cta = (units + g - 1) / g * WG_KT;
    *gx = (int)((M + *rows_per_cta - 1) / *rows_per_cta);
    if (*gx < 1) *gx = 1;
}

int64_t rows_wgrad_ws_bytes(int64_t M, int Cout, int Cin)
{
    int gx, nt, n_ci, n_co; int64_t rpc;
    wgrad_plan(M, Cout, Cin, &gx, &nt, &n_ci, &n_co, &rpc);
    return (int64_t)gx * Cout * Cin * 4;
}

int launch_rows_wgrad(const float *dy, int64_t ldy, const float *x, int64_t ldx, int64_t M, int Cout, int Cin, float *ws, float *dW, cudaStream_t st)
{
    PN2_REQUIRE(ldy % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0,
                "pn2_rows_wgrad_tf32: operand rows must start on 16-byte boundaries");
    PN2_REQUIRE(M >= 1 && M < (1ll << 31), "pn2_rows_wgrad_tf32: M = %lld outside [1, 2^31)", (long long)M);
    WgradParams p;
    int gx, n_co;
    wgrad_plan(M, Cout, Cin, &gx, &p.nt, &p.n_ci_tiles, &n_co, &p.rows_per_cta);
    p.M = M; p.Cout = Cout; p.Cin = Cin; p.nbb = (p.nt + 31) / 32; p.ws = ws;
    int tc = 32;
    while (tc < p.nt) tc <<= 1;
    p.tmem_cols = tc;
    int rc = make_tile_map(dy, 4, M, Cout, ldy * 4, WG_KT, &p.tmY, 1);
    if (rc) return rc;
    rc = make_tile_map(x, 4, M, Cin, ldx * 4, WG_KT, &p.tmX, 1);
    if (rc) return rc;
    const size_t smem = 1024 + (size_t)WG_STAGES * (4 + p.nbb) * WG_MNBLOCK + 256;
    cudaError_t e = cudaFuncSetAttribute(rows_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("rows_wgrad: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
    rows_wgrad_kernel<<<dim3((unsigned)gx, (unsigned)(n_co * p.n_ci_tiles)), WG_THREADS, smem, st>>>(p);
    rc = launch_status("pn2_rows_wgrad_tf32");
    if (rc) return rc;
    const int64_t n = (int64_t)Cout * Cin;
    rows_wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, gx, n, dW);
    return launch_status("pn2_rows_wgrad_tf32(reduce)");
}

}  // namespace

int launch_rows_gemm(const void *A, int64_t lda, int a_f32, const void *W, int64_t ldw, int w_f32, void *C, int64_t ldc, int c_bf16,
                     int64_t M, int N, int K, cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    const bool tf32 = w_f32 != 0;
    PN2_REQUIRE(!tf32 || a_f32, "pn2_rows_gemm: fp32 weights (TF32 product) need fp32 rows");
    const int cpe_a = a_f32 ? (tf32 ? 4 : 8) : 8;
    PN2_REQUIRE(K >= 1 && K % cpe_a == 0 && K % (tf32 ? 4 : 8) == 0, "pn2_rows_gemm: K = %d must be a multiple of %d (pad the operands with zero columns)", K, cpe_a);
    PN2_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && (a_f32 || lda % 8 == 0) && (tf32 || ldw % 8 == 0),
                "pn2_rows_gemm: operand rows must start on 16-byte boundaries");
    PN2_REQUIRE(ldc % (c_bf16 ? 8 : 4) == 0 && ((uintptr_t)C & 15) == 0, "pn2_rows_gemm: output rows must start on 16-byte boundaries");
    GemmParams p;
    p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
    p.a_f32 = a_f32; p.tf32 = tf32 ? 1 : 0; p.c_bf16 = c_bf16;
    const int kpb = tf32 ? 32 : 64;
    p.nkb = (K + kpb - 1) / kpb;
    // columns per CTA: W slice resident next to the A ring
    const int budget = TC_SMEM_LIMIT - (GM_STAGES + 2) * GM_BLOCK - 2048;
    int nt_max = budget / (p.nkb * 128) / 16 * 16;
    if (nt_max > 256) nt_max = 256;
    PN2_REQUIRE(nt_max >= 16, "pn2_rows_gemm: K = %d too large for a resident weight slice", K);
    const int n_tiles = (rup(N, 16) + nt_max - 1) / nt_max;
    p.Nt = rup((N + n_tiles - 1) / n_tiles, 16);
    p.tiles_m = (M + TC_ROWS - 1) / TC_ROWS;
    int tc = 32;
    while (tc < 2 * p.Nt) tc <<= 1;
    p.tmem_cols = tc;
    const size_t smem = 1024 + (size_t)p.nkb * p.Nt * 128 + (size_t)(GM_STAGES + 2) * GM_BLOCK + 256;
    const bool tma_a = !(a_f32 && !tf32);
    if (tma_a) {
        const int rc = make_tile_map(A, a_f32 ? 4 : 2, M, K, lda * (a_f32 ? 4 : 2), TC_ROWS, &p.tmA);
        if (rc) return rc;
    } else memset(&p.tmA, 0, sizeof(p.tmA));
    {
        const int rc = make_tile_map(C, c_bf16 ? 2 : 4, M, N, ldc * (c_bf16 ? 2 : 4), TC_ROWS, &p.tmC);
        if (rc) return rc;
    }
    int64_t gx = 148 / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > p.tiles_m) gx = p.tiles_m;
    dim3 grid((unsigned)gx, (unsigned)n_tiles);
    auto run = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("rows_gemm: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        kern<<<grid, GM_THREADS, smem, st>>>(p);
        return launch_status("pn2_rows_gemm");
    };
    if (tf32) return run(rows_gemm_kernel<true, true>);
    return a_f32 ? run(rows_gemm_kernel<true, false>) : run(rows_gemm_kernel<false, false>);
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_rows_gemm(const void *A, int64_t lda, int a_dtype, const void *W, int64_t ldw, int w_dtype,
                             void *C, int64_t ldc, int c_dtype, int64_t M, int N, int K, void *stream)
{
    PN2_REQUIRE(A && W && C, "pn2_rows_gemm: null pointer");
    PN2_REQUIRE(M >= 0 && N >= 0 && K >= 1, "pn2_rows_gemm: bad sizes");
    PN2_REQUIRE((a_dtype == PN2_DT_F32 || a_dtype == PN2_DT_BF16) && (w_dtype == PN2_DT_F32 || w_dtype == PN2_DT_BF16) &&
                (c_dtype == PN2_DT_F32 || c_dtype == PN2_DT_BF16), "pn2_rows_gemm: dtypes are PN2_DT_F32 or PN2_DT_BF16");
    PN2_REQUIRE(M < (1ll << 40), "pn2_rows_gemm: M too large");
    return launch_rows_gemm(A, lda, a_dtype == PN2_DT_F32, W, ldw, w_dtype == PN2_DT_F32, C, ldc, c_dtype == PN2_DT_BF16, M, N, K,
                            (cudaStream_t)stream);
}

extern "C" int64_t pn2_rows_wgrad_ws_bytes(int64_t M, int Cout, int Cin)
{
    if (M < 1 || Cout < 1 || Cin < 1) return 0;
    return rows_wgrad_ws_bytes(M, Cout, Cin);
}

extern "C" int pn2_rows_wgrad_tf32(const float *dy, int64_t ldy, const float *x, int64_t ldx, int64_t M, int Cout, int Cin,
                                   void *ws, float *dW, void *stream)
{
    PN2_REQUIRE(dy && x && ws && dW, "pn2_rows_wgrad_tf32: null pointer");
    PN2_REQUIRE(Cout >= 1 && Cin >= 1 && Cout % 4 == 0 && Cin % 4 == 0, "pn2_rows_wgrad_tf32: channel counts must be multiples of 4 (16-byte rows)");
    return launch_rows_wgrad(dy, ldy, x, ldx, M, Cout, Cin, (float *)ws, dW, (cudaStream_t)stream);
}
