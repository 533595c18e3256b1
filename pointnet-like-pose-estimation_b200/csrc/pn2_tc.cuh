// Shared by the tensor-core translation units (pn2_sa_tc.cu, pn2_sa_tc2.cu): packed-weight layout and the
// tcgen05 / mbarrier / bulk-copy PTX wrappers.  Not part of the ABI.
#pragma once
#include <stdlib.h>
#include <cuda_bf16.h>
#include "pn2_common.cuh"

namespace pn2 {

constexpr int TC_ROWS = 128;
constexpr int TC_CWARPS = 8;                       // compute warps (gather, MMA issue, epilogue)
constexpr int TC_CTHREADS = TC_CWARPS * 32;        // 256
constexpr int TC_THREADS = TC_CTHREADS + 32;       // + 1 producer warp for the weight ring
constexpr int A_LBO = TC_ROWS * 16;                // bytes between k-groups of the A operand
constexpr int KBLK = 64;                           // k extent of one weight chunk
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_LIMIT = 227 * 1024;

static inline int rup(int x, int a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------------
// Packed layout shared by host (pack + launch) and device.
// ------------------------------------------------------------------------------------------------------
struct TcLayer {
    int K;        // padded input channels (multiple of 16)
    int Cout;     // real output channels
    int Npad;     // output channels padded to 32 (hidden: = next layer's K) or 128 (last layer)
    int NS;       // columns per MMA slab (<= 256, divides Npad)
    int nslab;
    int nkb;      // number of 64-wide k blocks
    int w_off;    // byte offset of this layer's first chunk inside the packed buffer
    int b_off;    // float offset of this layer's bias inside the bias block
    int bs_off;   // hidden layers: byte offset (from the start of the packed buffer) of the layer's bias slice
};
struct TcLayout {
    int L;
    TcLayer ly[PN2_MAX_LAYERS];
    int w_bytes;       // all weight chunks
    int bias_floats;   // sum of Npad
    int max_chunk;     // bytes of the largest chunk
    int n_chunks;
    int total_bytes;   // weights + fp32 biases + bias slices
    int kblk;          // k extent of a weight chunk: KBLK, or 32 in the split layout (16 KB hi + lo chunks at NS = 128)
    int pair_halves;   // chunks packed as two contiguous column halves (CTA-pair kernel, see tc_pair_env)
    int replicas;      // identical copies of the packed image in global memory (streamed stacks, see tc_weight_replicas)
    int rep_stride;    // bytes between copies
    int split;         // fp32 mode on the tensor cores (pn2_sa_tcs.cu): every chunk is [bf16 hi image | bf16 lo image], twice the bytes
};
// Bias slice of a hidden layer (used by the pipelined kernels, pn2_sa_tc2.cu): one extra K = 16 slice of the weight
// operand, [Npad] rows in the canonical K-major no-swizzle layout (k-groups at Npad*16 bytes), holding
// k = 0: bf16(b), k = 1: bf16(b - bf16(b)), k = 2: the bf16 of what is left, zeros elsewhere.  Multiplied by an activation
// slice whose channels 0 and 1 are 1.0 it adds the folded bias inside the MMA (to ~2^-17 relative), so the epilogue is
// ReLU + convert only; the split (fp32 mode) kernels set channel 2 to 1.0 as well: three parts = the fp32 bias to ~2^-25.
constexpr int BIAS_SLICE_K = 16;

static __host__ __device__ inline int tc_chunk_bytes(const TcLayer &ly, int kb, int kblk = KBLK)
{
    const int kw = (ly.K - kb * kblk) < kblk ? (ly.K - kb * kblk) : kblk;
    return ly.NS * kw * 2;
}

// Streamed weight sets are kept in HBM / L2 as several identical copies, CTA i reading copy i % R (PN2_W_REPLICAS, read once):
// every persistent CTA walks the same chunk sequence at about the same time, so with one copy all 148 SMs pull the same few
// kilobytes -- the same few L2 slices -- at any moment.  Distinct addresses spread the same traffic over the whole L2.
static inline int tc_weight_replicas()
{
    static const int r = [] {
        const char *e = getenv("PN2_W_REPLICAS");
        int v = e ? atoi(e) : 1;
        return v < 1 ? 1 : (v > 16 ? 16 : v);
    }();
    return r;
}
// PN2_TC_PAIR=1 (read once): weight sets that can never be resident (> 227 KB) are packed for the CTA-pair kernel -- inside every
// chunk the columns [0, NS/2) of all k-groups first, then [NS/2, NS): each CTA's half of a chunk is one contiguous block
static inline bool tc_pair_env()
{
    static const bool v = [] { const char *e = getenv("PN2_TC_PAIR"); return e && atoi(e) == 1; }();
    return v;
}
static inline int tc_split_kblk()
{
    static const int k = [] { const char *e = getenv("PN2_F32_KBLK"); const int v = e ? atoi(e) : 32; return (v == 16 || v == 64) ? v : 32; }();
    return k;
}

static inline TcLayout tc_layout(int L, const int *dims, bool split = false)
{
    TcLayout t;
    const int sp = split ? 2 : 1;
    // split layout: 32-wide chunks (16 KB of hi + lo at NS = 128) so that an sa2-sized stack keeps two CTAs per SM beside its
    // ring (16.2 -> 10.7 ms); wide stacks stream megabytes per tile with one CTA per SM and want the larger chunks (3.2 vs 4.5 ms)
    bool wide = false;
    for (int l = 0; l <= L; ++l) if (rup(dims[l], 16) > 256) wide = true;
    const int kblk = (split && !wide) ? tc_split_kblk() : KBLK;
    t.split = split ? 1 : 0;
    t.kblk = kblk;
    t.L = L; t.w_bytes = 0; t.bias_floats = 0; t.max_chunk = 0; t.n_chunks = 0;
    int K = rup(dims[0], 16);
    // wide stacks (group_all heads) stream their weights through a small ring next to a large activation
    // buffer: keep the chunks at 16 KB there
    int ns_cap = split ? 128 : 256;
    for (int l = 0; l < L; ++l) if (rup(dims[l], 16) > 256) ns_cap = 128;
    for (int l = 0; l < L; ++l) {
        TcLayer &y = t.ly[l];
        // hidden layers pad to 32 columns (one tcgen05.ld block); the last layer pads to 128 because the
        // pipelined kernel uses its weight matrix as the M = 128 operand of a transposed product
        y.K = K; y.Cout = dims[l + 1]; y.Npad = rup(dims[l + 1], l == L - 1 ? 128 : 32);
        // one MMA slab per layer when the layer fits an instruction (N <= 256, any multiple of 16 is legal at
        // M = 128); otherwise the largest power-of-two slab that divides the width
        y.NS = 32;
        if (y.Npad <= ns_cap) y.NS = y.Npad;
        else for (int ns = ns_cap; ns >= 32; ns >>= 1) if (y.Npad % ns == 0) { y.NS = ns; break; }
        y.nslab = y.Npad / y.NS;
        y.nkb = (K + kblk - 1) / kblk;
        y.w_off = t.w_bytes; y.b_off = t.bias_floats;
        t.w_bytes += y.Npad * K * 2 * sp;
        t.bias_floats += y.Npad;
        for (int kb = 0; kb < y.nkb; ++kb) { const int c = sp * tc_chunk_bytes(y, kb, kblk); if (c > t.max_chunk) t.max_chunk = c; }
        t.n_chunks += y.nslab * y.nkb;
        K = y.Npad;
    }
    t.total_bytes = t.w_bytes + t.bias_floats * 4;
    for (int l = 0; l < L; ++l) {
        t.ly[l].bs_off = t.total_bytes;
        if (l < L - 1) t.total_bytes += t.ly[l].Npad * BIAS_SLICE_K * 2;
    }
    // weight sets too large to stay resident in shared memory are streamed per tile: those get the replicas
    t.replicas = t.w_bytes > 100 * 1024 ? tc_weight_replicas() : 1;
    t.pair_halves = (!split && tc_pair_env() && t.w_bytes > TC_SMEM_LIMIT) ? 1 : 0;
    t.rep_stride = rup(t.total_bytes, 1024);
    return t;
}

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// add `bytes` to the barrier's pending transaction count without arriving
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    // the suspend-time hint lets a waiting warp sleep in hardware (woken by the completing arrive) instead of
    // re-issuing the test: waiting warps stay out of the issue slots of the working ones
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// wait on a barrier that CTAs of the cluster arrive on (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA gather: rows r0..r3 (any order, out-of-range rows read as zero) x the tensor map's box of columns starting at
// `col` -> 4 consecutive rows of the box width at `dst` (swizzled as the tensor map says), bytes counted on `bar`
__device__ __forceinline__ void tma_gather4(void *dst, const void *tmap, int col, int r0, int r1, int r2, int r3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, M = 128, N from idesc, K = 16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// CTA-pair forms (cta_group::2): M = 256 over the two CTAs of a cluster, issued by the leader; the commit arrives on the barrier
// at the same shared-memory offset in both CTAs
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// ---- 128-byte-swizzled K-major operand (UMMA LayoutType::SWIZZLE_128B) ---------------------------------------
// A block holds 64 bf16 channels of 128 rows: row r is 128 contiguous bytes at r*128, its eight 16-byte chunks
// stored at chunk position (j ^ (r & 7)).  Whole rows are therefore contiguous (coalesced global copies, one
// 128-byte line per 8 lanes, bank-conflict free) and a warp writing one chunk of 32 consecutive rows is
// conflict free as well.  Blocks of 64 channels follow each other at SW128_BLOCK bytes; they must be 1024-byte
// aligned.  The K = 16 slice number ks of a block is addressed by advancing the start address by 32*ks bytes.
constexpr int SW128_BLOCK = TC_ROWS * 128;      // 16 KB
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int row, int chunk /* global chunk index = channel / 8 */)
{
    return (uint32_t)((chunk >> 3) * SW128_BLOCK + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr)
{
    return (uint64_t)desc_lo_sw128(saddr) | ((uint64_t)DESC_HI_SW128 << 32);
}
// low / high words of a no-swizzle descriptor (see smem_desc): stepping the start address by X bytes is lo += X >> 4
__device__ __forceinline__ uint32_t desc_lo_ns(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
constexpr uint32_t DESC_HI_NS = (128u >> 4) | (1u << 14);      // SBO = 128 B, descriptor version 1
__device__ __forceinline__ uint64_t mk_desc64(uint32_t lo, uint32_t hi)
{
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc_bf16(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void cbar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_CTHREADS) : "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits)
    return *reinterpret_cast<const uint32_t *>(&h);
}

// one reduce-scatter step of the max over rows: N live columns -> N/2; partner = lane ^ o
template <int N>
__device__ __forceinline__ void rs_step(float (&v)[32], int o, bool upper)
{
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
        const float a = v[j], b = v[j + N / 2];
        const float keep = upper ? b : a, send = upper ? a : b;
        v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
}


__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// max over the 32 lanes of a warp, result uniform (CREDUX.MAX.F32 on sm_100a)
__device__ __forceinline__ float warp_max_f32(float x)
{
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
    return r;
}
// relu(lo), relu(hi) -> packed bf16x2 (lo in the low half); one F2FP.RELU.BF16.F32.PACK_AB
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace pn2
