// Pieces shared by the pipelined fused set-abstraction kernels (pn2_sa_tc2.cu, pn2_sa_tc4.cu): parameter block,
// shared-memory carve-up, hidden-layer epilogue, max-pool, MMA issue.  Not part of the ABI.
#pragma once
#include <cuda.h>
#include "pn2_tc.cuh"

namespace pn2 {

constexpr int T2_MAX_UNITS = 4;
constexpr int T2_MAX_THREADS = 4 * T2_MAX_UNITS * 32;         // 512
constexpr int T2_SMEM_LIMIT = 227 * 1024;
constexpr int T2_MAX_L = 4;
constexpr int ONES_BYTES = 2 * A_LBO;                         // constant activation slice [2][128][16 B]

// per layer, precomputed on the host: everything the MMA issue needs, as descriptor words relative to the start of
// the weight image in shared memory
struct Tc2Layer {
    uint32_t w_lo_rel;    // weight operand: (byte offset >> 4) | (LBO >> 4) << 16
    uint32_t w_step;      // descriptor increment per K = 16 step
    uint32_t bs_lo_rel;   // hidden layers: the bias slice, same encoding
    uint32_t idesc;       // instruction descriptor (N = layer width; last layer: M = 128 channels, N = 128 rows)
    uint32_t s_cbf16, s_cbl16;   // split layout (fp32 mode): bytes >> 4 of one image (hi or lo) of a full / of the last weight chunk
    int s_spc;            // split layout: K = 16 steps per weight chunk (kblk / 16)
    int nks;              // K = 16 steps over the swizzled activation blocks (layer 0: the D feature channels)
    int nb32;             // hidden layers: Npad / 32
    int n_mh;             // last layer: Npad / 128
    int Cout, b_off;
};

struct Tc2Params {
    const float *points; int64_t sb, sn, sc;
    const float *feat;   int64_t fb, fn, fc;
    const float *centres; const int32_t *idx32;
    int N, S, K, C, D, order;
    int L;
    Tc2Layer ly[T2_MAX_L];
    int img_bytes, bias_off;   // weight image = the whole packed buffer; fp32 biases at bias_off
    const unsigned char *packed;
    float *out; int64_t ob, os; int co_off;
    int64_t G, rows, tiles;
    int U, cols_per_unit, tmem_cols, a_unit_bytes, xyz_off, feat_fast;   // xyz_off: the no-swizzle coordinate block inside a unit's buffer
    const __nv_bfloat16 *feat_bf16; __nv_bfloat16 *out_bf16;
    unsigned long long *trace;   // debug timeline (pn2_debug_trace), CTA 0 only; nullptr = off
    int use_tma, feat_rows;      // bf16 features gathered by TMA (tile::gather4) through `tmap` = [feat_rows = B*N][D]
    int static_sched;            // pn2_sa_tc4.cu: 1 = fixed (layer, unit) schedule, 0 = units polled dynamically
    int split;                   // fp32 mode: every operand as (bf16 hi, bf16 lo), four MMAs per K step (sa_tc2_kernel<0, ., true>)
    uint32_t a_lo_off;           // split: bytes from a unit's hi operand image to its lo image
    alignas(64) CUtensorMap tmap;
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t relu_pack_u(uint32_t lo, uint32_t hi) { return relu_pack_bf16x2(__uint_as_float(lo), __uint_as_float(hi)); }

// ReLU + bf16 of 32 accumulator columns of this thread's row -> chunks J0 .. J0+3 of a 64-channel swizzled block.
// a_rowx = address of the row inside block 0 with (row & 7) << 4 already added: the 16-byte chunk j of the row lives
// at a_rowx ^ (j << 4) (the block is 1024-byte aligned, so bits 4..6 of the row address are the swizzle term alone)
template <int J0>
__device__ __forceinline__ void epi_store32(const uint32_t (&r)[32], uint32_t a_rowx, uint32_t blk_off)
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
        sts128((a_rowx ^ (uint32_t)((J0 + i) << 4)) + blk_off, relu_pack_u(r[8 * i + 0], r[8 * i + 1]), relu_pack_u(r[8 * i + 2], r[8 * i + 3]),
               relu_pack_u(r[8 * i + 4], r[8 * i + 5]), relu_pack_u(r[8 * i + 6], r[8 * i + 7]));
}
// hidden-layer epilogue of NB consecutive 32-column blocks starting at column 0 (thread = row), two loads in flight
template <int NB>
__device__ __forceinline__ void epi_hidden(uint32_t tmem_row, uint32_t a_rowx)
{
#ifdef PN2_EPI_PAIR
#pragma unroll
    for (int b = 0; b < NB; b += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(tmem_row + (uint32_t)(32 * b), r0);
        if (b + 1 < NB) tmem_ld32_nowait(tmem_row + (uint32_t)(32 * b + 32), r1);
        tmem_ld_wait();
        epi_store32<0>(r0, a_rowx, (uint32_t)((b >> 1) * SW128_BLOCK));
        if (b + 1 < NB) epi_store32<4>(r1, a_rowx, (uint32_t)((b >> 1) * SW128_BLOCK));
    }
#else
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint32_t r0[32];
        tmem_ld32_nowait(tmem_row + (uint32_t)(32 * b), r0);
        tmem_ld_wait();
        if (b & 1) epi_store32<4>(r0, a_rowx, (uint32_t)((b >> 1) * SW128_BLOCK));
        else epi_store32<0>(r0, a_rowx, (uint32_t)((b >> 1) * SW128_BLOCK));
    }
#endif
}
// any width, blocks b0, b0 + step, ... < nb
__device__ __forceinline__ void epi_hidden_rt(uint32_t tmem_row, uint32_t a_rowx, int b0, int step, int nb)
{
#pragma unroll 1
    for (int b = b0; b < nb; b += step) {
        uint32_t r[32];
        tmem_ld32_nowait(tmem_row + (uint32_t)(32 * b), r);
        tmem_ld_wait();
        epi_store32<0>(r, a_rowx ^ (uint32_t)((b & 1) << 6), (uint32_t)((b >> 1) * SW128_BLOCK));
    }
}
__device__ __forceinline__ void epi_hidden_any(int nb32, uint32_t tmem_row, uint32_t a_rowx)
{
    switch (nb32) {
    case 1: epi_hidden<1>(tmem_row, a_rowx); break;
    case 2: epi_hidden<2>(tmem_row, a_rowx); break;
    case 3: epi_hidden<3>(tmem_row, a_rowx); break;
    case 4: epi_hidden<4>(tmem_row, a_rowx); break;
    case 8: epi_hidden<8>(tmem_row, a_rowx); break;
    default: epi_hidden_rt(tmem_row, a_rowx, 0, 1, nb32); break;
    }
}

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// max of v[0..N) as a tree of 3-input maxima
template <int N>
__device__ __forceinline__ float max_all(const float (&v)[N])
{
    if constexpr (N == 1) return v[0];
    else if constexpr (N == 2) return fmaxf(v[0], v[1]);
    else if constexpr (N == 3) return max3(v[0], v[1], v[2]);
    else {
        constexpr int M = (N + 2) / 3;
        float w[M];
#pragma unroll
        for (int i = 0; i < N / 3; ++i) w[i] = max3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        if constexpr (N % 3 == 1) w[M - 1] = v[N - 1];
        if constexpr (N % 3 == 2) w[M - 1] = fmaxf(v[N - 2], v[N - 1]);
        return max_all<M>(w);
    }
}
template <int OFF, int CNT>
__device__ __forceinline__ float max_range(const uint32_t (&r)[32])
{
    float v[CNT];
#pragma unroll
    for (int j = 0; j < CNT; ++j) v[j] = __uint_as_float(r[OFF + j]);
    return max_all<CNT>(v);
}

// where the pooled value of (group, channel) goes; groups are visited in order
struct PoolOut {
    float *po; __nv_bfloat16 *pb;       // element (cloud 0, group 0, this thread's channel)
    int64_t os, ob;
    int S, gg, G;
    float bv; bool c_ok, all_valid, has_bf, contig;      // contig: ob == S * os, group g lives at g * os
    bool use_atomic = false;                             // a tile is only part of its group (K > 128): combine with atomicMax
    // group finished: max over its rows -> + bias -> ReLU -> [B,S,Ctot]; then step to the next group
    __device__ __forceinline__ void emit(float m)
    {
        if (c_ok && gg < G) {
            const float y = fmaxf(m + bv, 0.f);       // an empty ball (every row masked to -inf) pools to 0, never to stale memory
            int64_t off;
            if (contig) off = (int64_t)gg * os;
            else { const int b = (int)((unsigned)gg / (unsigned)S); off = (int64_t)b * ob + (int64_t)(gg - b * S) * os; }
            if (use_atomic) atomicMax(reinterpret_cast<int *>(po + off), __float_as_int(y));      // y >= 0: int order = float order
            else po[off] = y;
            if (has_bf) pb[off] = __float2bfloat16_rn(y);
        }
        ++gg;
    }
};
__device__ __forceinline__ void mask_rows(uint32_t (&r)[32], uint32_t mk)
{
#pragma unroll
    for (int j = 0; j < 32; ++j) if (!((mk >> j) & 1u)) r[j] = 0xff800000u;   // -inf
}
// one 32-row block of the transposed accumulator (this thread = one channel)
template <int K>
__device__ __forceinline__ void pool_block(const uint32_t (&r)[32], int blk, float &run, PoolOut &o)
{
    if constexpr (K == 32) o.emit(max_range<0, 32>(r));
    else if constexpr (K > 32) {
        run = fmaxf(run, max_range<0, 32>(r));
        if ((((blk + 1) * 32) & (K - 1)) == 0) { o.emit(run); run = -INFINITY; }
    } else if constexpr (K == 16) {
        o.emit(max_range<0, 16>(r));
        o.emit(max_range<16, 16>(r));
    } else {
        o.emit(max_range<0, 8>(r));
        o.emit(max_range<8, 8>(r));
        o.emit(max_range<16, 8>(r));
        o.emit(max_range<24, 8>(r));
    }
}
// blocks [b_lo, b_hi) of one 128-channel slice; taddr = TMEM address of (this warp's lanes, column of block 0);
// vm = the tile's four valid-row ballots
template <int K, bool PAIR>
__device__ __forceinline__ void pool_slice(uint32_t taddr, int b_lo, int b_hi, const uint32_t *vm, PoolOut &o)
{
    float run = -INFINITY;
    int blk = b_lo;
#pragma unroll 1
    for (; PAIR && blk + 1 < b_hi; blk += 2) {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(taddr + (uint32_t)(blk * 32), r0);
        tmem_ld32_nowait(taddr + (uint32_t)(blk * 32 + 32), r1);
        tmem_ld_wait();
        if (!o.all_valid) { mask_rows(r0, vm[blk]); mask_rows(r1, vm[blk + 1]); }   // padding rows / empty balls
        pool_block<K>(r0, blk, run, o);
        pool_block<K>(r1, blk + 1, run, o);
    }
#pragma unroll 1
    for (; blk < b_hi; ++blk) {
        uint32_t r0[32];
        tmem_ld32_nowait(taddr + (uint32_t)(blk * 32), r0);
        tmem_ld_wait();
        if (!o.all_valid) mask_rows(r0, vm[blk]);
        pool_block<K>(r0, blk, run, o);
    }
}
template <bool PAIR>
__device__ __forceinline__ void pool_slice_any(int logK, uint32_t taddr, int b_lo, int b_hi, const uint32_t *vm, PoolOut &o)
{
    switch (logK) {
    case 3: pool_slice<8, PAIR>(taddr, b_lo, b_hi, vm, o); break;
    case 4: pool_slice<16, PAIR>(taddr, b_lo, b_hi, vm, o); break;
    case 5: pool_slice<32, PAIR>(taddr, b_lo, b_hi, vm, o); break;
    case 6: pool_slice<64, PAIR>(taddr, b_lo, b_hi, vm, o); break;
    default: pool_slice<128, PAIR>(taddr, b_lo, b_hi, vm, o); break;
    }
}

__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, bool pred)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(pred ? 16 : 0) : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
// one probe of a barrier phase, sleeping in hardware for up to `ns` when it has not completed
__device__ __forceinline__ bool mbar_probe(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
    return ok != 0;
}
int launch_sa_tc4(const Tc2Params &p, size_t smem, cudaStream_t st);

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

constexpr uint32_t DESC_HI = DESC_HI_NS;      // SBO = 128 B, descriptor version 1

// Every tcgen05.mma of layer l for one tile (called by ONE elected lane; all arguments warp-uniform).
//   wsm16  = shared address of the weight image >> 4      a_sw = descriptor low word of the unit's swizzled block 0
//   a_xyz  = ... of the unit's coordinate block            ones = ... of the constant slice      d_tm = TMEM accumulator
__device__ __forceinline__ void issue_layer_mmas(const Tc2Params &p, int l, uint32_t wsm16, uint32_t a_sw, uint32_t a_xyz, uint32_t ones,
                                                 uint32_t d_tm)
{
    const Tc2Layer &y = p.ly[l];
    const int nks = y.nks;
    const uint32_t idesc = y.idesc, w_step = y.w_step;
    if (l < p.L - 1) {
        // D[row][channel] = X (M operand) x W^T, then + 1 x bias
        uint32_t w = wsm16 + y.w_lo_rel, acc = 0;
        for (int ks = 0; ks < nks; ++ks) {
            umma_bf16(d_tm, mk_desc64(a_sw + (uint32_t)(((ks >> 2) << 10) + ((ks & 3) << 1)), DESC_HI_SW128), mk_desc64(w, DESC_HI), idesc, acc);
            w += w_step; acc = 1u;
        }
        if (l == 0) { umma_bf16(d_tm, mk_desc64(a_xyz, DESC_HI), mk_desc64(w, DESC_HI), idesc, acc); }
        umma_bf16(d_tm, mk_desc64(ones, DESC_HI), mk_desc64(wsm16 + y.bs_lo_rel, DESC_HI), idesc, 1u);
    } else {
        // transposed: D[channel][row] = W (128 output channels as M) x X^T (the tile's 128 rows as N)
        for (int mh = 0; mh < y.n_mh; ++mh) {
            uint32_t w = wsm16 + y.w_lo_rel + (uint32_t)(mh * TC_ROWS), acc = 0;
            const uint32_t d = d_tm + (uint32_t)(mh * TC_ROWS);
            for (int ks = 0; ks < nks; ++ks) {
                umma_bf16(d, mk_desc64(w, DESC_HI), mk_desc64(a_sw + (uint32_t)(((ks >> 2) << 10) + ((ks & 3) << 1)), DESC_HI_SW128), idesc, acc);
                w += w_step; acc = 1u;
            }
            if (l == 0) umma_bf16(d, mk_desc64(w, DESC_HI), mk_desc64(a_xyz, DESC_HI), idesc, acc);
        }
    }
}

// Split operands (fp32 mode, pn2_sa_tcs.cu explains the arithmetic): the same layer with every K = 16 step issued four times,
// (hi, hi), (hi, lo), (lo, hi), (lo, lo).  Weight chunks are [hi image | lo image] of s_spc steps each; `a_lo16` = distance
// (>> 4) from a unit's hi activation image to its lo image (swizzled blocks and coordinate block alike).
__device__ __forceinline__ void issue_layer_mmas_split(const Tc2Params &p, int l, uint32_t wsm16, uint32_t a_sw, uint32_t a_xyz, uint32_t ones,
                                                       uint32_t d_tm, uint32_t a_lo16)
{
    const Tc2Layer &y = p.ly[l];
    const int nks = y.nks;
    const int nst = nks + (l == 0 ? 1 : 0);             // layer 0: the coordinate block is the last K step
    const uint32_t idesc = y.idesc, w_step = y.w_step;
    const bool hidden = l < p.L - 1;
    const int n_mh = hidden ? 1 : y.n_mh;
    for (int mh = 0; mh < n_mh; ++mh) {
        const uint32_t wb = wsm16 + y.w_lo_rel + (uint32_t)(mh * TC_ROWS);
        const uint32_t d = d_tm + (uint32_t)(mh * TC_ROWS);
        uint32_t acc = 0;
        for (int t = 0; t < nst; ++t) {
            const int c = t / y.s_spc, r = t - c * y.s_spc;
            const bool last_chunk = (c + 1) * y.s_spc >= nst;
            const uint32_t wh = wb + (uint32_t)c * 2u * y.s_cbf16 + (uint32_t)r * w_step;
            const uint32_t wl = wh + (last_chunk ? y.s_cbl16 : y.s_cbf16);
            uint64_t ah, al;
            if (t < nks) {
                const uint32_t a = a_sw + (uint32_t)(((t >> 2) << 10) + ((t & 3) << 1));
                ah = mk_desc64(a, DESC_HI_SW128); al = mk_desc64(a + a_lo16, DESC_HI_SW128);
            } else {
                ah = mk_desc64(a_xyz, DESC_HI); al = mk_desc64(a_xyz + a_lo16, DESC_HI);
            }
            const uint64_t dwh = mk_desc64(wh, DESC_HI), dwl = mk_desc64(wl, DESC_HI);
            if (hidden) {
                umma_bf16(d, ah, dwh, idesc, acc);
                umma_bf16(d, ah, dwl, idesc, 1u);
                umma_bf16(d, al, dwh, idesc, 1u);
                umma_bf16(d, al, dwl, idesc, 1u);
            } else {                                    // transposed: weights are the M operand
                umma_bf16(d, dwh, ah, idesc, acc);
                umma_bf16(d, dwl, ah, idesc, 1u);
                umma_bf16(d, dwh, al, idesc, 1u);
                umma_bf16(d, dwl, al, idesc, 1u);
            }
            acc = 1u;
        }
        if (hidden) umma_bf16(d, mk_desc64(ones, DESC_HI), mk_desc64(wsm16 + y.bs_lo_rel, DESC_HI), idesc, 1u);
    }
}
// hidden-layer epilogue on split operands: accumulator (three-part bias included) -> ReLU in fp32 -> (bf16 hi, bf16 lo) ->
// the hi and lo images of the next layer's operand (same swizzled addressing as epi_store32)
__device__ __forceinline__ void split_pair(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
    float x = __uint_as_float(a), y = __uint_as_float(b);
    x = x > 0.f ? x : (x != x ? x : 0.f);
    y = y > 0.f ? y : (y != y ? y : 0.f);
    hi = pack_bf16x2(x, y);
    lo = pack_bf16x2(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void epi_hidden_split(int nb32, uint32_t tmem_row, uint32_t a_rowx, uint32_t a_lo_off)
{
#pragma unroll 1
    for (int b = 0; b < nb32; ++b) {
        uint32_t r[32];
        tmem_ld32_nowait(tmem_row + (uint32_t)(32 * b), r);
        tmem_ld_wait();
        const uint32_t base = (a_rowx ^ (uint32_t)((b & 1) << 6)) + (uint32_t)((b >> 1) * SW128_BLOCK);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_pair(r[8 * i + 2 * j], r[8 * i + 2 * j + 1], h[j], l[j]);
            const uint32_t addr = base ^ (uint32_t)(i << 4);
            sts128(addr, h[0], h[1], h[2], h[3]);
            sts128(addr + a_lo_off, l[0], l[1], l[2], l[3]);
        }
    }
}

// shared-memory carve-up, identical in both kernels
struct Tc2Smem {
    unsigned char *A_base, *ones, *Wsm;
    float *bias_s;
    uint32_t *vmask;                    // [U][2][4] valid-row ballots, by tile parity
    uint64_t *acc_ready, *a_ready, *feat_ready, *buf_free, *w_bar;
    uint32_t *tmem_slot;
};
__device__ __forceinline__ Tc2Smem carve(unsigned char *smem, const Tc2Params &p)
{
    Tc2Smem s;
    // activation buffers first: their 128-byte-swizzled blocks need 1024-byte alignment
    s.A_base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    s.ones = s.A_base + (size_t)p.U * p.a_unit_bytes;
    s.Wsm = s.ones + ONES_BYTES;
    s.bias_s = reinterpret_cast<float *>(s.Wsm + p.bias_off);
    s.vmask = reinterpret_cast<uint32_t *>(s.Wsm + p.img_bytes);
    s.acc_ready = reinterpret_cast<uint64_t *>(s.vmask + 8 * T2_MAX_UNITS);
    s.a_ready = s.acc_ready + T2_MAX_UNITS;
    s.feat_ready = s.a_ready + T2_MAX_UNITS;
    s.buf_free = s.feat_ready + T2_MAX_UNITS;
    s.w_bar = s.buf_free + T2_MAX_UNITS;
    s.tmem_slot = reinterpret_cast<uint32_t *>(s.w_bar + 1);
    return s;
}
static size_t tc2_smem_bytes(int U, int a_unit_bytes, int img_bytes)
{
    return (size_t)U * a_unit_bytes + ONES_BYTES + (size_t)img_bytes + 8 * T2_MAX_UNITS * 4 + (4 * T2_MAX_UNITS + 1) * 8 + 16 + 1024;   // + alignment slack
}
// start of every CTA: barriers, weight image (bulk copy), constant slice
__device__ __forceinline__ void cta_prologue(const Tc2Smem &s, const Tc2Params &p, int tid, uint32_t a_ready_count, uint32_t feat_ready_count = 1)
{
    if (tid == 0) {
        for (int u = 0; u < T2_MAX_UNITS; ++u) {
            mbar_init(s.acc_ready + u, 1); mbar_init(s.a_ready + u, a_ready_count);
            mbar_init(s.feat_ready + u, feat_ready_count); mbar_init(s.buf_free + u, 1);
        }
        mbar_init(s.w_bar, 1);
        fence_mbar_init();
        mbar_expect_tx(s.w_bar, (uint32_t)p.img_bytes);
        for (int off = 0; off < p.img_bytes; off += 32768) {
            const int nb = min(32768, p.img_bytes - off);
            bulk_g2s(s.Wsm + off, p.packed + off, (uint32_t)nb, s.w_bar);
        }
    }
    // constant slice: k-group 0 = {1, 1, 0, ...} per row ({1, 1, 1, 0, ...} for split operands: three bias parts), k-group 1 = 0
    for (int i = tid; i < 2 * TC_ROWS; i += (int)blockDim.x)
        *reinterpret_cast<uint4 *>(s.ones + i * 16) = make_uint4(i < TC_ROWS ? 0x3F803F80u : 0u, (i < TC_ROWS && p.split) ? 0x00003F80u : 0u, 0u, 0u);
    fence_proxy_async();
}


}  // namespace pn2
