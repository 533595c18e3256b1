// Internal helpers shared by the translation units of libpn2_sm100.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "pn2.h"

namespace pn2 {

void set_error(const char *fmt, ...);  // thread-local message behind pn2_last_error_string()

#define PN2_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::pn2::set_error(__VA_ARGS__);     \
            return -1;                         \
        }                                      \
    } while (0)

inline int launch_status(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// Order in which ATen's CPU sum reduces the C squared terms of one point ("summation
// order").  Each add is rounded (no FMA, no reassociation).
//   contiguous operand: 4 interleaved partial sums ("ILP-4", C < 8) or 8 vector lanes + tail first (C >= 8);
//   strided operand (the [B,C,N].permute(0,2,1) view, unit stride along the point index): one accumulator
//   in ascending k for the points ATen vectorises over (n < 32*floor(N/32)), ILP-4 for the leftover points.
template <int C>
__device__ __forceinline__ float sum_ilp4(const float (&t)[C])
{
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int n4 = C / 4;
#pragma unroll
    for (int i = 0; i < n4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = (i == 0) ? t[k] : __fadd_rn(p[k], t[4 * i + k]);
#pragma unroll
    for (int k = 4 * n4; k < C; ++k) p[0] = (n4 == 0 && k == 0) ? t[0] : __fadd_rn(p[0], t[k]);
    return __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[2]), p[3]);
}

// `tail`: the point is one of the leftover points of a strided operand (ignored for contiguous operands;
// for C <= 4 both strided orders coincide)
template <int C, bool CONTIG>
__device__ __forceinline__ float sum_order(const float (&t)[C], bool tail = false)
{
    if constexpr (!CONTIG) {
        if constexpr (C > 4) { if (tail) return sum_ilp4<C>(t); }
        float a = t[0];
#pragma unroll
        for (int k = 1; k < C; ++k) a = __fadd_rn(a, t[k]);
        return a;
    } else if constexpr (C < 8) {
        return sum_ilp4<C>(t);
    } else {
        float lane[8];
        constexpr int n8 = C / 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) lane[j] = t[j];
#pragma unroll
        for (int i = 1; i < n8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) lane[j] = __fadd_rn(lane[j], t[8 * i + j]);
        float a = 0.f;
#pragma unroll
        for (int k = 8 * n8; k < C; ++k) a = __fadd_rn(a, t[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) a = __fadd_rn(a, lane[j]);
        return a;
    }
}

// ---- two points per instruction (add.rn.f32x2 / mul.rn.f32x2 / fma.rn.f32x2, sm_100): lane-wise IEEE
// round-to-nearest, so each half is bit-identical to the scalar sequence -------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 unpk2(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
// Separately rounded product.  NOT mul.rn.f32x2: ptxas (12.9, sm_100a) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// despite the explicit rounding modifiers and -fmad=false (tools/probes/f32x2_contract.cu: 22 % of |p - c|^2 values one
// ulp off, which moves an FPS argmax about once in ten clouds).  fma(a, b, +0) rounds the product once and cannot be
// fused with a following add; the only difference from a multiply is that a -0 product comes out as +0, which none of the
// callers can observe (squares, and dot products that only feed sums and comparisons).
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    const f32x2 zero = 0ull;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(zero));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// sum_order for a pair of points (neither of them a leftover point of a strided operand)
template <int C, bool CONTIG>
__device__ __forceinline__ f32x2 sum_order2(const f32x2 (&t)[C])
{
    if constexpr (!CONTIG || C < 8) {
        if constexpr (!CONTIG) {
            f32x2 a = t[0];
#pragma unroll
            for (int k = 1; k < C; ++k) a = add2(a, t[k]);
            return a;
        } else {
            f32x2 p[4];
            constexpr int n4 = C / 4;
            const f32x2 zero = pk2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = zero;
#pragma unroll
            for (int i = 0; i < n4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) p[k] = (i == 0) ? t[k] : add2(p[k], t[4 * i + k]);
#pragma unroll
            for (int k = 4 * n4; k < C; ++k) p[0] = (n4 == 0 && k == 0) ? t[0] : add2(p[0], t[k]);
            return add2(add2(add2(p[0], p[1]), p[2]), p[3]);
        }
    } else {
        f32x2 lane[8];
        constexpr int n8 = C / 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) lane[j] = t[j];
#pragma unroll
        for (int i = 1; i < n8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) lane[j] = add2(lane[j], t[8 * i + j]);
        f32x2 a = pk2(0.f, 0.f);
#pragma unroll
        for (int k = 8 * n8; k < C; ++k) a = add2(a, t[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) a = add2(a, lane[j]);
        return a;
    }
}

// Same rule for a run-time C (generic fallback kernels).
__device__ __forceinline__ float sum_order_rt(const float *t, int C, bool contig, bool tail = false)
{
    if (!contig && !tail) {
        float a = t[0];
        for (int k = 1; k < C; ++k) a = __fadd_rn(a, t[k]);
        return a;
    }
    if (!contig || C < 8) {
        float p[4] = {0.f, 0.f, 0.f, 0.f};
        int n4 = C / 4;
        for (int i = 0; i < n4; ++i)
            for (int k = 0; k < 4; ++k) p[k] = __fadd_rn(p[k], t[4 * i + k]);
        for (int k = 4 * n4; k < C; ++k) p[0] = __fadd_rn(p[0], t[k]);
        return __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), p[2]), p[3]);
    }
    float lane[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int n8 = C / 8;
    for (int i = 0; i < n8; ++i)
        for (int j = 0; j < 8; ++j) lane[j] = __fadd_rn(lane[j], t[8 * i + j]);
    float a = 0.f;
    for (int k = 8 * n8; k < C; ++k) a = __fadd_rn(a, t[k]);
    for (int j = 0; j < 8; ++j) a = __fadd_rn(a, lane[j]);
    return a;
}

// per-translation-unit launchers (defined in the .cu files, called from pn2_api.cu)
int launch_fps(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
               const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, cudaStream_t st);
int launch_index_points(const float *src, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                        const int64_t *idx, int64_t M, float *out, int32_t *err_flag, cudaStream_t st);
int launch_square_distance(const float *src, int64_t ssb, int64_t ssn, int64_t ssc,
                           const float *dst, int64_t dsb, int64_t dsn, int64_t dsc,
                           int B, int S, int N, int C, float *out, cudaStream_t st);
int launch_ball_query(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                      const float *points, int64_t sb, int64_t sn, int64_t sc,
                      int B, int S, int N, int C, float r2, int K,
                      int64_t *idx64, int32_t *idx32, int32_t *empty_flag, cudaStream_t st);
int launch_gather_rows(const float *feat, int64_t fb, int64_t fn, const int32_t *idx32, int B, int N, int64_t M, int D, float *out, cudaStream_t st);
int launch_scatter_add_rows(const float *grad, const int32_t *idx32, int B, int N, int64_t M, int D, float *gfeat, cudaStream_t st);
int launch_group(const float *points, int64_t sb, int64_t sn, int64_t sc,
                 const float *feat, int64_t fb, int64_t fn, int64_t fc,
                 const int32_t *idx32, const float *centres,
                 int B, int N, int S, int K, int C, int D, int order, float *out, cudaStream_t st);

struct SaArgs {
    const float *points; int64_t sb, sn, sc;
    const float *feat;   int64_t fb, fn, fc;
    const float *centres; const int32_t *idx32;
    int B, N, S, K, C, D, order;
    int L; int dims[PN2_MAX_LAYERS + 1];
    const void *packed;
    float *out; int64_t ob, os; int co_off;
    const void *feat_bf16;   // optional bf16 copy of feat, [B,N,D] contiguous (produced by a previous launch)
    void *out_bf16;          // optional bf16 copy of the pooled output, same (ob, os, co_off) addressing as out
};
int64_t packed_bytes_f32(int L, const int *dims);
int64_t packed_bytes_bf16(int L, const int *dims);
int launch_pack_f32(int L, const int *dims, const void *const *ptrs, float eps, int C, int order, void *packed, cudaStream_t st);
int launch_pack_bf16(int L, const int *dims, const void *const *ptrs, float eps, int C, int order, void *packed, cudaStream_t st, bool split = false);
int64_t packed_bytes_f32_split(int L, const int *dims);
int sa_f32_split_tile_rows(int K, int L, const int *dims);
int launch_sa_f32_split(const SaArgs &a, const void *packed_split, cudaStream_t st);   // -100: configuration not covered
int launch_sa_f32_split_units(const SaArgs &a, const void *packed_split, cudaStream_t st);   // -100: not covered (stacks with feature channels, ...)
int sa_f32_needs_zero(int S, int K, int L, const int *dims);
int sa_bf16_needs_zero(int S, int K, int L, const int *dims);
int launch_sa_f32(const SaArgs &a, cudaStream_t st);
int launch_sa_bf16(const SaArgs &a, cudaStream_t st);
int launch_sa_bf16_pipelined(const SaArgs &a, cudaStream_t st);   // -100: configuration not covered
int sa_bf16_pipelined_covers(int K, int C, int D, int order, int L, const int *dims);
void set_debug_trace(void *buf);
unsigned long long *get_debug_trace();     // the buffer of pn2_debug_trace (NULL = tracing off)
int launch_ball_query_multi(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                            const float *points, int64_t sb, int64_t sn, int64_t sc,
                            int B, int S, int N, int C, int NR, const float *r2, const int *K,
                            int64_t *const *idx64, int32_t *const *idx32, int32_t *empty_flag, cudaStream_t st);
int64_t fps_np_scratch_bytes(int B, int N, int elem_bytes);
int launch_fps_np_f64(const double *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, const int64_t *start, int S,
                      int64_t *idx64, void *scratch, cudaStream_t stream);
int launch_fps_np_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, const int64_t *start, int S,
                      int64_t *idx64, void *scratch, cudaStream_t stream);
int64_t bn_train_ws_bytes(int C);
int launch_bn_fwd(const float *y, int64_t M, int C, int K, const float *gamma, const float *beta, float eps,
                  void *ws, float *stats, float *z, int32_t *arg,
                  float *running_mean, float *running_var, const float *conv_bias, float factor, cudaStream_t st);
int launch_bn_bwd(const float *dz, const float *y, const int32_t *arg, int64_t M, int C, int K, const float *stats,
                  void *ws, float *sums, float *dy, cudaStream_t st);
int launch_normalize_splice(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                            const int64_t *category, int num_category, float *out, cudaStream_t st);
int make_row_gather_map(const void *base, int64_t rows, int cols, void *out_map);
int make_tile_map(const void *base, int elem_bytes, int64_t rows, int64_t cols, int64_t pitch_bytes, int box_rows, void *out_map, int atom32 = 0);
int launch_selftest_tma_gather(const void *src_bf16, int64_t rows, int cols, const int32_t *rows4, int n_inst, int col0, void *out, cudaStream_t st);
int launch_selftest_umma(int n_cols, int k_dim, int variant, float *max_err, float *dump, cudaStream_t st);

}  // namespace pn2
