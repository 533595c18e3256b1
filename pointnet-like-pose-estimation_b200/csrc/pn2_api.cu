// extern "C" surface of libpn2_sm100.so (declared in include/pn2.h): argument validation on the host,
// thread-local error string, dispatch to the per-kernel launchers.  No torch types, no allocation, no sync.
#include <stdarg.h>
#include <string.h>
#include "pn2_common.cuh"

namespace pn2 {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static bool valid_dims(int L, const int *dims)
{
    if (L < 1 || L > PN2_MAX_LAYERS || !dims) return false;
    for (int l = 0; l <= L; ++l) if (dims[l] < 1 || dims[l] > 8192) return false;
    return true;
}
}  // namespace pn2

using namespace pn2;

extern "C" {

int pn2_version(void) { return PN2_ABI_VERSION; }
const char *pn2_last_error_string(void) { return g_err; }

int pn2_fps_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, void *stream)
{
    PN2_REQUIRE(points && start, "pn2_fps_f32: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1 && S >= 0, "pn2_fps_f32: bad sizes B=%d N=%d S=%d", B, N, S);
    PN2_REQUIRE(C >= 1 && C <= PN2_MAX_COORD, "pn2_fps_f32: C=%d outside [1,%d] (summation order only verified there)", C, PN2_MAX_COORD);
    return launch_fps(points, sb, sn, sc, B, N, C, start, S, idx64, idx32, new_points, (cudaStream_t)stream);
}

int pn2_index_points_f32(const float *src, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                         const int64_t *idx, int64_t M, float *out, int32_t *err_flag, void *stream)
{
    PN2_REQUIRE(src && idx && out, "pn2_index_points_f32: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1 && C >= 0 && M >= 0, "pn2_index_points_f32: bad sizes");
    return launch_index_points(src, sb, sn, sc, B, N, C, idx, M, out, err_flag, (cudaStream_t)stream);
}

int pn2_square_distance_f32(const float *src, int64_t ssb, int64_t ssn, int64_t ssc,
                            const float *dst, int64_t dsb, int64_t dsn, int64_t dsc,
                            int B, int S, int N, int C, float *out, void *stream)
{
    PN2_REQUIRE(src && dst && out, "pn2_square_distance_f32: null pointer");
    PN2_REQUIRE(C >= 1 && C <= PN2_MAX_COORD, "pn2_square_distance_f32: C=%d outside [1,%d]", C, PN2_MAX_COORD);
    return launch_square_distance(src, ssb, ssn, ssc, dst, dsb, dsn, dsc, B, S, N, C, out, (cudaStream_t)stream);
}

int pn2_ball_query_f32(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                       const float *points, int64_t sb, int64_t sn, int64_t sc,
                       int B, int S, int N, int C, float r2, int K,
                       int64_t *idx64, int32_t *idx32, int32_t *empty_flag, void *stream)
{
    PN2_REQUIRE(centres && points, "pn2_ball_query_f32: null pointer");
    PN2_REQUIRE(idx64 || idx32, "pn2_ball_query_f32: no output requested");
    PN2_REQUIRE(C >= 1 && C <= PN2_MAX_COORD, "pn2_ball_query_f32: C=%d outside [1,%d]", C, PN2_MAX_COORD);
    PN2_REQUIRE(K >= 1 && K <= N, "pn2_ball_query_f32: number=%d must be in [1, N=%d] (the reference raises for K > N)", K, N);
    PN2_REQUIRE(K <= 4096, "pn2_ball_query_f32: number=%d > 4096 unsupported", K);
    return launch_ball_query(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, C, r2, K, idx64, idx32, empty_flag,
                             (cudaStream_t)stream);
}

int pn2_ball_query_multi_f32(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                             const float *points, int64_t sb, int64_t sn, int64_t sc,
                             int B, int S, int N, int C, int n_radii, const float *r2, const int *K,
                             int64_t *const *idx64, int32_t *const *idx32, int32_t *empty_flag, void *stream)
{
    PN2_REQUIRE(centres && points && r2 && K, "pn2_ball_query_multi_f32: null pointer");
    PN2_REQUIRE(n_radii >= 1 && n_radii <= 3, "pn2_ball_query_multi_f32: n_radii=%d outside [1,3]", n_radii);
    PN2_REQUIRE(idx64 || idx32, "pn2_ball_query_multi_f32: no output requested");
    PN2_REQUIRE(C >= 1 && C <= PN2_MAX_COORD, "pn2_ball_query_multi_f32: C=%d outside [1,%d]", C, PN2_MAX_COORD);
    for (int r = 0; r < n_radii; ++r) {
        PN2_REQUIRE((idx64 && idx64[r]) || (idx32 && idx32[r]), "pn2_ball_query_multi_f32: no output requested for radius %d", r);
        PN2_REQUIRE(K[r] >= 1 && K[r] <= N, "pn2_ball_query_multi_f32: number=%d must be in [1, N=%d] (the reference raises for K > N)", K[r], N);
        PN2_REQUIRE(K[r] <= 4096, "pn2_ball_query_multi_f32: number=%d > 4096 unsupported", K[r]);
    }
    return launch_ball_query_multi(centres, csb, csn, csc, points, sb, sn, sc, B, S, N, C, n_radii, r2, K, idx64, idx32,
                                   empty_flag, (cudaStream_t)stream);
}

int pn2_group_f32(const float *points, int64_t sb, int64_t sn, int64_t sc,
                  const float *feat, int64_t fb, int64_t fn, int64_t fc,
                  const int32_t *idx32, const float *centres,
                  int B, int N, int S, int K, int C, int D, int order, float *out, void *stream)
{
    PN2_REQUIRE(points && out, "pn2_group_f32: null pointer");
    PN2_REQUIRE(order == PN2_ORDER_ALL || (idx32 && centres), "pn2_group_f32: idx/centres required");
    PN2_REQUIRE(order != PN2_ORDER_ALL || (S == 1 && K == N), "pn2_group_f32: group_all needs S=1, K=N");
    PN2_REQUIRE((D == 0) == (feat == nullptr), "pn2_group_f32: feat pointer / D mismatch");
    return launch_group(points, sb, sn, sc, feat, fb, fn, fc, idx32, centres, B, N, S, K, C, D, order, out, (cudaStream_t)stream);
}

int64_t pn2_packed_bytes(int L, const int *dims, int prec)
{
    if (!valid_dims(L, dims)) { set_error("pn2_packed_bytes: bad layer list"); return -1; }
    if (prec == PN2_PREC_F32) return packed_bytes_f32(L, dims);
    if (prec == PN2_PREC_BF16) return packed_bytes_bf16(L, dims);
    set_error("pn2_packed_bytes: unknown precision %d", prec);
    return -1;
}

int pn2_pack_weights(int L, const int *dims, const void *const *ptrs, float eps, int prec, int C, int order,
                     void *packed, void *stream)
{
    PN2_REQUIRE(valid_dims(L, dims), "pn2_pack_weights: bad layer list");
    PN2_REQUIRE(ptrs && packed, "pn2_pack_weights: null pointer");
    for (int i = 0; i < 6 * L; ++i) PN2_REQUIRE(ptrs[i], "pn2_pack_weights: null parameter pointer %d", i);
    PN2_REQUIRE(C >= 1 && C <= dims[0] && order >= 0 && order <= 2, "pn2_pack_weights: bad C=%d / order=%d", C, order);
    if (prec == PN2_PREC_F32) return launch_pack_f32(L, dims, ptrs, eps, C, order, packed, (cudaStream_t)stream);
    if (prec == PN2_PREC_BF16) return launch_pack_bf16(L, dims, ptrs, eps, C, order, packed, (cudaStream_t)stream);
    PN2_REQUIRE(false, "pn2_pack_weights: unknown precision %d", prec);
}

int pn2_sa_needs_zero_init(int S, int K, int L, const int *dims, int prec)
{
    if (!valid_dims(L, dims)) { set_error("pn2_sa_needs_zero_init: bad layer list"); return -1; }
    return prec == PN2_PREC_BF16 ? sa_bf16_needs_zero(S, K, L, dims) : sa_f32_needs_zero(S, K, L, dims);
}

int pn2_sa_bf16_side_io(int K, int C, int order, int L, const int *dims, int prec)
{
    if (!valid_dims(L, dims)) { set_error("pn2_sa_bf16_side_io: bad layer list"); return -1; }
    if (C < 1 || C > dims[0]) { set_error("pn2_sa_bf16_side_io: bad C"); return -1; }
    return prec == PN2_PREC_BF16 ? sa_bf16_pipelined_covers(K, C, dims[0] - C, order, L, dims) : 0;
}

int pn2_sa_fused(const float *points, int64_t sb, int64_t sn, int64_t sc,
                 const float *feat, int64_t fb, int64_t fn, int64_t fc,
                 const float *centres, const int32_t *idx32,
                 int B, int N, int S, int K, int C, int D, int order,
                 int L, const int *dims, const void *packed, int prec,
                 float *out, int64_t ob, int64_t os, int co_off,
                 const void *feat_bf16, void *out_bf16, void *stream)
{
    PN2_REQUIRE(points && packed && out, "pn2_sa_fused: null pointer");
    PN2_REQUIRE(valid_dims(L, dims), "pn2_sa_fused: bad layer list");
    PN2_REQUIRE(order >= 0 && order <= 2, "pn2_sa_fused: bad order %d", order);
    PN2_REQUIRE(order == PN2_ORDER_ALL || (idx32 && centres), "pn2_sa_fused: idx/centres required");
    PN2_REQUIRE(order != PN2_ORDER_ALL || (S == 1 && K == N), "pn2_sa_fused: group_all needs S=1, K=N");
    // features may be given as the bf16 side tensor alone (bf16 mode; the launcher refuses configurations whose kernel
    // would need the fp32 rows)
    PN2_REQUIRE((D == 0) == (feat == nullptr && !(prec == PN2_PREC_BF16 && feat_bf16)), "pn2_sa_fused: feat pointer / D mismatch");
    PN2_REQUIRE(dims[0] == C + D, "pn2_sa_fused: first layer expects %d channels, rows have %d (+%d)", dims[0], C, D);
    PN2_REQUIRE(B >= 0 && N >= 1 && S >= 1 && K >= 1 && C >= 1, "pn2_sa_fused: bad sizes");
    SaArgs a;
    a.points = points; a.sb = sb; a.sn = sn; a.sc = sc;
    a.feat = feat; a.fb = fb; a.fn = fn; a.fc = fc;
    a.centres = centres; a.idx32 = idx32;
    a.B = B; a.N = N; a.S = S; a.K = K; a.C = C; a.D = D; a.order = order;
    a.L = L;
    for (int l = 0; l <= L; ++l) a.dims[l] = dims[l];
    a.packed = packed; a.out = out; a.ob = ob; a.os = os; a.co_off = co_off;
    a.feat_bf16 = (prec == PN2_PREC_BF16) ? feat_bf16 : nullptr;
    a.out_bf16 = (prec == PN2_PREC_BF16) ? out_bf16 : nullptr;
    if (prec == PN2_PREC_F32) return launch_sa_f32(a, (cudaStream_t)stream);
    if (prec == PN2_PREC_BF16) return launch_sa_bf16(a, (cudaStream_t)stream);
    PN2_REQUIRE(false, "pn2_sa_fused: unknown precision %d", prec);
}

int pn2_normalize_splice_f32(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                             const int64_t *category, int num_category, float *out, void *stream)
{
    PN2_REQUIRE(xyz && out, "pn2_normalize_splice_f32: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1, "pn2_normalize_splice_f32: bad sizes B=%d N=%d", B, N);
    PN2_REQUIRE(num_category >= 0 && num_category <= 1024 && (num_category == 0 || category), "pn2_normalize_splice_f32: bad category arguments");
    return launch_normalize_splice(xyz, sb, sn, sc, B, N, category, num_category, out, (cudaStream_t)stream);
}

int64_t pn2_fps_xyz_scratch_bytes(int B, int N, int elem_bytes)
{
    return (B < 0 || N < 1 || (elem_bytes != 4 && elem_bytes != 8)) ? -1 : fps_np_scratch_bytes(B, N, elem_bytes);
}

int pn2_fps_xyz_f64(const double *points, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                    const int64_t *start, int S, int64_t *idx64, void *scratch, void *stream)
{
    PN2_REQUIRE(points && start && idx64, "pn2_fps_xyz_f64: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1 && S >= 0, "pn2_fps_xyz_f64: bad sizes B=%d N=%d S=%d", B, N, S);
    return launch_fps_np_f64(points, sb, sn, sc, B, N, start, S, idx64, scratch, (cudaStream_t)stream);
}

int pn2_fps_xyz_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                    const int64_t *start, int S, int64_t *idx64, void *scratch, void *stream)
{
    PN2_REQUIRE(points && start && idx64, "pn2_fps_xyz_f32: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1 && S >= 0, "pn2_fps_xyz_f32: bad sizes B=%d N=%d S=%d", B, N, S);
    return launch_fps_np_f32(points, sb, sn, sc, B, N, start, S, idx64, scratch, (cudaStream_t)stream);
}

int64_t pn2_bn_train_ws_bytes(int C) { return C >= 1 ? bn_train_ws_bytes(C) : -1; }

int pn2_bn_relu_train_fwd_f32(const float *y, int64_t M, int C, int K, const float *gamma, const float *beta, float eps,
                              void *ws, float *stats, float *z, int32_t *arg,
                              float *running_mean, float *running_var, const float *conv_bias, float factor, void *stream)
{
    PN2_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "pn2_bn_relu_train_fwd_f32: running_mean and running_var go together");
    PN2_REQUIRE(y && gamma && beta && ws && stats && z, "pn2_bn_relu_train_fwd_f32: null pointer");
    PN2_REQUIRE(M >= 2 && C >= 1, "pn2_bn_relu_train_fwd_f32: batch statistics need more than one value per channel (M=%lld, C=%d)", (long long)M, C);
    PN2_REQUIRE(K == 0 || (K > 0 && M % K == 0 && arg), "pn2_bn_relu_train_fwd_f32: K=%d must divide M=%lld and needs arg", K, (long long)M);
    return launch_bn_fwd(y, M, C, K, gamma, beta, eps, ws, stats, z, arg, running_mean, running_var, conv_bias, factor,
                         (cudaStream_t)stream);
}

int pn2_bn_relu_train_bwd_f32(const float *dz, const float *y, const int32_t *arg, int64_t M, int C, int K,
                              const float *stats, void *ws, float *sums, float *dy, void *stream)
{
    PN2_REQUIRE(dz && y && stats && ws && sums && dy, "pn2_bn_relu_train_bwd_f32: null pointer");
    PN2_REQUIRE(M >= 2 && C >= 1, "pn2_bn_relu_train_bwd_f32: bad sizes M=%lld C=%d", (long long)M, C);
    PN2_REQUIRE(K == 0 || (K > 0 && M % K == 0 && arg), "pn2_bn_relu_train_bwd_f32: K=%d must divide M=%lld and needs arg", K, (long long)M);
    return launch_bn_bwd(dz, y, arg, M, C, K, stats, ws, sums, dy, (cudaStream_t)stream);
}

int pn2_selftest_tma_gather(const void *src_bf16, int64_t rows, int cols, const int32_t *rows4, int n_inst, int col0, void *out, void *stream)
{
    return launch_selftest_tma_gather(src_bf16, rows, cols, rows4, n_inst, col0, out, (cudaStream_t)stream);
}

int pn2_debug_trace(void *buf)
{
    set_debug_trace(buf);
    return 0;
}

int pn2_selftest_umma(int n_cols, int k_dim, int variant, float *max_err, float *dump, void *stream)
{
    PN2_REQUIRE(max_err, "pn2_selftest_umma: null pointer");
    return launch_selftest_umma(n_cols, k_dim, variant, max_err, dump, (cudaStream_t)stream);
}


int pn2_gather_rows_f32(const float *feat, int64_t fb, int64_t fn, int64_t fc, const int32_t *idx32, int B, int N, int64_t M, int D,
                        float *out, void *stream)
{
    PN2_REQUIRE(feat && idx32 && out, "pn2_gather_rows_f32: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1 && M >= 0 && D >= 0 && fc == 1, "pn2_gather_rows_f32: bad sizes / rows must have unit channel stride");
    return launch_gather_rows(feat, fb, fn, idx32, B, N, M, D, out, (cudaStream_t)stream);
}

int pn2_scatter_add_rows_f32(const float *grad, const int32_t *idx32, int B, int N, int64_t M, int D, float *grad_feat, void *stream)
{
    PN2_REQUIRE(grad && idx32 && grad_feat, "pn2_scatter_add_rows_f32: null pointer");
    PN2_REQUIRE(B >= 0 && N >= 1 && M >= 0 && D >= 0, "pn2_scatter_add_rows_f32: bad sizes");
    return launch_scatter_add_rows(grad, idx32, B, N, M, D, grad_feat, (cudaStream_t)stream);
}

// see include/pn2.h
__global__ void upload_i64_kernel(const int64_t *__restrict__ src, int64_t *__restrict__ dst, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
int pn2_upload_i64(const int64_t *host_mapped, int64_t *dst, int64_t n, void *stream)
{
    PN2_REQUIRE(host_mapped && dst && n >= 0, "pn2_upload_i64: bad arguments");
    if (n == 0) return 0;
    upload_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(host_mapped, dst, n);
    return launch_status("pn2_upload_i64");
}
}  // extern "C"
