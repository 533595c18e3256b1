// Training-mode BatchNorm + ReLU of the shared MLP on channels-last rows.
//
// The reference's training step (train_rotation.py:95-123 through pointnet2_utils.py:166-172 / :212-218) evaluates
// relu(bn(conv(x))) on the grouped [B, C, K, S] tensor; on a GPU that is cuDNN's per-channel NCHW BatchNorm plus separate
// bias-add / ReLU / threshold kernels, which take 75 % of the configs[4] step (profiles/r01_j_train_profile_aten.txt: the
// 1x1 convolutions themselves are under 10 %).  Here the grouped rows stay in the [M = B*S*K, C] layout the grouping
// kernel writes, the convolution is a plain GEMM on them, and the rest of the layer is four streaming kernels:
//
//   forward   stats : per-channel sum / sum of squares of y = x W^T (fp64 accumulation)      reads y once
//             apply : z = max(0, y * a + b), a = gamma / sqrt(var + eps), b = beta - mean * a   reads y, writes z
//   backward  stats : s1 = sum g, s2 = sum g * xhat with g = dz * [z > 0], xhat = (y - mean) * invstd   reads dz, y
//             apply : dy = a * (g - s1/M - xhat * s2/M)                                        reads dz, y, writes dy
//
// bn_out and the ReLU mask are recomputed from y (the same fma, so the mask is the forward's bit for bit), so autograd
// keeps y and z per layer instead of y, bn(y) and z.  The "pooled" variants fuse the max over the K rows of a group
// (pointnet2_utils.py:172 / :218) into the last layer: the forward writes only the pooled [G, C] maximum and its
// argmax row, the backward takes the pooled gradient, and z / dz of the widest layer never exist in HBM.
//
// All four are HBM-bound column-wise streams: a thread owns one 16-byte column group and walks rows, so a warp reads
// whole 128-byte lines; per-CTA partial sums go to a workspace and a second tiny kernel adds them in a fixed order
// (deterministic, no atomics).  Biased batch variance for the normalisation, like nn.BatchNorm2d (:155).
#include "pn2_common.cuh"

namespace pn2 {

namespace {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_CTAS = 148 * 4;

// stats buffer rows: [0] mean  [1] biased variance  [2] invstd  [3] a = gamma * invstd  [4] b = beta - mean * a
// sums buffer rows:  [0] s1 = sum g  [1] s2 = sum g * xhat

// ATen's relu: negative -> 0, NaN stays NaN (fmaxf would drop it)
__device__ __forceinline__ float relu_f(float v) { return v < 0.f ? 0.f : v; }

struct ColMap {
    int cv;      // this thread's column group (V floats), < CV
    int rl;      // row lane
    int RL;      // row lanes per CTA
    int per;     // column groups per CTA
    bool live;
};

template <int V>
__device__ __forceinline__ ColMap col_map(int C)
{
    const int CV = C / V;
    ColMap m;
    const int per = CV < BN_THREADS ? CV : BN_THREADS;     // column groups handled per CTA pass (grid.y tiles the rest)
    m.RL = BN_THREADS / per;
    m.per = per;
    const int t = threadIdx.x;
    m.rl = t / per;
    m.cv = blockIdx.y * per + (t % per);
    m.live = m.rl < m.RL && m.cv < CV;
    return m;
}

template <int V> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<1> { using T = float; };

template <int V> __device__ __forceinline__ void ldv(const float *p, float (&v)[V]);
template <> __device__ __forceinline__ void ldv<4>(const float *p, float (&v)[4])
{
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<1>(const float *p, float (&v)[1]) { v[0] = __ldg(p); }

template <int V> __device__ __forceinline__ void stv(float *p, const float (&v)[V]);
template <> __device__ __forceinline__ void stv<4>(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ void stv<1>(float *p, const float (&v)[1]) { *p = v[0]; }

// rows [r0, r1) of this CTA
__device__ __forceinline__ void row_range(int64_t M, int64_t &r0, int64_t &r1)
{
    const int64_t per = (M + gridDim.x - 1) / gridDim.x;
    r0 = (int64_t)blockIdx.x * per;
    r1 = r0 + per < M ? r0 + per : M;
}

// Adds the row lanes of a CTA (shared memory) and writes the CTA's partials: ws[(q * gridDim.x + cta) * C + c].
template <int V, int Q>
__device__ __forceinline__ void cta_partials(const ColMap &m, int C, double (&acc)[Q][V], double *__restrict__ ws)
{
    extern __shared__ double s_red[];                       // [Q][RL][per * V]
    const int per = m.per;
    const int lc = (m.cv - blockIdx.y * per) * V;
    if (m.live) {
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int j = 0; j < V; ++j) s_red[(q * m.RL + m.rl) * per * V + lc + j] = acc[q][j];
    }
    __syncthreads();
    if (m.live && m.rl == 0) {
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int j = 0; j < V; ++j) {
                double s = 0.0;
                for (int r = 0; r < m.RL; ++r) s += s_red[(q * m.RL + r) * per * V + lc + j];
                ws[((int64_t)q * gridDim.x + blockIdx.x) * C + m.cv * V + j] = s;
            }
    }
}

template <int V>
__global__ void __launch_bounds__(BN_THREADS) bn_fwd_stats_kernel(const float *__restrict__ y, int64_t M, int C, double *__restrict__ ws)
{
    const ColMap m = col_map<V>(C);
    double acc[2][V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] = acc[1][j] = 0.0;
    int64_t r0, r1;
    row_range(M, r0, r1);
    if (m.live) {
        const float *p = y + (int64_t)m.cv * V;
        int64_t r = r0 + m.rl;
        for (; r + 3 * m.RL < r1; r += 4 * m.RL) {          // four independent loads in flight
            float v[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) ldv<V>(p + (r + u * m.RL) * C, v[u]);
            // four rows are first added in fp32 (sum and sum of squares: four-term partials, error far below the rounding
            // of y itself), then once into the fp64 accumulators: the fp64 pipe would otherwise co-limit this pure-read kernel
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float s4 = (v[0][j] + v[1][j]) + (v[2][j] + v[3][j]);
                const float q4 = fmaf(v[3][j], v[3][j], fmaf(v[2][j], v[2][j], fmaf(v[1][j], v[1][j], v[0][j] * v[0][j])));
                acc[0][j] += (double)s4;
                acc[1][j] += (double)q4;
            }
        }
        for (; r < r1; r += m.RL) {
            float v[V];
            ldv<V>(p + r * C, v);
#pragma unroll
            for (int j = 0; j < V; ++j) { const double d = v[j]; acc[0][j] += d; acc[1][j] = fma(d, d, acc[1][j]); }
        }
    }
    cta_partials<V, 2>(m, C, acc, ws);
}

// Sum of the per-CTA partials of FIN_CH channels by a (FIN_CH, FIN_LANES) thread block: lane ly adds partials ly,
// ly + FIN_LANES, ... (64-byte segments), lane 0 then adds the FIN_LANES sub-sums in order -- a fixed order, so deterministic.
constexpr int FIN_LANES = 64, FIN_CH = 8;

__device__ __forceinline__ bool finalize_sums(const double *__restrict__ ws, int nctas, int C, double &s, double &q)
{
    __shared__ double s_part[2][FIN_LANES][FIN_CH];
    const int c = blockIdx.x * FIN_CH + threadIdx.x, ly = threadIdx.y;
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int i = ly; i < nctas; i += FIN_LANES) {
            a0 += ws[(int64_t)i * C + c];
            a1 += ws[((int64_t)nctas + i) * C + c];
        }
    }
    s_part[0][ly][threadIdx.x] = a0;
    s_part[1][ly][threadIdx.x] = a1;
    __syncthreads();
    if (ly != 0 || c >= C) return false;
    s = q = 0.0;
#pragma unroll 8
    for (int l = 0; l < FIN_LANES; ++l) { s += s_part[0][l][threadIdx.x]; q += s_part[1][l][threadIdx.x]; }
    return true;
}

// With running_mean / running_var it also does nn.BatchNorm2d's bookkeeping: running <- (1 - f) running + f batch, the
// mean shifted by the convolution bias that was left out of y, the variance unbiased (M / (M - 1)) like ATen.
__global__ void __launch_bounds__(FIN_CH * FIN_LANES) bn_fwd_finalize_kernel(const double *__restrict__ ws, int nctas, int64_t M, int C,
                                       const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                                       float *__restrict__ stats, float *__restrict__ running_mean,
                                       float *__restrict__ running_var, const float *__restrict__ conv_bias, float factor)
{
    double s, q;
    if (!finalize_sums(ws, nctas, C, s, q)) return;
    const int c = blockIdx.x * FIN_CH + threadIdx.x;
    const double mean = s / (double)M;
    double var = q / (double)M - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = gamma[c] * invstd;
    stats[c] = (float)mean;
    stats[C + c] = (float)var;
    stats[2 * C + c] = invstd;
    stats[3 * C + c] = a;
    stats[4 * C + c] = fmaf(-(float)mean, a, beta[c]);
    if (running_mean) {
        const float mb = (float)mean + (conv_bias ? conv_bias[c] : 0.f);
        running_mean[c] = fmaf(factor, mb, (1.f - factor) * running_mean[c]);
        running_var[c] = fmaf(factor, (float)(var * ((double)M / (double)(M - 1))), (1.f - factor) * running_var[c]);
    }
}

// z = max(0, y * a + b).  POOLED: rows come in groups of K; only max over the group and its first argmax row are kept.
template <int V, bool POOLED>
__global__ void __launch_bounds__(BN_THREADS) bn_fwd_apply_kernel(const float *__restrict__ y, int64_t M, int C, int K,
                                                                  const float *__restrict__ stats, float *__restrict__ z,
                                                                  int32_t *__restrict__ arg)
{
    const ColMap m = col_map<V>(C);
    if (!m.live) return;
    float a[V], b[V];
    ldv<V>(stats + 3 * C + m.cv * V, a);
    ldv<V>(stats + 4 * C + m.cv * V, b);
    const float *p = y + (int64_t)m.cv * V;
    if (!POOLED) {
        int64_t r0, r1;
        row_range(M, r0, r1);
        float *o = z + (int64_t)m.cv * V;
#pragma unroll 4
        for (int64_t r = r0 + m.rl; r < r1; r += m.RL) {
            float v[V];
            ldv<V>(p + r * C, v);
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = relu_f(fmaf(v[j], a[j], b[j]));
            stv<V>(o + r * C, v);
        }
    } else {
        const int64_t G = M / K;
        for (int64_t g = (int64_t)blockIdx.x * m.RL + m.rl; g < G; g += (int64_t)gridDim.x * m.RL) {
            float best[V];
            int bi[V];
#pragma unroll
            for (int j = 0; j < V; ++j) { best[j] = -1.f; bi[j] = 0; }
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                float v[V];
                ldv<V>(p + (g * K + k) * C, v);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float t = relu_f(fmaf(v[j], a[j], b[j]));
                    // torch.max: the first maximal row; NaN propagates (a NaN is "greater" than everything before it)
                    if (t > best[j] || (t != t && best[j] == best[j])) { best[j] = t; bi[j] = k; }
                }
            }
            stv<V>(z + g * C + (int64_t)m.cv * V, best);
#pragma unroll
            for (int j = 0; j < V; ++j) arg[g * C + m.cv * V + j] = bi[j];
        }
    }
}

template <int V, bool POOLED>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_stats_kernel(const float *__restrict__ dz, const float *__restrict__ y,
                                                                  const int32_t *__restrict__ arg, int64_t M, int C, int K,
                                                                  const float *__restrict__ stats, double *__restrict__ ws)
{
    const ColMap m = col_map<V>(C);
    double acc[2][V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] = acc[1][j] = 0.0;
    if (m.live) {
        float mean[V], istd[V], a[V], b[V];
        ldv<V>(stats + m.cv * V, mean);
        ldv<V>(stats + 2 * C + m.cv * V, istd);
        ldv<V>(stats + 3 * C + m.cv * V, a);
        ldv<V>(stats + 4 * C + m.cv * V, b);
        const float *py = y + (int64_t)m.cv * V;
        if (!POOLED) {
            int64_t r0, r1;
            row_range(M, r0, r1);
            const float *pg = dz + (int64_t)m.cv * V;
#pragma unroll 4
            for (int64_t r = r0 + m.rl; r < r1; r += m.RL) {
                float v[V], g[V];
                ldv<V>(py + r * C, v);
                ldv<V>(pg + r * C, g);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float gg = fmaf(v[j], a[j], b[j]) > 0.f ? g[j] : 0.f;
                    const float xh = (v[j] - mean[j]) * istd[j];
                    acc[0][j] += (double)gg;
                    acc[1][j] = fma((double)gg, (double)xh, acc[1][j]);
                }
            }
        } else {
            // only the argmax row of each group carries gradient: one y element per (group, channel)
            const int64_t G = M / K;
            for (int64_t g = (int64_t)blockIdx.x * m.RL + m.rl; g < G; g += (int64_t)gridDim.x * m.RL) {
                float gr[V];
                ldv<V>(dz + g * C + (int64_t)m.cv * V, gr);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const int k = arg[g * C + m.cv * V + j];
                    const float v = __ldg(py + (g * K + k) * C + j);
                    const float gg = fmaf(v, a[j], b[j]) > 0.f ? gr[j] : 0.f;
                    const float xh = (v - mean[j]) * istd[j];
                    acc[0][j] += (double)gg;
                    acc[1][j] = fma((double)gg, (double)xh, acc[1][j]);
                }
            }
        }
    }
    cta_partials<V, 2>(m, C, acc, ws);
}

__global__ void __launch_bounds__(FIN_CH * FIN_LANES) bn_bwd_finalize_kernel(const double *__restrict__ ws, int nctas, int C, float *__restrict__ sums)
{
    double s1, s2;
    if (!finalize_sums(ws, nctas, C, s1, s2)) return;
    const int c = blockIdx.x * FIN_CH + threadIdx.x;
    sums[c] = (float)s1;
    sums[C + c] = (float)s2;
}

template <int V, bool POOLED>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const float *__restrict__ dz, const float *__restrict__ y,
                                                                  const int32_t *__restrict__ arg, int64_t M, int C, int K,
                                                                  const float *__restrict__ stats, const float *__restrict__ sums,
                                                                  float *__restrict__ dy)
{
    const ColMap m = col_map<V>(C);
    if (!m.live) return;
    float mean[V], istd[V], a[V], b[V], k1[V], k2[V];
    ldv<V>(stats + m.cv * V, mean);
    ldv<V>(stats + 2 * C + m.cv * V, istd);
    ldv<V>(stats + 3 * C + m.cv * V, a);
    ldv<V>(stats + 4 * C + m.cv * V, b);
    ldv<V>(sums + m.cv * V, k1);
    ldv<V>(sums + C + m.cv * V, k2);
    const float inv_m = 1.f / (float)M;
#pragma unroll
    for (int j = 0; j < V; ++j) { k1[j] *= inv_m; k2[j] *= inv_m; }
    const float *py = y + (int64_t)m.cv * V;
    float *po = dy + (int64_t)m.cv * V;
    if (!POOLED) {
        int64_t r0, r1;
        row_range(M, r0, r1);
        const float *pg = dz + (int64_t)m.cv * V;
#pragma unroll 4
        for (int64_t r = r0 + m.rl; r < r1; r += m.RL) {
            float v[V], g[V];
            ldv<V>(py + r * C, v);
            ldv<V>(pg + r * C, g);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float gg = fmaf(v[j], a[j], b[j]) > 0.f ? g[j] : 0.f;
                const float xh = (v[j] - mean[j]) * istd[j];
                v[j] = a[j] * (gg - k1[j] - xh * k2[j]);
            }
            stv<V>(po + r * C, v);
        }
    } else {
        const int64_t G = M / K;
        for (int64_t g = (int64_t)blockIdx.x * m.RL + m.rl; g < G; g += (int64_t)gridDim.x * m.RL) {
            float gr[V];
            int ak[V];
            ldv<V>(dz + g * C + (int64_t)m.cv * V, gr);
#pragma unroll
            for (int j = 0; j < V; ++j) ak[j] = arg[g * C + m.cv * V + j];
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                float v[V];
                ldv<V>(py + (g * K + k) * C, v);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float gg = (k == ak[j] && fmaf(v[j], a[j], b[j]) > 0.f) ? gr[j] : 0.f;
                    const float xh = (v[j] - mean[j]) * istd[j];
                    v[j] = a[j] * (gg - k1[j] - xh * k2[j]);
                }
                stv<V>(po + (g * K + k) * C, v);
            }
        }
    }
}

struct Grid {
    dim3 grid;
    size_t smem;
    int V;
};

Grid plan(int64_t rows, int C)
{
    Grid g;
    g.V = (C % 4 == 0) ? 4 : 1;
    const int CV = C / g.V;
    const int per = CV < BN_THREADS ? CV : BN_THREADS;
    const int RL = BN_THREADS / per;
    int64_t want = (rows + RL - 1) / RL;                    // at least one row per lane
    const int ytiles = (CV + per - 1) / per;
    int64_t cap = BN_MAX_CTAS / ytiles;
    if (cap < 1) cap = 1;
    g.grid = dim3((unsigned)(want < cap ? (want < 1 ? 1 : want) : cap), (unsigned)ytiles, 1);
    g.smem = (size_t)2 * RL * per * g.V * sizeof(double);
    return g;
}

}  // namespace

int64_t bn_train_ws_bytes(int C) { return (int64_t)2 * BN_MAX_CTAS * C * sizeof(double); }

int launch_bn_fwd(const float *y, int64_t M, int C, int K, const float *gamma, const float *beta, float eps,
                  void *ws, float *stats, float *z, int32_t *arg,
                  float *running_mean, float *running_var, const float *conv_bias, float factor, cudaStream_t st)
{
    const Grid g = plan(M, C);
    double *w = reinterpret_cast<double *>(ws);
    if (g.V == 4) bn_fwd_stats_kernel<4><<<g.grid, BN_THREADS, g.smem, st>>>(y, M, C, w);
    else bn_fwd_stats_kernel<1><<<g.grid, BN_THREADS, g.smem, st>>>(y, M, C, w);
    bn_fwd_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, dim3(FIN_CH, FIN_LANES), 0, st>>>(w, (int)g.grid.x, M, C, gamma, beta, eps, stats,
                                                                               running_mean, running_var, conv_bias, factor);
    if (K <= 0) {
        if (g.V == 4) bn_fwd_apply_kernel<4, false><<<g.grid, BN_THREADS, 0, st>>>(y, M, C, 1, stats, z, nullptr);
        else bn_fwd_apply_kernel<1, false><<<g.grid, BN_THREADS, 0, st>>>(y, M, C, 1, stats, z, nullptr);
    } else {
        const Grid p = plan(M / K, C);
        if (g.V == 4) bn_fwd_apply_kernel<4, true><<<p.grid, BN_THREADS, 0, st>>>(y, M, C, K, stats, z, arg);
        else bn_fwd_apply_kernel<1, true><<<p.grid, BN_THREADS, 0, st>>>(y, M, C, K, stats, z, arg);
    }
    return launch_status("bn_fwd");
}

int launch_bn_bwd(const float *dz, const float *y, const int32_t *arg, int64_t M, int C, int K, const float *stats,
                  void *ws, float *sums, float *dy, cudaStream_t st)
{
    double *w = reinterpret_cast<double *>(ws);
    if (K <= 0) {
        const Grid g = plan(M, C);
        if (g.V == 4) bn_bwd_stats_kernel<4, false><<<g.grid, BN_THREADS, g.smem, st>>>(dz, y, nullptr, M, C, 1, stats, w);
        else bn_bwd_stats_kernel<1, false><<<g.grid, BN_THREADS, g.smem, st>>>(dz, y, nullptr, M, C, 1, stats, w);
        bn_bwd_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, dim3(FIN_CH, FIN_LANES), 0, st>>>(w, (int)g.grid.x, C, sums);
        if (g.V == 4) bn_bwd_apply_kernel<4, false><<<g.grid, BN_THREADS, 0, st>>>(dz, y, nullptr, M, C, 1, stats, sums, dy);
        else bn_bwd_apply_kernel<1, false><<<g.grid, BN_THREADS, 0, st>>>(dz, y, nullptr, M, C, 1, stats, sums, dy);
    } else {
        const Grid p = plan(M / K, C);
        if (p.V == 4) bn_bwd_stats_kernel<4, true><<<p.grid, BN_THREADS, p.smem, st>>>(dz, y, arg, M, C, K, stats, w);
        else bn_bwd_stats_kernel<1, true><<<p.grid, BN_THREADS, p.smem, st>>>(dz, y, arg, M, C, K, stats, w);
        bn_bwd_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, dim3(FIN_CH, FIN_LANES), 0, st>>>(w, (int)p.grid.x, C, sums);
        if (p.V == 4) bn_bwd_apply_kernel<4, true><<<p.grid, BN_THREADS, 0, st>>>(dz, y, arg, M, C, K, stats, sums, dy);
        else bn_bwd_apply_kernel<1, true><<<p.grid, BN_THREADS, 0, st>>>(dz, y, arg, M, C, K, stats, sums, dy);
    }
    return launch_status("bn_bwd");
}

}  // namespace pn2
