/*
 * pn2_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference path).
 *
 * Plain-C restatement of the PointNet++ set-abstraction arithmetic of the reference's
 * model/pointnet2_utils.py, written from the behaviour of that file (and of the ATen / MKL
 * CPU kernels it calls) so that the CUDA kernels can be diffed bit-for-bit on indices and
 * within a stated tolerance on features.  Nothing in the product imports or links this file:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity pin: every function here is checked against outputs of the *live* reference
 * (imported from /root/reference/model in the build container) by oracle/make_golden.py; the
 * resulting vectors are committed under tests/golden/ and re-checked by the CPU test-suite.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fPIC -shared (see oracle/Makefile).  -ffp-contract=off
 * is REQUIRED: the contract below distinguishes fused from unfused multiply-adds.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PN2O_MAX_C 16 /* the summation-order rules below were verified against ATen for C <= 16 */

/* ---------------------------------------------------------------------------------------------
 * Summation order of torch.sum(v, -1) over the C coordinate channels (pointnet2_utils.py:24-25,
 * :64).  ATen picks the order from the memory layout of the [B,.,C] operand:
 *   stride(-1) == 1, C <  8 : 4 interleaved partial sums, tail into p0, ((p0+p1)+p2)+p3   ("ILP-4").
 *   stride(-1) == 1, C >= 8 : 8 vector lanes over the first 8*floor(C/8) terms, then
 *                             acc = tail terms in order, then += lane0..lane7.
 *   stride(-1) != 1 (the [B,C,N].permute(0,2,1) view the first layer receives: unit stride along
 *   the point index n): ATen vectorises over n, 32 points per step, one accumulator per point in
 *   ascending k; the points left over, n >= 32*floor(N/32), go through the scalar ILP-4 loop.
 * `n` / `n_total` locate the point inside its cloud for the strided rule (ignored otherwise).
 * ------------------------------------------------------------------------------------------- */
static float sum_ilp4(const float *v, int C)
{
    float p[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int n4 = C / 4;
    for (int i = 0; i < n4; ++i)
        for (int k = 0; k < 4; ++k) p[k] = p[k] + v[4 * i + k];
    for (int t = 4 * n4; t < C; ++t) p[0] = p[0] + v[t];
    return ((p[0] + p[1]) + p[2]) + p[3];
}

static float sum_order(const float *v, int C, int contiguous, int64_t n, int64_t n_total)
{
    if (!contiguous) {
        if (n >= (n_total / 32) * 32) return sum_ilp4(v, C);
        float a = 0.0f;
        for (int k = 0; k < C; ++k) a = a + v[k];
        return a;
    }
    if (C < 8) return sum_ilp4(v, C);
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n8 = C / 8;
    for (int i = 0; i < n8; ++i)
        for (int j = 0; j < 8; ++j) lane[j] = lane[j] + v[8 * i + j];
    float a = 0.0f;
    for (int t = 8 * n8; t < C; ++t) a = a + v[t];
    for (int j = 0; j < 8; ++j) a = a + lane[j];
    return a;
}

/* mode 0: strided, vectorised body; 1: contiguous; 2: strided, leftover points */
float pn2o_sum_order(const float *v, int C, int mode) { return sum_order(v, C, mode == 1, mode == 2 ? 1 : 0, mode == 2 ? 1 : 32); }

/* ---------------------------------------------------------------------------------------------
 * farthest_point_sample (pointnet2_utils.py:47-68).
 * pts is [B,N,C] addressed with element strides (sb,sn,sc); start[b] is the random first index the
 * reference draws at :59.  Per iteration: out[i]=far; d_n = sum_order((p_n-p_far)^2) with sub, mul,
 * add each rounded (no FMA); dist_n takes d_n only when d_n < dist_n (:65-66, NaN never replaces);
 * far = argmax dist, lowest index on ties (:67).
 * ------------------------------------------------------------------------------------------- */
int pn2o_fps(const float *pts, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
             const int64_t *start, int S, int64_t *out)
{
    if (C < 1 || C > PN2O_MAX_C || N < 1 || S < 0) return -1;
    int contiguous = (sc == 1);
    float *dist = (float *)malloc(sizeof(float) * (size_t)N);
    if (!dist) return -2;
    for (int b = 0; b < B; ++b) {
        const float *P = pts + (int64_t)b * sb;
        for (int n = 0; n < N; ++n) dist[n] = 1e10f;
        int64_t far = start[b];
        if (far < 0 || far >= N) { free(dist); return -3; }
        for (int i = 0; i < S; ++i) {
            out[(int64_t)b * S + i] = far;
            float c[PN2O_MAX_C], t[PN2O_MAX_C];
            for (int k = 0; k < C; ++k) c[k] = P[far * sn + k * sc];
            float best = -1.0f;
            int64_t besti = 0;
            for (int n = 0; n < N; ++n) {
                for (int k = 0; k < C; ++k) {
                    float df = P[(int64_t)n * sn + k * sc] - c[k];
                    t[k] = df * df;
                }
                float d = sum_order(t, C, contiguous, n, N);
                if (d < dist[n]) dist[n] = d;
                if (dist[n] > best) { best = dist[n]; besti = n; }
            }
            far = besti;
        }
    }
    free(dist);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * The numpy copies of farthest point sampling (provider.py:183-204, data_utils/ModelDataLoader.py:10-31,
 * point_collect/transformer.py:120-141): one cloud, first three columns of an [N,C] array, arithmetic in the
 * array's dtype.  numpy evaluates `xyz - centroid` (:196), `** 2` and `np.sum(., -1)` as three array passes, so
 * every sub, square and add is rounded on its own, and a 3-term contiguous sum is sequential (probed: 200 000
 * random triples, float64 and float32).  `distance` is float64 in the reference (:191) even for float32 input; a
 * float32 value widens exactly, so keeping it in the array's dtype decides `dist < distance` (:197) and
 * np.argmax (:199, first of equal maxima) identically.
 * ------------------------------------------------------------------------------------------- */
#define PN2O_FPS_XYZ(NAME, T)                                                                       \
    int NAME(const T *pts, int64_t sn, int64_t sc, int N, int64_t start, int S, int64_t *out)       \
    {                                                                                               \
        if (N < 1 || S < 0 || start < 0 || start >= N) return -1;                                   \
        T *dist = (T *)malloc(sizeof(T) * (size_t)N);                                               \
        if (!dist) return -2;                                                                       \
        for (int n = 0; n < N; ++n) dist[n] = (T)1e10;                                              \
        int64_t far = start;                                                                        \
        for (int i = 0; i < S; ++i) {                                                               \
            out[i] = far;                                                                           \
            const T cx = pts[far * sn], cy = pts[far * sn + sc], cz = pts[far * sn + 2 * sc];       \
            T best = (T)-1;                                                                         \
            int64_t besti = 0;                                                                      \
            for (int n = 0; n < N; ++n) {                                                           \
                volatile T dx = pts[(int64_t)n * sn] - cx;                                          \
                volatile T dy = pts[(int64_t)n * sn + sc] - cy;                                     \
                volatile T dz = pts[(int64_t)n * sn + 2 * sc] - cz;                                 \
                volatile T qx = dx * dx, qy = dy * dy, qz = dz * dz;                                \
                volatile T s = qx + qy;                                                             \
                T d = s + qz;                                                                       \
                if (d < dist[n]) dist[n] = d;                                                       \
                if (dist[n] > best) { best = dist[n]; besti = n; }                                  \
            }                                                                                       \
            far = besti;                                                                            \
        }                                                                                           \
        free(dist);                                                                                 \
        return 0;                                                                                   \
    }
PN2O_FPS_XYZ(pn2o_fps_xyz_f64, double)
PN2O_FPS_XYZ(pn2o_fps_xyz_f32, float)

/* ---------------------------------------------------------------------------------------------
 * square_distance as the reference evaluates it (pointnet2_utils.py:23-25): expanded form
 *   dot = s0*p0; dot = fma(s_k, p_k, dot) k=1..C-1   (what MKL sgemm does for this K; :23)
 *   d = -2*dot; d += sum_order(s^2) [src always contiguous]; d += sum_order(p^2) [dst layout]
 * src = centres [B,S,C] contiguous; dst = pts with strides.  out [B,S,N].
 * ------------------------------------------------------------------------------------------- */
static void sq_norms(const float *P, int64_t sn, int64_t sc, int N, int C, int contiguous, float *out)
{
    float t[PN2O_MAX_C];
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < C; ++k) {
            float x = P[(int64_t)n * sn + k * sc];
            t[k] = x * x;
        }
        out[n] = sum_order(t, C, contiguous, n, N);
    }
}

static inline float pair_d(const float *s, const float *P, int64_t pn_off, int64_t sc, int C, float s2, float p2)
{
    float dot = s[0] * P[pn_off];
    for (int k = 1; k < C; ++k) dot = fmaf(s[k], P[pn_off + k * sc], dot);
    float d = -2.0f * dot;
    d = d + s2;
    d = d + p2;
    return d;
}

int pn2o_square_distance(const float *src, const float *pts, int64_t sb, int64_t sn, int64_t sc,
                         int B, int S, int N, int C, float *out)
{
    if (C < 1 || C > PN2O_MAX_C) return -1;
    float *p2 = (float *)malloc(sizeof(float) * (size_t)N);
    float *s2 = (float *)malloc(sizeof(float) * (size_t)(S > 0 ? S : 1));
    if (!p2 || !s2) { free(p2); free(s2); return -2; }
    for (int b = 0; b < B; ++b) {
        const float *P = pts + (int64_t)b * sb;
        const float *Sx = src + (int64_t)b * S * C;
        sq_norms(P, sn, sc, N, C, sc == 1, p2);
        sq_norms(Sx, C, 1, S, C, 1, s2);
        for (int s = 0; s < S; ++s)
            for (int n = 0; n < N; ++n)
                out[((int64_t)b * S + s) * N + n] = pair_d(Sx + (int64_t)s * C, P, (int64_t)n * sn, sc, C, s2[s], p2[n]);
    }
    free(p2); free(s2);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * query_ball_point (pointnet2_utils.py:70-90): the first K indices n (ascending) with
 * !(d(s,n) > r2) where r2 = (float)(radius**2) computed in double by Python (:85); remaining slots
 * repeat the first hit (:87-89).  Returns 1 if some centre has no in-radius point (the reference
 * would then index out of range at the following index_points); slots are then filled with N.
 * ------------------------------------------------------------------------------------------- */
int pn2o_ball_query(const float *centres, const float *pts, int64_t sb, int64_t sn, int64_t sc,
                    int B, int S, int N, int C, float r2, int K, int64_t *out)
{
    if (C < 1 || C > PN2O_MAX_C || K < 1 || K > N) return -1;
    float *p2 = (float *)malloc(sizeof(float) * (size_t)N);
    if (!p2) return -2;
    int empty = 0;
    float t[PN2O_MAX_C];
    for (int b = 0; b < B; ++b) {
        const float *P = pts + (int64_t)b * sb;
        sq_norms(P, sn, sc, N, C, sc == 1, p2);
        for (int s = 0; s < S; ++s) {
            const float *sx = centres + ((int64_t)b * S + s) * C;
            for (int k = 0; k < C; ++k) t[k] = sx[k] * sx[k];
            float s2 = sum_order(t, C, 1, 0, 0);
            int64_t *row = out + ((int64_t)b * S + s) * K;
            int cnt = 0;
            for (int n = 0; n < N && cnt < K; ++n) {
                float d = pair_d(sx, P, (int64_t)n * sn, sc, C, s2, p2[n]);
                if (!(d > r2)) row[cnt++] = n;
            }
            if (cnt == 0) { empty = 1; for (int j = 0; j < K; ++j) row[j] = N; }
            else for (int j = cnt; j < K; ++j) row[j] = row[0];
        }
    }
    free(p2);
    return empty;
}

/* ---------------------------------------------------------------------------------------------
 * Grouping (pointnet2_utils.py:109-116 SSG, :205-211 MSG, :134-140 group_all).
 * order: 0 = SSG  [xyz_norm(C), feat(D)], 1 = MSG [feat(D), xyz_norm(C)], 2 = group_all [xyz_raw, feat]
 * pts strides (sb,sn,sc); feat [B,N,D] strides (fb,fn,fc) or NULL (D=0).
 * idx [B,S,K] int64; centres [B,S,C] contiguous.  out [B,S,K,C+D] contiguous.
 * ------------------------------------------------------------------------------------------- */
int pn2o_group(const float *pts, int64_t sb, int64_t sn, int64_t sc,
               const float *feat, int64_t fb, int64_t fn, int64_t fc,
               const int64_t *idx, const float *centres,
               int B, int N, int S, int K, int C, int D, int order, float *out)
{
    int W = C + D;
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < S; ++s)
            for (int j = 0; j < K; ++j) {
                int64_t n = (order == 2) ? j : idx[((int64_t)b * S + s) * K + j];
                if (n < 0 || n >= N) return -3;
                float *row = out + (((int64_t)b * S + s) * K + j) * W;
                float *xo = (order == 1) ? row + D : row;
                float *fo = (order == 1) ? row : row + C;
                for (int k = 0; k < C; ++k) {
                    float x = pts[b * sb + n * sn + k * sc];
                    xo[k] = (order == 2) ? x : x - centres[((int64_t)b * S + s) * C + k];
                }
                for (int k = 0; k < D; ++k) fo[k] = feat[b * fb + n * fn + k * fc];
            }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Shared MLP + eval-mode BatchNorm + ReLU, then max over the K rows of each group
 * (pointnet2_utils.py:167-172, :213-218).  rows [G,K,Cin] contiguous.  Per layer l (L layers):
 *   y = W_l x + b_l ; y = (y - mean)/sqrt(var + eps)*gamma + beta ; x' = max(y, 0)
 * Accumulated in double, each layer's activations rounded to fp32 (the reference holds fp32
 * tensors between layers; its own accumulation order is oneDNN's and unspecified, hence the
 * tolerance-based comparison).  params: for each layer, W[Cout*Cin], b, gamma, beta, mean, var
 * concatenated (all fp32).  dims[L+1] = Cin, C1, ..., CL.  out [G, CL].
 * ------------------------------------------------------------------------------------------- */
int pn2o_mlp_max(const float *rows, int64_t G, int K, int L, const int *dims, const float *params,
                 float eps, float *out)
{
    int maxc = 0;
    for (int l = 0; l <= L; ++l) if (dims[l] > maxc) maxc = dims[l];
    float *a = (float *)malloc(sizeof(float) * (size_t)maxc);
    float *bb = (float *)malloc(sizeof(float) * (size_t)maxc);
    if (!a || !bb) { free(a); free(bb); return -2; }
    int CL = dims[L];
    for (int64_t g = 0; g < G; ++g) {
        float *o = out + g * CL;
        for (int j = 0; j < K; ++j) {
            memcpy(a, rows + (g * K + j) * dims[0], sizeof(float) * (size_t)dims[0]);
            const float *p = params;
            for (int l = 0; l < L; ++l) {
                int ci = dims[l], co = dims[l + 1];
                const float *W = p, *bias = W + (size_t)co * ci, *gam = bias + co, *bet = gam + co,
                            *mu = bet + co, *var = mu + co;
                for (int c = 0; c < co; ++c) {
                    double acc = 0.0;
                    for (int k = 0; k < ci; ++k) acc += (double)W[(size_t)c * ci + k] * (double)a[k];
                    double y = acc + (double)bias[c];
                    y = (y - (double)mu[c]) / sqrt((double)var[c] + (double)eps) * (double)gam[c] + (double)bet[c];
                    bb[c] = (float)(y > 0.0 ? y : 0.0);
                }
                float *tmp = a; a = bb; bb = tmp;
                p = var + co;
            }
            if (j == 0) memcpy(o, a, sizeof(float) * (size_t)CL);
            else for (int c = 0; c < CL; ++c) if (a[c] > o[c]) o[c] = a[c];
        }
    }
    free(a); free(bb);
    return 0;
}

int pn2o_version(void) { return 1; }
