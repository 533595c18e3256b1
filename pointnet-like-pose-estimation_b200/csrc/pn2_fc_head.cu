// Fused FC head of the PointNet++ / PointNet-v1 models, eval mode:
//     x [B,K0] -> relu(bn1(fc1 x)) -> relu(bn2(fc2 .)) -> fc3 -> log_softmax + argmax | sigmoid + sign | + mean | identity
// (pointnet2_cls_ssg.py:31-38, translation_ssg.py:31-33,39-44, sign_ssg.py:29-36, rotation_ssg.py:32-38, width.py:39-44).
// The reference spends ~14 library launches here (three GEMMs, two BatchNorms, bias adds, ReLUs, softmax, max); this is
// one launch.  BatchNorm1d is folded into the weights by the caller (W' = W s, b' = (b - mu) s + beta), dropout is the
// identity in eval mode.
//
// Work split: a tile of TB clouds is served by a CLUSTER of 8 CTAs.  Every CTA stages the tile's TB x K0 inputs in shared
// memory and computes one eighth of each hidden layer's outputs for all TB clouds (fp32 FFMA, weights streamed from L2
// once per tile: [K/4][H][4] packing makes a warp's weight load one 512-byte line and the activations a shared-memory
// broadcast), then writes its slice of the hidden activations into the shared memory of all eight CTAs through
// distributed shared memory; one cluster barrier per layer.  The last layer (<= 16 outputs) and the epilogue are done by
// every CTA for TB/8 of the clouds.  B = 24: 24 CTAs of ~3 us instead of ~14 launches; B = 2048: the 2.6 MB of weights
// are read 128 times from L2 (0.34 GB) instead of the activations making 14 round trips through HBM.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include "pn2_common.cuh"

namespace cg = cooperative_groups;

namespace pn2 {

constexpr int FC_CS = 8;          // CTAs per cluster
constexpr int FC_THREADS = 256;

struct FcHeadParams {
    const float *x; int64_t xs;
    int B, K0, H1, H2, OUT, kind;
    const float *w1p, *b1, *w2p, *b2, *w3, *b3;
    const float *mean; int64_t ms; const float *mean_mlp;
    float *out; void *aux;
    int alias_h2;
};

// one eighth of a hidden layer for TB clouds: in [TB][K] (shared), weights wp [K/4][H][4] (global), this CTA's outputs
// [rank*OPC, (rank+1)*OPC); result (bias + ReLU applied) broadcast into dst [TB][H] of every CTA of the cluster
template <int TB>
__device__ __forceinline__ void hidden_layer(cg::cluster_group &cluster, const float *in, int K, int H, const float *__restrict__ wp,
                                             const float *__restrict__ bias, float *scratch, float *dst, int rank)
{
    const int tid = threadIdx.x;
    const int OPC = H / FC_CS;                 // outputs per CTA
    const int KS = FC_THREADS / OPC;           // K splits
    const int o = tid % OPC, kq = tid / OPC;
    const int kspan = K / KS;
    float acc[TB];
#pragma unroll
    for (int t = 0; t < TB; ++t) acc[t] = 0.f;
    const float4 *w4 = reinterpret_cast<const float4 *>(wp) + (size_t)(kq * kspan / 4) * H + rank * OPC + o;
    const float *xin = in + kq * kspan;
    // weights come from L2 (~600+ cycles): four 16-byte loads in flight per thread ahead of the FMAs that use them
    // (kspan is a multiple of 16 for every supported width)
    float4 wq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wq[i] = __ldg(w4 + (size_t)i * H);
#pragma unroll 1
    for (int k = 0; k < kspan; k += 16) {
        float4 wc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wc[i] = wq[i];
        w4 += (size_t)4 * H;
        if (k + 16 < kspan) {
#pragma unroll
            for (int i = 0; i < 4; ++i) wq[i] = __ldg(w4 + (size_t)i * H);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 w = wc[i];
#pragma unroll
            for (int t = 0; t < TB; ++t) {
                const float4 v = *reinterpret_cast<const float4 *>(xin + t * K + k + 4 * i);
                acc[t] = fmaf(v.x, w.x, acc[t]);
                acc[t] = fmaf(v.y, w.y, acc[t]);
                acc[t] = fmaf(v.z, w.z, acc[t]);
                acc[t] = fmaf(v.w, w.w, acc[t]);
            }
        }
    }
    // scratch may alias `in` of the FIRST layer (the staged inputs): everyone is done reading before anyone writes
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TB; ++t) scratch[(kq * TB + t) * OPC + o] = acc[t];
    __syncthreads();
    for (int e = tid; e < TB * OPC; e += FC_THREADS) {
        const int t = e / OPC, oo = e - t * OPC;
        float v = __ldg(bias + rank * OPC + oo);
        for (int q = 0; q < KS; ++q) v += scratch[(q * TB + t) * OPC + oo];
        v = v > 0.f ? v : (v != v ? v : 0.f);                  // ReLU that propagates NaN like ATen's
#pragma unroll
        for (int r = 0; r < FC_CS; ++r) cluster.map_shared_rank(dst, r)[t * H + rank * OPC + oo] = v;
    }
    cluster.sync();                                            // every CTA now holds the whole layer for the tile
}

template <int TB>
__global__ void __cluster_dims__(FC_CS, 1, 1) __launch_bounds__(FC_THREADS) fc_head_kernel(const FcHeadParams p)
{
    extern __shared__ __align__(16) float fsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tile = blockIdx.x / FC_CS;
    const int t0 = tile * TB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *xs = fsm;                            // [TB][K0]; later the K-split scratch (its first TB * 256 floats)
    float *h1 = xs + TB * p.K0;                 // [TB][H1]
    // the second hidden layer lives in the staged-input region behind the scratch when it fits there (the inputs are dead in
    // every CTA of the cluster once the first layer's closing cluster barrier has passed): 98 instead of 115 KB for the
    // 1024 -> 512 -> 256 heads, i.e. two CTAs per SM
    float *h2 = p.alias_h2 ? xs + TB * FC_THREADS : h1 + TB * p.H1;        // [TB][H2]
    float *lg = p.alias_h2 ? h1 + TB * p.H1 : h2 + TB * p.H2;              // [TB/CS .. ][16] logits of this CTA's clouds
    // all CTAs of the cluster must have started before anyone writes into a peer's shared memory
    cluster.sync();
    const int k4n = p.K0 / 4;
    for (int i = tid; i < TB * k4n; i += FC_THREADS) {
        const int t = i / k4n, c = i - t * k4n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t0 + t < p.B) v = __ldg(reinterpret_cast<const float4 *>(p.x + (int64_t)(t0 + t) * p.xs) + c);
        reinterpret_cast<float4 *>(xs)[i] = v;
    }
    __syncthreads();
    hidden_layer<TB>(cluster, xs, p.K0, p.H1, p.w1p, p.b1, xs, h1, rank);
    hidden_layer<TB>(cluster, h1, p.H1, p.H2, p.w2p, p.b2, xs, h2, rank);
    // last layer + epilogue: this CTA takes clouds t = rank, rank + 8, ...; one warp per (cloud, output)
    constexpr int TL = (TB + FC_CS - 1) / FC_CS;
    for (int e = warp; e < TL * p.OUT; e += FC_THREADS / 32) {
        const int tl = e / p.OUT, oo = e - tl * p.OUT;
        const int t = rank + tl * FC_CS;
        float s = 0.f;
        if (t < TB) {
            const float *w = p.w3 + (size_t)oo * p.H2;
            for (int k = lane; k < p.H2; k += 32) s = fmaf(h2[t * p.H2 + k], __ldg(w + k), s);
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) lg[tl * 16 + oo] = s + __ldg(p.b3 + oo);
    }
    __syncthreads();
    if (tid < TL) {
        const int t = rank + tid * FC_CS;
        const int64_t b = (int64_t)t0 + t;
        if (t < TB && b < p.B) {
            float *l = lg + tid * 16;
            float *o = p.out + b * p.OUT;
            if (p.kind == PN2_HEAD_CLS) {                      // log_softmax(-1) and x.data.max(1)[1]
                float m = l[0];
                int am = 0;
                for (int j = 1; j < p.OUT; ++j) if (l[j] > m || (l[j] != l[j] && m == m)) { m = l[j]; am = j; }
                float se = 0.f;
                for (int j = 0; j < p.OUT; ++j) se += expf(l[j] - m);
                const float lse = logf(se);
                for (int j = 0; j < p.OUT; ++j) o[j] = (l[j] - m) - lse;
                if (p.aux) reinterpret_cast<int64_t *>(p.aux)[b] = am;
            } else if (p.kind == PN2_HEAD_SIGN) {              // sigmoid(x), sign(sigmoid(x) - 0.5)
                for (int j = 0; j < p.OUT; ++j) {
                    const float sg = 1.f / (1.f + expf(-l[j]));
                    o[j] = sg;
                    const float d = sg - 0.5f;
                    if (p.aux) reinterpret_cast<float *>(p.aux)[b * p.OUT + j] = d > 0.f ? 1.f : (d < 0.f ? -1.f : d);
                }
            } else if (p.kind == PN2_HEAD_TRANS) {             // x + mean, mean optionally through its own 3 -> 6 -> 3 MLP
                float mv[3];
                for (int j = 0; j < 3; ++j) mv[j] = __ldg(p.mean + b * p.ms + j);
                if (p.mean_mlp) {
                    const float *q = p.mean_mlp;               // W1' [6][3], b1' [6], W2 [3][6], b2 [3]
                    float y[6];
                    for (int i = 0; i < 6; ++i) {
                        float a = q[18 + i];
                        for (int j = 0; j < 3; ++j) a = fmaf(q[3 * i + j], mv[j], a);
                        y[i] = a > 0.f ? a : (a != a ? a : 0.f);
                    }
                    for (int j = 0; j < 3; ++j) {
                        float a = q[42 + j];
                        for (int i = 0; i < 6; ++i) a = fmaf(q[24 + 6 * j + i], y[i], a);
                        mv[j] = a;
                    }
                }
                for (int j = 0; j < p.OUT; ++j) o[j] = l[j] + mv[j];
            } else {
                for (int j = 0; j < p.OUT; ++j) o[j] = l[j];
            }
        }
    }
}

template <int TB>
static int launch_fc_tb(FcHeadParams p, cudaStream_t st)
{
    static const bool alias_off = getenv("PN2_FC_ALIAS") && atoi(getenv("PN2_FC_ALIAS")) == 0;
    p.alias_h2 = (!alias_off && p.K0 >= FC_THREADS + p.H2) ? 1 : 0;
    const size_t smem = sizeof(float) * ((size_t)TB * (p.K0 + p.H1 + (p.alias_h2 ? 0 : p.H2)) + 16 * ((TB + FC_CS - 1) / FC_CS));
    auto kern = fc_head_kernel<TB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("fc_head: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
    }
    const unsigned tiles = (unsigned)((p.B + TB - 1) / TB);
    kern<<<tiles * FC_CS, FC_THREADS, smem, st>>>(p);
    return launch_status("pn2_fc_head_f32");
}

int launch_fc_head(const FcHeadParams &p, cudaStream_t st)
{
    if (p.B == 0) return 0;
    // few clouds: small tiles so that more SMs share the weight stream; many clouds: fewer passes over the weights
    if (p.B <= 64) return launch_fc_tb<8>(p, st);
    return launch_fc_tb<16>(p, st);
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_fc_head_f32(const float *x, int64_t x_stride, int B, int K0, int H1, int H2, int OUT,
                               const float *w1p, const float *b1, const float *w2p, const float *b2,
                               const float *w3, const float *b3, int kind,
                               const float *mean, int64_t mean_stride, const float *mean_mlp,
                               float *out, void *aux, void *stream)
{
    PN2_REQUIRE(x && w1p && b1 && w2p && b2 && w3 && b3 && out, "pn2_fc_head_f32: null pointer");
    PN2_REQUIRE(B >= 0 && OUT >= 1 && OUT <= 16, "pn2_fc_head_f32: bad sizes B=%d OUT=%d (OUT <= 16)", B, OUT);
    PN2_REQUIRE(kind >= PN2_HEAD_PLAIN && kind <= PN2_HEAD_SIGN, "pn2_fc_head_f32: bad kind %d", kind);
    PN2_REQUIRE(kind != PN2_HEAD_TRANS || (mean && OUT == 3), "pn2_fc_head_f32: the translation epilogue needs mean [B,3] and OUT = 3");
    auto layer_ok = [](int K, int H) {
        if (H % FC_CS) return false;
        const int opc = H / FC_CS;
        if (opc < 1 || opc > FC_THREADS || FC_THREADS % opc) return false;
        const int ks = FC_THREADS / opc;
        return K % (16 * ks) == 0;
    };
    PN2_REQUIRE(K0 >= 4 && K0 <= 2048 && layer_ok(K0, H1) && layer_ok(H1, H2) && H1 <= 1024 && H2 <= 1024,
                "pn2_fc_head_f32: unsupported widths %d -> %d -> %d (hidden widths must be 8 * a divisor of 256, K a multiple of 16 * the K split)", K0, H1, H2);
    PN2_REQUIRE(K0 >= FC_THREADS, "pn2_fc_head_f32: K0 = %d < %d (the K-split scratch reuses the staged input tile)", K0, FC_THREADS);
    PN2_REQUIRE(x_stride % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w1p & 15) == 0 && ((uintptr_t)w2p & 15) == 0,
                "pn2_fc_head_f32: x / packed weights must be 16-byte aligned with a row stride that is a multiple of 4");
    FcHeadParams p;
    p.x = x; p.xs = x_stride; p.B = B; p.K0 = K0; p.H1 = H1; p.H2 = H2; p.OUT = OUT; p.kind = kind;
    p.w1p = w1p; p.b1 = b1; p.w2p = w2p; p.b2 = b2; p.w3 = w3; p.b3 = b3;
    p.mean = mean; p.ms = mean_stride; p.mean_mlp = mean_mlp; p.out = out; p.aux = aux;
    return launch_fc_head(p, (cudaStream_t)stream);
}
