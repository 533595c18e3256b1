// K3 (fp32 mode) -- fused set abstraction on the FP32 pipes: gather -> shared MLP with folded BatchNorm ->
// ReLU -> max over the K neighbours (reference: model/pointnet2_utils.py:166-172 and :204-218).
//
// A CTA owns a tile of R grouped rows (R/K whole groups, or an R-row slice of one large group).  The rows
// are gathered straight into shared memory (channel-major, so row-parallel threads are bank-conflict free),
// every layer is a [R x Cin] x [Cin x Cout] register-tiled FFMA GEMM whose weights stream through a
// cp.async double buffer, activations ping-pong between two shared-memory buffers, and only the pooled
// [groups x Cout] result is written to HBM: the [B,S,K,Cin] tensor and every intermediate activation of
// the reference stay on chip.  This is the <= 1e-5 path; the tensor-core path is pn2_sa_tc.cu.
#include "pn2_common.cuh"

namespace pn2 {

constexpr int SA_THREADS = 256;
constexpr int SA_KC = 16;        // k rows of W per staged chunk
constexpr int SA_CW = 8;         // output channels per thread (x 4 rows)
constexpr int F32_CIN_ALIGN = 16;
constexpr int F32_COUT_ALIGN = 256;   // a pass covers up to 256 output channels (R = 32)

static inline int rup(int x, int a) { return (x + a - 1) / a * a; }

// packed layout (floats): for each layer  Wt[CinPad][CoutPad] (BN scale folded, zero padded) then bias[CoutPad]
static int64_t ffma_image_bytes(int L, const int *dims)
{
    int64_t n = 0;
    for (int l = 0; l < L; ++l) {
        const int64_t cip = rup(dims[l], F32_CIN_ALIGN), cop = rup(dims[l + 1], F32_COUT_ALIGN);
        n += cip * cop + cop;
    }
    return (n * (int64_t)sizeof(float) + 255) / 256 * 256;
}
// fp32 mode carries two images of the folded weights: the FFMA kernel's (this file) and, behind it, the split (bf16 hi, bf16 lo)
// image of the tensor-core kernel (pn2_sa_tcs.cu); the launcher picks the kernel per configuration
int64_t packed_bytes_f32(int L, const int *dims)
{
    return ffma_image_bytes(L, dims) + packed_bytes_f32_split(L, dims);
}

struct PackArgs {
    int L;
    int dims[PN2_MAX_LAYERS + 1];
    const float *w[PN2_MAX_LAYERS], *b[PN2_MAX_LAYERS], *g[PN2_MAX_LAYERS], *be[PN2_MAX_LAYERS],
        *mu[PN2_MAX_LAYERS], *var[PN2_MAX_LAYERS];
    float eps;
};

__global__ void pack_f32_kernel(PackArgs a, float *__restrict__ packed)
{
    int64_t off = 0;
    for (int l = 0; l < a.L; ++l) {
        const int ci = a.dims[l], co = a.dims[l + 1];
        const int cip = (ci + F32_CIN_ALIGN - 1) / F32_CIN_ALIGN * F32_CIN_ALIGN;
        const int cop = (co + F32_COUT_ALIGN - 1) / F32_COUT_ALIGN * F32_COUT_ALIGN;
        const int64_t nW = (int64_t)cip * cop;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nW + cop; i += (int64_t)gridDim.x * blockDim.x) {
            const bool is_bias = i >= nW;
            const int c = is_bias ? (int)(i - nW) : (int)(i % cop);
            const int k = is_bias ? 0 : (int)(i / cop);
            float v = 0.f;
            if (c < co && k < ci) {
                const float s = __fdiv_rn(a.g[l][c], __fsqrt_rn(__fadd_rn(a.var[l][c], a.eps)));
                if (is_bias) v = __fadd_rn(__fmul_rn(__fsub_rn(a.b[l][c], a.mu[l][c]), s), a.be[l][c]);
                else v = __fmul_rn(a.w[l][(int64_t)c * ci + k], s);
            }
            packed[off + i] = v;
        }
        off += nW + cop;
    }
}

int launch_pack_f32(int L, const int *dims, const void *const *ptrs, float eps, int C, int order, void *packed, cudaStream_t st)
{
    {
        const int rc = launch_pack_bf16(L, dims, ptrs, eps, C, order, (unsigned char *)packed + ffma_image_bytes(L, dims), st, true);
        if (rc) return rc;
    }
    PackArgs a;
    a.L = L; a.eps = eps;
    for (int l = 0; l <= L; ++l) a.dims[l] = dims[l];
    for (int l = 0; l < L; ++l) {
        a.w[l] = (const float *)ptrs[6 * l + 0]; a.b[l] = (const float *)ptrs[6 * l + 1];
        a.g[l] = (const float *)ptrs[6 * l + 2]; a.be[l] = (const float *)ptrs[6 * l + 3];
        a.mu[l] = (const float *)ptrs[6 * l + 4]; a.var[l] = (const float *)ptrs[6 * l + 5];
    }
    pack_f32_kernel<<<148, 256, 0, st>>>(a, (float *)packed);
    return launch_status("pn2_pack_weights(f32)");
}

struct SaF32Params {
    const float *points; int64_t sb, sn, sc;
    const float *feat;   int64_t fb, fn, fc;
    const float *centres; const int32_t *idx32;
    int N, S, K, C, D, order;
    int L; int dims[PN2_MAX_LAYERS + 1];
    const float *packed;
    float *out; int64_t ob, os; int co_off;
    int64_t G;          // number of groups (B*S)
    int gpt;            // groups per tile (K <= R) or 0
    int cpg;            // row chunks per group (K > R) or 1
    int64_t tiles;
    int bufc[2];        // channel capacity of the two activation buffers
    int feat_vec4;      // feature rows can be read as float4
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int R>
__global__ void __launch_bounds__(SA_THREADS) sa_f32_kernel(const SaF32Params p)
{
    constexpr int TXN = R / 4;              // thread columns over rows (4 rows each)
    constexpr int TYN = SA_THREADS / TXN;   // thread rows over output channels (SA_CW each)
    constexpr int CCH = SA_CW * TYN;        // output channels per pass (128 for R=64, 256 for R=32)
    constexpr int W4 = SA_KC * CCH / 4 / SA_THREADS;  // float4 per thread per staged W chunk

    extern __shared__ __align__(16) float smem[];
    // the two activation buffers as offsets from `smem` (an indexed array of pointers would demote every access to a
    // generic load / store)
    const int buf1_off = p.bufc[0] * R;
    float *Ws = smem + buf1_off + (size_t)p.bufc[1] * R;        // [2][SA_KC][CCH]
    int *s_n = (int *)(Ws + 2 * SA_KC * CCH);          // [R] source point of each row, -1 = padding row
    int64_t *s_g = (int64_t *)(s_n + R);               // [R] group of each row

    const int tid = threadIdx.x, lane = tid & 31;
    const int tx = tid % TXN, ty = tid / TXN;
    const int K = p.K, C = p.C, D = p.D;
    const int Cin0 = C + D;
    const int xo = (p.order == PN2_ORDER_MSG) ? D : 0;
    const int fo = (p.order == PN2_ORDER_MSG) ? 0 : C;

    for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        // ---- row bookkeeping ---------------------------------------------------------------------------
        __syncthreads();
        for (int r = tid; r < R; r += SA_THREADS) {
            int64_t g; int j;
            if (p.gpt) { g = tile * p.gpt + r / K; j = r % K; if (r >= p.gpt * K) g = p.G; }
            else       { g = tile / p.cpg; j = (int)(tile % p.cpg) * R + r; }
            int n = -1;
            if (g < p.G && j < K) n = (p.order == PN2_ORDER_ALL) ? j : p.idx32[g * K + j];
            if (n >= p.N) n = -1;                // 'no neighbour' marker of an empty ball (pointnet2_utils.py:85)
            s_n[r] = n; s_g[r] = g;
        }
        __syncthreads();
        // ---- gather layer-0 activations X0[c][r] ----------------------------------------------------------
        float *X0 = smem;
        const int cip0 = (Cin0 + F32_CIN_ALIGN - 1) / F32_CIN_ALIGN * F32_CIN_ALIGN;
        for (int i = tid; i < R * (cip0 - Cin0); i += SA_THREADS) X0[(Cin0 + i / R) * R + i % R] = 0.f;
        for (int i = tid; i < R * C; i += SA_THREADS) {
            const int r = i % R, k = i / R;
            const int n = s_n[r];
            float v = 0.f;
            if (n >= 0) {
                const int64_t g = s_g[r];
                const int64_t b = g / p.S;
                v = __ldg(p.points + b * p.sb + (int64_t)n * p.sn + (int64_t)k * p.sc);
                if (p.order != PN2_ORDER_ALL) v = __fsub_rn(v, __ldg(p.centres + g * C + k));
            }
            X0[(xo + k) * R + r] = v;
        }
        if (D) {
            // lane <-> row, loop over channels: conflict-free transposed store, sector reuse through L1
            const int warp = tid >> 5;
            for (int r0 = warp * 32; r0 < R; r0 += (SA_THREADS / 32) * 32) {
                const int r = r0 + lane;
                const int n = s_n[r];
                const float *f = nullptr;
                if (n >= 0) f = p.feat + (s_g[r] / p.S) * p.fb + (int64_t)n * p.fn;
                if (p.feat_vec4) {
                    for (int k = 0; k < D; k += 4) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (f) v = __ldg(reinterpret_cast<const float4 *>(f + k));
                        X0[(fo + k + 0) * R + r] = v.x; X0[(fo + k + 1) * R + r] = v.y;
                        X0[(fo + k + 2) * R + r] = v.z; X0[(fo + k + 3) * R + r] = v.w;
                    }
                } else {
                    for (int k = 0; k < D; ++k) X0[(fo + k) * R + r] = f ? __ldg(f + (int64_t)k * p.fc) : 0.f;
                }
            }
        }
        // (visibility of X0 is covered by the first __syncthreads of the k loop below)

        // ---- layers ---------------------------------------------------------------------------------------
        const float *wl = p.packed;
        for (int l = 0; l < p.L; ++l) {
            const int cin = p.dims[l], cout = p.dims[l + 1];
            const int cip = (cin + F32_CIN_ALIGN - 1) / F32_CIN_ALIGN * F32_CIN_ALIGN;
            const int cop = (cout + F32_COUT_ALIGN - 1) / F32_COUT_ALIGN * F32_COUT_ALIGN;
            const float *Wt = wl, *bias = wl + (size_t)cip * cop;
            const float *Xin = smem + ((l & 1) ? buf1_off : 0);
            float *Xout = smem + ((l & 1) ? 0 : buf1_off);
            const bool last = (l == p.L - 1);
            const int next_cip = last ? 0 : (cout + F32_CIN_ALIGN - 1) / F32_CIN_ALIGN * F32_CIN_ALIGN;
            const int nk = cip / SA_KC;

            for (int co0 = 0; co0 < cout; co0 += CCH) {
                // accumulators as packed pairs over two adjacent output channels: Blackwell's FP32 pipe issues one scalar FFMA
                // every other cycle per scheduler, fma.rn.f32x2 does two in the same slot -- same per-element rounding
                // 4 rows x 8 channels per thread: one activation quad + two weight quads from shared memory feed 32 FMAs
                // (the 4 x 4 tile this replaces was bound by shared-memory bandwidth, 3 wavefronts per 16 FMAs).
                f32x2 acc2[4][SA_CW / 2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < SA_CW / 2; ++j) acc2[i][j] = pk2(0.f, 0.f);

                auto prefetch = [&](int kc) {
                    float *dst = Ws + (kc & 1) * SA_KC * CCH;
#pragma unroll
                    for (int q = 0; q < W4; ++q) {
                        const int e = (tid + q * SA_THREADS) * 4;          // float index inside the chunk
                        const int kk = e / CCH, cc = e % CCH;
                        cp_async16(dst + e, Wt + (size_t)(kc * SA_KC + kk) * cop + co0 + cc);
                    }
                    cp_async_commit();
                };
                const bool active = co0 + SA_CW * ty < cout;
                prefetch(0);
                for (int kc = 0; kc < nk; ++kc) {
                    cp_async_wait_all();
                    __syncthreads();
                    if (kc + 1 < nk) prefetch(kc + 1);
                    const float *wsb = Ws + (kc & 1) * SA_KC * CCH + SA_CW * ty;
                    const float *xb = Xin + (size_t)(kc * SA_KC) * R + 4 * tx;
                    if (!active) continue;             // this thread's channels are padding of a narrow layer (whole warps: ty is warp-uniform up to 2 values)
                    // operands of step kk + 1 are requested before the FMAs of step kk (register double buffer)
                    float4 a_n = *reinterpret_cast<const float4 *>(xb);
                    float4 w_n = *reinterpret_cast<const float4 *>(wsb);
                    float4 u_n = *reinterpret_cast<const float4 *>(wsb + 4);
#pragma unroll
                    for (int kk = 0; kk < SA_KC; ++kk) {
                        const float4 a = a_n, w = w_n, u = u_n;
                        if (kk + 1 < SA_KC) {
                            a_n = *reinterpret_cast<const float4 *>(xb + (kk + 1) * R);
                            w_n = *reinterpret_cast<const float4 *>(wsb + (kk + 1) * CCH);
                            u_n = *reinterpret_cast<const float4 *>(wsb + (kk + 1) * CCH + 4);
                        }
                        const f32x2 wv[4] = {pk2(w.x, w.y), pk2(w.z, w.w), pk2(u.x, u.y), pk2(u.z, u.w)};
                        const f32x2 av[4] = {pk2(a.x, a.x), pk2(a.y, a.y), pk2(a.z, a.z), pk2(a.w, a.w)};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < SA_CW / 2; ++j) acc2[i][j] = fma2(av[i], wv[j], acc2[i][j]);
                    }
                }
                float acc[4][SA_CW];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < SA_CW / 2; ++j) {
                        const float2 v = unpk2(acc2[i][j]);
                        acc[i][2 * j] = v.x; acc[i][2 * j + 1] = v.y;
                    }
                // ---- epilogue: bias + ReLU ---------------------------------------------------------------------
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + co0 + SA_CW * ty));
                const float4 b5 = __ldg(reinterpret_cast<const float4 *>(bias + co0 + SA_CW * ty + 4));
                const float bv[SA_CW] = {b4.x, b4.y, b4.z, b4.w, b5.x, b5.y, b5.z, b5.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < SA_CW; ++j) acc[i][j] = fmaxf(acc[i][j] + bv[j], 0.f);

                if (!last) {
#pragma unroll
                    for (int j = 0; j < SA_CW; ++j) {
                        const int co = co0 + SA_CW * ty + j;
                        if (co < next_cip)
                            *reinterpret_cast<float4 *>(Xout + (size_t)co * R + 4 * tx) =
                                make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                    }
                } else {
                    // ---- max over the rows of each group ----------------------------------------------------
                    const int rows_per_group = p.gpt ? K : R;   // rows of one group inside this tile
                    const bool pow2 = (rows_per_group >= 4) && ((rows_per_group & (rows_per_group - 1)) == 0);
                    if (pow2) {
                        float m[SA_CW];
#pragma unroll
                        for (int j = 0; j < SA_CW; ++j) {
                            m[j] = 0.f;                               // post-ReLU values are >= 0
#pragma unroll
                            for (int i = 0; i < 4; ++i) if (s_n[4 * tx + i] >= 0) m[j] = fmaxf(m[j], acc[i][j]);
                        }
                        const int q = rows_per_group / 4;             // thread columns per group
                        for (int o = 1; o < q; o <<= 1)
#pragma unroll
                            for (int j = 0; j < SA_CW; ++j) m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], o));
                        if (tx % q == 0) {
                            const int64_t g = s_g[4 * tx];
                            if (g < p.G) {            // an empty ball (no valid row) pools to 0
                                float *o = p.out + (g / p.S) * p.ob + (g % p.S) * p.os + p.co_off + co0 + SA_CW * ty;
#pragma unroll
                                for (int j = 0; j < SA_CW; ++j) {
                                    if (co0 + SA_CW * ty + j < cout) {
                                        if (p.gpt) o[j] = m[j];
                                        else atomicMax(reinterpret_cast<int *>(o + j), __float_as_int(m[j]));
                                    }
                                }
                            }
                        }
                    } else {
                        // generic K: stage the chunk in the free buffer, then one thread per (channel, group)
                        float *Y = Xout;
#pragma unroll
                        for (int j = 0; j < SA_CW; ++j)
                            *reinterpret_cast<float4 *>(Y + (size_t)(SA_CW * ty + j) * R + 4 * tx) =
                                make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                        __syncthreads();
                        const int ng = p.gpt ? p.gpt : 1;
                        for (int it = tid; it < CCH * ng; it += SA_THREADS) {
                            const int cl = it % CCH, gl = it / CCH;
                            const int r_begin = gl * rows_per_group;
                            const int64_t g = s_g[r_begin];
                            if (g >= p.G || co0 + cl >= cout) continue;
                            float m = 0.f; bool any = false;
                            for (int r = r_begin; r < r_begin + rows_per_group && r < R; ++r)
                                if (s_n[r] >= 0) { m = fmaxf(m, Y[(size_t)cl * R + r]); any = true; }
                            (void)any;                                // an empty ball pools to 0
                            float *o = p.out + (g / p.S) * p.ob + (g % p.S) * p.os + p.co_off + co0 + cl;
                            if (p.gpt) *o = m;
                            else atomicMax(reinterpret_cast<int *>(o), __float_as_int(m));
                        }
                    }
                }
                __syncthreads();   // Ws and (generic pool) Y are reused by the next pass
            }
            wl += (size_t)cip * cop + cop;
        }
    }
}

static void f32_plan(int S, int K, int L, const int *dims, int *R_out, int bufc[2], size_t *smem_out)
{
    (void)S; (void)K;
    for (int R = 64; R >= 32; R >>= 1) {
        const int CCH = SA_CW * (SA_THREADS / (R / 4));
        int c[2] = {0, 0};
        for (int l = 0; l < L; ++l) {
            const int cip = rup(dims[l], F32_CIN_ALIGN);
            if (cip > c[l & 1]) c[l & 1] = cip;
        }
        if (CCH > c[L & 1]) c[L & 1] = CCH;       // generic pooling stages one pass of the last layer there
        const size_t smem = sizeof(float) * ((size_t)(c[0] + c[1]) * R + 2 * SA_KC * CCH) + sizeof(int) * R + sizeof(int64_t) * R + 16;
        if (smem <= 220 * 1024 || R == 32) {
            *R_out = R; bufc[0] = c[0]; bufc[1] = c[1]; *smem_out = smem;
            return;
        }
    }
}

int sa_f32_needs_zero(int S, int K, int L, const int *dims)
{
    const int tr = sa_f32_split_tile_rows(K, L, dims);
    if (tr) return K > tr;
    int R, bufc[2]; size_t smem;
    f32_plan(S, K, L, dims, &R, bufc, &smem);
    return K > R;
}

int launch_sa_f32(const SaArgs &a, cudaStream_t st)
{
    {   // tensor-core kernel on split operands where it covers the configuration
        const int rc = launch_sa_f32_split(a, (const unsigned char *)a.packed + ffma_image_bytes(a.L, a.dims), st);
        if (rc != -100) return rc;
    }
    SaF32Params p;
    p.points = a.points; p.sb = a.sb; p.sn = a.sn; p.sc = a.sc;
    p.feat = a.feat; p.fb = a.fb; p.fn = a.fn; p.fc = a.fc;
    p.centres = a.centres; p.idx32 = a.idx32;
    p.N = a.N; p.S = a.S; p.K = a.K; p.C = a.C; p.D = a.D; p.order = a.order;
    p.L = a.L;
    for (int l = 0; l <= a.L; ++l) p.dims[l] = a.dims[l];
    p.packed = (const float *)a.packed;
    p.out = a.out; p.ob = a.ob; p.os = a.os; p.co_off = a.co_off;
    p.G = (int64_t)a.B * a.S;
    int R; size_t smem;
    f32_plan(a.S, a.K, a.L, a.dims, &R, p.bufc, &smem);
    PN2_REQUIRE(smem <= 227 * 1024, "pn2_sa_fused(f32): layer widths need %zu bytes of shared memory", smem);
    if (a.K <= R) { p.gpt = R / a.K; p.cpg = 1; p.tiles = (p.G + p.gpt - 1) / p.gpt; }
    else          { p.gpt = 0; p.cpg = (a.K + R - 1) / R; p.tiles = p.G * p.cpg; }
    p.feat_vec4 = (a.D > 0 && a.fc == 1 && a.D % 4 == 0 && a.fn % 4 == 0 && a.fb % 4 == 0 &&
                   ((uintptr_t)a.feat & 15) == 0);
    if (p.tiles == 0) return 0;
    int64_t grid = p.tiles;
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    const int64_t cap = (int64_t)148 * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
    if (grid > cap) grid = cap;
    cudaError_t e;
    if (R == 64) {
        e = cudaFuncSetAttribute(sa_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("sa_f32: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        sa_f32_kernel<64><<<(unsigned)grid, SA_THREADS, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(sa_f32_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("sa_f32: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
        sa_f32_kernel<32><<<(unsigned)grid, SA_THREADS, smem, st>>>(p);
    }
    return launch_status("pn2_sa_fused(f32)");
}

}  // namespace pn2
