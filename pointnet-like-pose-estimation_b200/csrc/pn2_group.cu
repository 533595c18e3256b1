// Row gathers: index_points (pointnet2_utils.py:28-45) and the materialised grouped tensor of
// sample_and_group / sample_and_group_all / the MSG branch (:109-116, :134-140, :205-211).
// Pure data movement (HBM/L2 bound): one warp per output row, lanes over channels, so that a row whose
// source is channel-contiguous is read and written with full sectors.  The fused set-abstraction kernels
// never call these; they exist for the stand-alone functions and for training mode.
#include "pn2_common.cuh"

namespace pn2 {

__global__ void index_points_kernel(const float *__restrict__ src, int64_t sb, int64_t sn, int64_t sc, int N, int C,
                                    const int64_t *__restrict__ idx, int64_t M, int64_t rows,
                                    float *__restrict__ out, int32_t *__restrict__ err_flag)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t b = r / M;
        int64_t n = idx[r];
        if (n < 0) n += N;                       // torch advanced indexing wraps negative indices
        if (n < 0 || n >= N) {
            if (lane == 0 && err_flag) atomicOr(err_flag, 1);
            continue;
        }
        const float *p = src + b * sb + n * sn;
        for (int k = lane; k < C; k += 32) out[r * C + k] = __ldg(p + (int64_t)k * sc);
    }
}

int launch_index_points(const float *src, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                        const int64_t *idx, int64_t M, float *out, int32_t *err_flag, cudaStream_t st)
{
    const int64_t rows = (int64_t)B * M;
    if (rows == 0 || C == 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    index_points_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, sb, sn, sc, N, C, idx, M, rows, out, err_flag);
    return launch_status("pn2_index_points_f32");
}

__global__ void group_kernel(const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                             const float *__restrict__ feat, int64_t fb, int64_t fn, int64_t fc,
                             const int32_t *__restrict__ idx32, const float *__restrict__ centres,
                             int N, int S, int K, int C, int D, int order, int64_t rows, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int W = C + D;
    const int xo = (order == PN2_ORDER_MSG) ? D : 0;
    const int fo = (order == PN2_ORDER_MSG) ? 0 : C;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int64_t g = r / K;                 // (b, s)
        const int64_t b = g / S;
        const int n = (order == PN2_ORDER_ALL) ? (int)(r % K) : idx32[r];
        float *o = out + r * W;
        if (n < 0 || n >= N) {                   // 'no neighbour' marker of an empty ball (:85): a zero row, never an OOB read
            for (int k = lane; k < W; k += 32) o[k] = 0.f;
            continue;
        }
        const float *p = points + b * sb + (int64_t)n * sn;
        for (int k = lane; k < C; k += 32) {
            float x = __ldg(p + (int64_t)k * sc);
            if (order != PN2_ORDER_ALL) x = __fsub_rn(x, __ldg(centres + g * C + k));
            o[xo + k] = x;
        }
        if (D) {
            const float *f = feat + b * fb + (int64_t)n * fn;
            for (int k = lane; k < D; k += 32) o[fo + k] = __ldg(f + (int64_t)k * fc);
        }
    }
}

// Coordinates only (D = 0) and at most 16 of them -- the grouped rows of training mode: one warp per output row would
// use 3 to 10 of its lanes and write 12 to 40 bytes per instruction.  Here a warp takes 32 consecutive rows, one per lane
// (gather + centre subtraction into shared memory), and writes the 32 * C floats as contiguous 128-byte lines.
__global__ void __launch_bounds__(256) group_rows_small_kernel(const float *__restrict__ points, int64_t sb, int64_t sn, int64_t sc,
                                                               const int32_t *__restrict__ idx32, const float *__restrict__ centres,
                                                               int N, int S, int K, int C, int order, int64_t rows, float *__restrict__ out)
{
    __shared__ float s_rows[8][32 * 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float *sr = s_rows[warp];
    for (int64_t base = warp0 * 32; base < rows; base += nwarps * 32) {
        const int64_t r = base + lane;
        if (r < rows) {
            const int64_t g = r / K;
            const int64_t b = g / S;
            const int n = (order == PN2_ORDER_ALL) ? (int)(r % K) : idx32[r];
            const bool ok = n >= 0 && n < N;      // n == N: empty ball (:85) -> zero row
            const float *p = points + b * sb + (int64_t)(ok ? n : 0) * sn;
            for (int k = 0; k < C; ++k) {
                float x = __ldg(p + (int64_t)k * sc);
                if (order != PN2_ORDER_ALL) x = __fsub_rn(x, __ldg(centres + g * C + k));
                sr[lane * C + k] = ok ? x : 0.f;
            }
        }
        __syncwarp();
        const int cnt = (int)((rows - base < 32 ? rows - base : 32) * C);
        for (int i = lane; i < cnt; i += 32) out[base * C + i] = sr[i];
        __syncwarp();
    }
}

int launch_group(const float *points, int64_t sb, int64_t sn, int64_t sc,
                 const float *feat, int64_t fb, int64_t fn, int64_t fc,
                 const int32_t *idx32, const float *centres,
                 int B, int N, int S, int K, int C, int D, int order, float *out, cudaStream_t st)
{
    const int64_t rows = (int64_t)B * S * K;
    if (rows == 0) return 0;
    if (D == 0 && C <= 16) {
        int64_t blocks = (rows + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        group_rows_small_kernel<<<(unsigned)blocks, 256, 0, st>>>(points, sb, sn, sc, idx32, centres, N, S, K, C, order, rows, out);
        return launch_status("pn2_group_f32");
    }
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    group_kernel<<<(unsigned)blocks, 256, 0, st>>>(points, sb, sn, sc, feat, fb, fn, fc, idx32, centres,
                                                   N, S, K, C, D, order, rows, out);
    return launch_status("pn2_group_f32");
}

// ---------------------------------------------------------------------------------------------------
// provider.normalization (provider.py:5-21) + provider.splice_torch (:166-180) + transpose (train_rotation.py:116):
// one CTA per cloud, three passes over the (L1/L2-resident) cloud: centroid, largest norm, write-out.
__device__ __forceinline__ float block_sum(float v, float *scratch)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    return t;
}
__global__ void __launch_bounds__(256) normalize_splice_kernel(const float *__restrict__ xyz, int64_t sb, int64_t sn, int64_t sc, int N,
                                                               const int64_t *__restrict__ category, int num_category, float *__restrict__ out)
{
    __shared__ float scratch[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *P = xyz + (int64_t)b * sb;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int n = tid; n < N; n += blockDim.x) {
        sx += __ldg(P + (int64_t)n * sn); sy += __ldg(P + (int64_t)n * sn + sc); sz += __ldg(P + (int64_t)n * sn + 2 * sc);
    }
    const float cx = __fdiv_rn(block_sum(sx, scratch), (float)N), cy = __fdiv_rn(block_sum(sy, scratch), (float)N),
                cz = __fdiv_rn(block_sum(sz, scratch), (float)N);
    float m2 = 0.f;
    for (int n = tid; n < N; n += blockDim.x) {
        const float x = __fsub_rn(__ldg(P + (int64_t)n * sn), cx), y = __fsub_rn(__ldg(P + (int64_t)n * sn + sc), cy),
                    z = __fsub_rn(__ldg(P + (int64_t)n * sn + 2 * sc), cz);
        m2 = fmaxf(m2, __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    }
    for (int o = 16; o; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = m2;
    __syncthreads();
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m2 = fmaxf(m2, scratch[w]);
    const float m = __fsqrt_rn(m2);                      // max of sqrt = sqrt of max (monotonic, correctly rounded)
    const int Ct = 3 + num_category;
    float *O = out + (int64_t)b * Ct * N;
    const int cat = num_category ? (int)category[b] : -1;
    for (int n = tid; n < N; n += blockDim.x) {
        O[n] = __fdiv_rn(__fsub_rn(__ldg(P + (int64_t)n * sn), cx), m);
        O[N + n] = __fdiv_rn(__fsub_rn(__ldg(P + (int64_t)n * sn + sc), cy), m);
        O[2 * N + n] = __fdiv_rn(__fsub_rn(__ldg(P + (int64_t)n * sn + 2 * sc), cz), m);
        for (int k = 0; k < num_category; ++k) O[(int64_t)(3 + k) * N + n] = (k == cat) ? 1.f : 0.f;
    }
}
int launch_normalize_splice(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                            const int64_t *category, int num_category, float *out, cudaStream_t st)
{
    if (B == 0) return 0;
    normalize_splice_kernel<<<B, 256, 0, st>>>(xyz, sb, sn, sc, N, category, num_category, out);
    return launch_status("pn2_normalize_splice_f32");
}

// ---------------------------------------------------------------------------------------------------
// Differentiable feature gather of training mode: index_points(feature, idx) of pointnet2_utils.py:112 / :208 and its
// backward (a scatter-add into the feature gradient).  One warp per gathered row, 16-byte accesses when the rows allow it;
// the 'no neighbour' index (n == N, empty ball) gathers zeros and scatters nothing.
template <bool VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ feat, int64_t fb, int64_t fn, const int32_t *__restrict__ idx32,
                                                          int N, int64_t M, int D, int64_t rows, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int n = idx32[r];
        const bool ok = n >= 0 && n < N;
        const float *f = feat + (r / M) * fb + (int64_t)(ok ? n : 0) * fn;
        float *o = out + r * D;
        if (VEC) {
            for (int k = lane * 4; k < D; k += 128) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) v = __ldg(reinterpret_cast<const float4 *>(f + k));
                *reinterpret_cast<float4 *>(o + k) = v;
            }
        } else {
            for (int k = lane; k < D; k += 32) o[k] = ok ? __ldg(f + k) : 0.f;
        }
    }
}
template <bool VEC>
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float *__restrict__ grad, const int32_t *__restrict__ idx32, int N, int64_t M,
                                                               int D, int64_t rows, float *__restrict__ gfeat)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const int n = idx32[r];
        if (n < 0 || n >= N) continue;
        const float *g = grad + r * D;
        float *o = gfeat + ((r / M) * N + n) * (int64_t)D;
        if (VEC) {
            for (int k = lane * 4; k < D; k += 128) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(g + k));
                atomicAdd(reinterpret_cast<float4 *>(o + k), v);        // one 16-byte reduction (sm_90+)
            }
        } else {
            for (int k = lane; k < D; k += 32) atomicAdd(o + k, __ldg(g + k));
        }
    }
}
int launch_gather_rows(const float *feat, int64_t fb, int64_t fn, const int32_t *idx32, int B, int N, int64_t M, int D, float *out, cudaStream_t st)
{
    const int64_t rows = (int64_t)B * M;
    if (rows == 0 || D == 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const bool vec = D % 4 == 0 && fb % 4 == 0 && fn % 4 == 0 && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)out & 15) == 0;
    if (vec) gather_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(feat, fb, fn, idx32, N, M, D, rows, out);
    else gather_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(feat, fb, fn, idx32, N, M, D, rows, out);
    return launch_status("pn2_gather_rows_f32");
}
int launch_scatter_add_rows(const float *grad, const int32_t *idx32, int B, int N, int64_t M, int D, float *gfeat, cudaStream_t st)
{
    const int64_t rows = (int64_t)B * M;
    if (rows == 0 || D == 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const bool vec = D % 4 == 0 && ((uintptr_t)grad & 15) == 0 && ((uintptr_t)gfeat & 15) == 0;
    if (vec) scatter_add_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(grad, idx32, N, M, D, rows, gfeat);
    else scatter_add_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(grad, idx32, N, M, D, rows, gfeat);
    return launch_status("pn2_scatter_add_rows_f32");
}

}  // namespace pn2
