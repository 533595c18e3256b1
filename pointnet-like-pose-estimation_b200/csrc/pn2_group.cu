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

int launch_group(const float *points, int64_t sb, int64_t sn, int64_t sc,
                 const float *feat, int64_t fb, int64_t fn, int64_t fc,
                 const int32_t *idx32, const float *centres,
                 int B, int N, int S, int K, int C, int D, int order, float *out, cudaStream_t st)
{
    const int64_t rows = (int64_t)B * S * K;
    if (rows == 0) return 0;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    group_kernel<<<(unsigned)blocks, 256, 0, st>>>(points, sb, sn, sc, feat, fb, fn, fc, idx32, centres,
                                                   N, S, K, C, D, order, rows, out);
    return launch_status("pn2_group_f32");
}

}  // namespace pn2
