// Farthest point sampling with the arithmetic of the reference's NUMPY copies of the algorithm (provider.py:183-204,
// data_utils/ModelDataLoader.py:10-31, point_collect/transformer.py:120-141): the first
// three columns of an [N, C] array in the array's own dtype -- float64 for everything np.loadtxt / the shape generators
// of data_build/ produce, float32 where the loader casts first (ModelNetDataLoader.py:102-105) -- and any N (the
// DBSCAN clusters of point_collect/collect.py:64-68 and the 10 000-point ModelNet files are larger than K1's clouds).
//
// Exactness contract: dist_n = fl(fl(fl(dx^2) + fl(dy^2)) + fl(dz^2)) with dx = fl(x_n - x_far), every operation
// rounded on its own in the array's dtype (numpy evaluates `xyz - centroid`, `** 2` and `np.sum(., -1)` as three
// array passes; the sum of three contiguous terms is sequential); distance[n] is replaced only when dist_n <
// distance[n] (NaN never replaces); the next centroid is np.argmax(distance): lowest index among equal maxima.
//
// One CTA per cloud.  Coordinates and the running distances stay in shared memory when 4 * N elements fit (N <= 6400
// float64 / 12 800 float32); larger clouds read the coordinates through L1/L2 and keep the distances in caller-provided
// scratch.  Every thread owns the same points in every iteration, so the loop needs one barrier per iteration (the
// per-warp argmax slots are double-buffered by iteration parity).
#include "pn2_common.cuh"

namespace pn2 {

namespace {

__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

constexpr int NP_SMEM_BUDGET = 200 * 1024;

template <typename T>
__device__ __forceinline__ void better(T &v, int &i, T ov, int oi)
{
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

template <typename T, bool ONCHIP>
__global__ void __launch_bounds__(1024) fps_np_kernel(const T *__restrict__ points, int64_t sb, int64_t sn, int64_t sc, int N,
                                                      const int64_t *__restrict__ start, int S, int64_t *__restrict__ idx64,
                                                      T *__restrict__ scratch)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ T s_val[2][32];
    __shared__ int s_idx[2][32];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int b = blockIdx.x;
    const T *P = points + (int64_t)b * sb;
    T *sx = reinterpret_cast<T *>(s_raw), *sy = sx + N, *sz = sy + N;
    T *dist = ONCHIP ? sz + N : scratch + (int64_t)b * N;

    for (int n = tid; n < N; n += nthr) {
        if (ONCHIP) {
            sx[n] = P[(int64_t)n * sn];
            sy[n] = P[(int64_t)n * sn + sc];
            sz[n] = P[(int64_t)n * sn + 2 * sc];
        }
        dist[n] = (T)1e10;
    }
    int far = (int)min(max(start[b], (int64_t)0), (int64_t)N - 1);
    if (ONCHIP) __syncthreads();

    for (int i = 0; i < S; ++i) {
        if (tid == 0) idx64[(int64_t)b * S + i] = far;
        T cx, cy, cz;
        if (ONCHIP) { cx = sx[far]; cy = sy[far]; cz = sz[far]; }
        else { cx = __ldg(P + (int64_t)far * sn); cy = __ldg(P + (int64_t)far * sn + sc); cz = __ldg(P + (int64_t)far * sn + 2 * sc); }
        T bv = (T)-1;
        int bi = 0x7fffffff;
        for (int n = tid; n < N; n += nthr) {
            T x, y, z;
            if (ONCHIP) { x = sx[n]; y = sy[n]; z = sz[n]; }
            else { x = __ldg(P + (int64_t)n * sn); y = __ldg(P + (int64_t)n * sn + sc); z = __ldg(P + (int64_t)n * sn + 2 * sc); }
            const T dx = sub_rn(x, cx), dy = sub_rn(y, cy), dz = sub_rn(z, cz);
            const T d = add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz));
            T dn = dist[n];
            if (d < dn) { dn = d; dist[n] = d; }
            if (dn > bv) { bv = dn; bi = n; }          // ascending n: the first of equal values stays
        }
#pragma unroll
        for (int off = 16; off; off >>= 1)
            better(bv, bi, __shfl_xor_sync(0xffffffffu, bv, off), __shfl_xor_sync(0xffffffffu, bi, off));
        const int par = i & 1;
        if (lane == 0) { s_val[par][warp] = bv; s_idx[par][warp] = bi; }
        __syncthreads();
        bv = lane < nwarps ? s_val[par][lane] : (T)-1;
        bi = lane < nwarps ? s_idx[par][lane] : 0x7fffffff;
#pragma unroll
        for (int off = 16; off; off >>= 1)
            better(bv, bi, __shfl_xor_sync(0xffffffffu, bv, off), __shfl_xor_sync(0xffffffffu, bi, off));
        far = bi;
    }
}

template <typename T>
int launch_fps_np(const T *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, const int64_t *start, int S,
                  int64_t *idx64, void *scratch, cudaStream_t stream)
{
    if (B == 0 || S == 0) return 0;
    const size_t need = 4 * (size_t)N * sizeof(T);
    const bool onchip = need <= (size_t)NP_SMEM_BUDGET;
    int threads = (N + 3) / 4;                              // about four points per thread, whole warps, 64..1024
    threads = ((threads + 31) / 32) * 32;
    threads = threads < 64 ? 64 : (threads > 1024 ? 1024 : threads);
    if (onchip) {
        static bool attr_set = false;                       // idempotent; a race only repeats the call
        if (!attr_set) {
            cudaFuncSetAttribute(fps_np_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NP_SMEM_BUDGET);
            attr_set = true;
        }
        fps_np_kernel<T, true><<<B, threads, need, stream>>>(points, sb, sn, sc, N, start, S, idx64, nullptr);
    } else {
        PN2_REQUIRE(scratch, "pn2_fps_xyz: N=%d needs %lld bytes of scratch (pn2_fps_xyz_scratch_bytes)", N,
                    (long long)B * N * (long long)sizeof(T));
        fps_np_kernel<T, false><<<B, threads, 0, stream>>>(points, sb, sn, sc, N, start, S, idx64, reinterpret_cast<T *>(scratch));
    }
    return launch_status("fps_np_kernel");
}

}  // namespace

int64_t fps_np_scratch_bytes(int B, int N, int elem_bytes)
{
    return 4 * (int64_t)N * elem_bytes <= NP_SMEM_BUDGET ? 0 : (int64_t)B * N * elem_bytes;
}

int launch_fps_np_f64(const double *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, const int64_t *start, int S,
                      int64_t *idx64, void *scratch, cudaStream_t stream)
{
    return launch_fps_np<double>(points, sb, sn, sc, B, N, start, S, idx64, scratch, stream);
}

int launch_fps_np_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, const int64_t *start, int S,
                      int64_t *idx64, void *scratch, cudaStream_t stream)
{
    return launch_fps_np<float>(points, sb, sn, sc, B, N, start, S, idx64, scratch, stream);
}

}  // namespace pn2
