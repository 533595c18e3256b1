// Probe: does ptxas contract mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (single rounding)?  The FPS / ball-query
// exactness contract needs separately rounded products and sums.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../pointnet-like-pose-estimation_b200/csrc/pn2_common.cuh"
using namespace pn2;
template <int MODE>
__global__ void k(const float *a, const float *c, float *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f32x2 t[3];
    for (int kk = 0; kk < 3; ++kk) {
        f32x2 p = pk2(a[(2 * i) * 3 + kk], a[(2 * i + 1) * 3 + kk]);
        f32x2 nc = pk2(-c[kk], -c[kk]);
        f32x2 df = add2(p, nc);
        if (MODE == 0) t[kk] = mul2(df, df);
        else if (MODE == 1) t[kk] = fma2(df, df, pk2(0.f, 0.f));
        else { float2 v = unpk2(df); t[kk] = pk2(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)); }
    }
    float2 d = unpk2(sum_order2<3, false>(t));
    out[2 * i] = d.x; out[2 * i + 1] = d.y;
}
int main()
{
    const int n = 1 << 20;
    float *a = (float *)malloc(sizeof(float) * n * 6), c[3] = {0.137f, -0.422f, 0.731f}, *o = (float *)malloc(sizeof(float) * 2 * n);
    srand(1);
    for (int i = 0; i < n * 6; ++i) a[i] = 2.f * rand() / RAND_MAX - 1.f;
    float *da, *dc, *dout;
    cudaMalloc(&da, sizeof(float) * n * 6); cudaMalloc(&dc, 12); cudaMalloc(&dout, sizeof(float) * 2 * n);
    cudaMemcpy(da, a, sizeof(float) * n * 6, cudaMemcpyHostToDevice); cudaMemcpy(dc, c, 12, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 3; ++mode) {
    if (mode == 0) k<0><<<n / 256, 256>>>(da, dc, dout, n);
    if (mode == 1) k<1><<<n / 256, 256>>>(da, dc, dout, n);
    if (mode == 2) k<2><<<n / 256, 256>>>(da, dc, dout, n);
    cudaMemcpy(o, dout, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < 2 * n; ++i) {
        volatile float t0, t1, t2, s;
        volatile float d0 = a[i * 3 + 0] - c[0], d1 = a[i * 3 + 1] - c[1], d2 = a[i * 3 + 2] - c[2];
        t0 = d0 * d0; t1 = d1 * d1; t2 = d2 * d2; s = t0 + t1; s = s + t2;
        if (s != o[i]) ++bad;
    }
    printf("mode %d (0 mul2+add2, 1 fma2(x,x,+0)+add2, 2 scalar mul + add2): %ld of %d results differ from separately rounded arithmetic (%s)\n", mode, bad, 2 * n, cudaGetErrorString(cudaGetLastError()));
  }
    return 0;
}
