// K3 (bf16 mode), warp-specialised schedule of the pipelined fused set-abstraction kernel (pn2_sa_tc2.cu describes the
// data path: tiles of 128 grouped rows, in-place activation buffers, bias inside the MMA, transposed last layer with an
// in-register max-pool).  Reference: model/pointnet2_utils.py:166-172 (SSG) and :204-218 (one MSG branch).
//
// One persistent CTA per SM keeps U tiles ("units") in flight and splits the work by ROLE instead of by tile:
//
//   4 producer warps   warp p owns rows 32p..32p+31 of every tile: neighbour index -> feature rows (bf16, 16-byte
//                      loads, 8 lanes per 128-byte line) and coordinates, all in flight in registers while the unit's
//                      previous tile is still being computed; once the unit's operand buffer is free
//                      (tcgen05.commit on buf_free[u]) they are stored into the 128-byte-swizzled operand, fenced to the
//                      async proxy and announced on full[u].  Nobody else issues a global load, so nobody else stalls
//                      in fence.proxy.async (a MEMBAR that waits for the thread's loads in flight); TMA gather4 was
//                      measured at ~4000 cycles per 32 KB tile (latency-bound inside the TMA unit), the load/store
//                      path lands a tile in ~1000.
//   1 MMA warp         walks the (layer, unit) schedule; when a unit's next layer has its operand (full[u] / w_done[u]) it
//                      issues the layer (one elected lane, descriptors on the uniform datapath) and commits to
//                      acc_ready[u] (+ buf_free[u] after the last layer).
//   16 worker warps    warp (q, hb) covers TMEM lane quarter q and column share hb of EVERY tile: it waits on the units'
//                      acc_ready barriers in the same order and runs its share of the epilogue / pooling phase, then
//                      arrives on w_done[u].  A phase is 4x shorter than with one warpgroup per tile and the MMAs of one
//                      unit run underneath the CUDA-core phases of the other.  (Both loops can also POLL the units and
//                      take whichever is ready, PN2_TC4_STATIC=0; see the status note.)
//
// STATUS (round 1), cls_ssg at B = 2048 on B200: with the FIXED (layer, unit) schedule (default, PN2_TC4_STATIC=1: the
// polling loops below block instead of moving on) this is the default for two-unit stacks -- sa2 2.68 ms against 2.79 ms
// for the cooperative kernel with its TMA gather.  Four-unit stacks stay on the unit kernel (sa1: 3.6 ms here against
// 1.46 ms; a 16-column share per phase is all latency).  With DYNAMIC polling (PN2_TC4_STATIC=0) CTA 0 runs an sa2 tile
// pair in 9.5-10.9 k cycles, but 10-20 % of the CTAs, different ones every launch, fall into a slow mode (up to 1.6x)
// with the two units in phase, and the launch takes 3.8 ms; see profiles/README.md.
#include <stdlib.h>
#include "pn2_sa_pipe.cuh"

namespace pn2 {

constexpr int T4_WORKERS = 16, T4_PRODUCERS = 4;
// warpgroups (setmaxnreg moves registers between them): 0-3 workers, 4 producers, 5 = the MMA warp + three idle warps
constexpr int T4_PROD_WARP0 = T4_WORKERS;                 // warps 16..19
constexpr int T4_MMA_WARP = T4_WORKERS + T4_PRODUCERS;    // warp 20
constexpr int T4_THREADS = 24 * 32;                       // 768 threads: 80 registers per thread at launch
constexpr int T4_REGS_PRODUCER = 120, T4_REGS_MMA = 40;   // producers +40 x 128 threads, MMA warpgroup -40 x 128
constexpr uint32_t T4_POLL_NS = 20;
constexpr bool T4_POOL_PAIR = false;                      // workers stay at the launch allocation of 80 registers: one tcgen05.ld in flight in the pool

__device__ __forceinline__ void sts_relu8(uint32_t addr, const uint32_t *r)
{
    sts128(addr, relu_pack_u(r[0], r[1]), relu_pack_u(r[2], r[3]), relu_pack_u(r[4], r[5]), relu_pack_u(r[6], r[7]));
}
// column share hb (of 4) of a hidden layer's epilogue for this thread's row: 16-byte chunks [hb*CPS, (hb+1)*CPS),
// CPS = Npad / 32; chunk c of the row lives at (a_rowx ^ ((c & 7) << 4)) + (c >> 3) * SW128_BLOCK
__device__ __forceinline__ void epi_share(int cps, int hb, uint32_t tmem_row, uint32_t a_rowx)
{
    const int c0 = hb * cps;
    const uint32_t t0 = tmem_row + (uint32_t)(c0 * 8);
    const uint32_t a0 = (a_rowx ^ (uint32_t)((c0 & 7) << 4)) + (uint32_t)((c0 >> 3) * SW128_BLOCK);   // valid for cps = 1, 2, 4, 8
    if (cps == 4) {
        uint32_t r[32];
        tmem_ld32_nowait(t0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) sts_relu8(a0 ^ (uint32_t)(i << 4), r + 8 * i);
    } else if (cps == 8) {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(t0, r0);
        tmem_ld32_nowait(t0 + 32u, r1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) sts_relu8(a0 ^ (uint32_t)(i << 4), r0 + 8 * i);
#pragma unroll
        for (int i = 0; i < 4; ++i) sts_relu8(a0 ^ (uint32_t)((i + 4) << 4), r1 + 8 * i);
    } else if (cps == 2) {
        uint32_t r[16];
        tmem_ld16_nowait(t0, r);
        tmem_ld_wait();
        sts_relu8(a0, r);
        sts_relu8(a0 ^ 16u, r + 8);
    } else {
#pragma unroll 1
        for (int c = c0; c < c0 + cps; ++c) {
            uint32_t r[8];
            tmem_ld8_nowait(tmem_row + (uint32_t)(c * 8), r);
            tmem_ld_wait();
            sts_relu8((a_rowx ^ (uint32_t)((c & 7) << 4)) + (uint32_t)((c >> 3) * SW128_BLOCK), r);
        }
    }
}

__device__ __forceinline__ uint4 ldg128(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <bool SMALL_C>
__global__ void __launch_bounds__(T4_THREADS, 1) sa_tc4_kernel(const __grid_constant__ Tc2Params p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const Tc2Smem sm = carve(smem, p);
    const int U = p.U;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = p.L;
    uint64_t *const full = sm.feat_ready, *const w_done = sm.a_ready, *const acc_ready = sm.acc_ready, *const buf_free = sm.buf_free;

    cta_prologue(sm, p, tid, T4_WORKERS, T4_PRODUCERS);
    if (warp == T4_MMA_WARP) tmem_alloc(sm.tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    const int tiles = (int)p.tiles;
    const int stride = (int)gridDim.x * U;
    const int first_tile = (int)blockIdx.x * U;
    // tiles of unit u: first_tile + u + it * stride, it < n_it(u)
    auto n_it_of = [&](int u) -> int { const int t0 = first_tile + u; return t0 < tiles ? (tiles - t0 + stride - 1) / stride : 0; };
    const int K = p.K, C = p.C, D = p.D, S = p.S, N = p.N;
    const int logK = 31 - __clz(K);
    const bool static_sched = p.static_sched != 0;

    if (warp >= T4_MMA_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T4_REGS_MMA));
    }
    if (warp > T4_MMA_WARP) {
        // idle warps of the MMA warpgroup
    } else if (warp == T4_MMA_WARP) {
        // ================================ MMA warp ====================================================================
        const uint32_t wsm16 = smem_u32(sm.Wsm) >> 4;
        const uint32_t ones_lo = desc_lo_ns(smem_u32(sm.ones), A_LBO);
        int ml[T2_MAX_UNITS], mit[T2_MAX_UNITS], nit[T2_MAX_UNITS];
        uint32_t wd_par = 0;                          // parity of the next w_done completion to consume, per unit
        int remaining = 0;
#pragma unroll
        for (int u = 0; u < T2_MAX_UNITS; ++u) { ml[u] = 0; mit[u] = 0; nit[u] = u < U ? n_it_of(u) : 0; remaining += nit[u] * L; }
        mbar_wait(sm.w_bar, 0);
        Tracer tr; tr.init(p.trace, 4, blockIdx.x == 0 && lane == 0);
        while (remaining > 0) {
#pragma unroll
            for (int u = 0; u < T2_MAX_UNITS; ++u) {
                if (mit[u] >= nit[u]) continue;
                const int l = ml[u];
                // layer 0 needs the gathered operand and (after the first tile) the previous tile's pool to have drained
                // TMEM; layer l > 0 needs epilogue l-1.  full[u] cannot advance before this layer is issued, so probing
                // it repeatedly without consuming it is safe.
                if (static_sched) {          // fixed (layer, unit) order: block instead of moving on to the next unit
                    if (l == 0) mbar_wait(full + u, (uint32_t)mit[u] & 1u);
                    if (l > 0 || mit[u] > 0) mbar_wait(w_done + u, (wd_par >> u) & 1u);
                } else {
                    if (l == 0 && !mbar_probe(full + u, (uint32_t)mit[u] & 1u, T4_POLL_NS)) continue;
                    if ((l > 0 || mit[u] > 0) && !mbar_probe(w_done + u, (wd_par >> u) & 1u, T4_POLL_NS)) continue;
                }
                if (l > 0 || mit[u] > 0) wd_par ^= 1u << u;
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sm.A_base) + (uint32_t)(u * p.a_unit_bytes);
                if (elect_one()) {
                    issue_layer_mmas(p, l, wsm16, desc_lo_sw128(a_addr), desc_lo_ns(a_addr + (uint32_t)p.xyz_off, A_LBO), ones_lo,
                                     tmem_base + (uint32_t)(u * p.cols_per_unit));
                    umma_commit(acc_ready + u);
                    if (l == L - 1) umma_commit(buf_free + u);
                }
                __syncwarp();
                tr.mark(0x100 + l * 16 + u);
                if (++ml[u] == L) { ml[u] = 0; ++mit[u]; }
                --remaining;
            }
        }
    } else if (warp >= T4_PROD_WARP0) {
        // ================================ producer warps ==============================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T4_REGS_PRODUCER));
        const int pw = warp - T4_PROD_WARP0;               // rows 32*pw .. 32*pw+31 of every tile
        const int row = pw * 32 + lane;
        const int rows = (int)p.rows;
        const int nblk = D / 64;                           // 64-channel blocks of bf16 feature rows (D % 64 == 0 or D == 0)
        const int ch = lane & 7, rsub = lane >> 3;         // feature copies: lanes = (4 rows) x (8 chunks of 16 B)
        Tracer tr; tr.init(p.trace, 1, blockIdx.x == 0 && pw == 0 && lane == 0);
        int n_it_max = 0;
        for (int u = 0; u < U; ++u) n_it_max = max(n_it_max, n_it_of(u));
        auto load_index = [&](int tile) -> int {
            const int r = tile * TC_ROWS + row;
            return (tile < tiles && r < rows) ? __ldg(p.idx32 + r) : -1;
        };
        int n_nxt = load_index(first_tile);
        for (int it = 0; it < n_it_max; ++it) {
            for (int u = 0; u < U; ++u) {
                const int tile = first_tile + u + it * stride;
                if (tile >= tiles) continue;
                const int n = n_nxt;
                // index of this warp's rows in the next tile of the sequence (a tile past the end reads as -1)
                n_nxt = load_index(first_tile + (u + 1 < U ? u + 1 : 0) + (u + 1 < U ? it : it + 1) * stride);
                const bool valid = n >= 0 && n < N;        // n == N marks an empty ball (:85)
                const int g = (tile * TC_ROWS + row) >> logK;
                const int b = (int)((unsigned)g / (unsigned)S);
                const int grow = valid ? b * N + n : -1;   // global feature row of this lane's tile row
                unsigned char *A = sm.A_base + (size_t)u * p.a_unit_bytes;
                // coordinates (C <= 4): in flight together with the first feature batch, issued AFTER it (if the register
                // allocator spills them, the spill store blocks on the loads: better once the feature loads are out)
                float px[4], pc[4];
                auto load_xyz = [&]() {
                    if (!SMALL_C) return;
                    const float *pp = p.points + (int64_t)b * p.sb + (int64_t)n * p.sn;
                    const float *cc = p.centres + (int64_t)g * C;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        px[i] = pc[i] = 0.f;
                        if (valid && i < C) { px[i] = __ldg(pp); pc[i] = __ldg(cc + i); }
                        pp += p.sc;
                    }
                };
                if (nblk == 0) load_xyz();
                bool waited = (it == 0);
                auto wait_free = [&]() {
                    if (!waited) { mbar_wait(buf_free + u, (uint32_t)(it - 1) & 1u); waited = true; }
                };
                // ---- feature rows, two 64-channel blocks (16 loads of 16 B per lane) per batch ------------------------------
#pragma unroll 1
                for (int blk0 = 0; blk0 < nblk; blk0 += 2) {
                    const int nb = min(2, nblk - blk0);
                    uint4 v[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int gr = __shfl_sync(0xffffffffu, grow, 4 * i + rsub);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            v[2 * i + k] = make_uint4(0u, 0u, 0u, 0u);
                            if (k < nb && gr >= 0) v[2 * i + k] = ldg128(p.feat_bf16 + (int64_t)gr * D + (blk0 + k) * 64 + ch * 8);
                        }
                    }
                    if (blk0 == 0) load_xyz();
                    wait_free();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = pw * 32 + 4 * i + rsub;
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            if (k < nb)
                                *reinterpret_cast<uint4 *>(A + (blk0 + k) * SW128_BLOCK + r * 128 + ((ch ^ (r & 7)) << 4)) = v[2 * i + k];
                    }
                }
                wait_free();
                // ---- coordinates minus centre (+ zero padding) -> the unit's coordinate block [2][128][16 B] ----------------
                unsigned char *X_row = A + p.xyz_off + row * 16;
                if (SMALL_C) {
                    *reinterpret_cast<uint4 *>(X_row) = make_uint4(pack_bf16x2(__fsub_rn(px[0], pc[0]), __fsub_rn(px[1], pc[1])),
                                                                   pack_bf16x2(__fsub_rn(px[2], pc[2]), __fsub_rn(px[3], pc[3])), 0u, 0u);
                    *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(0u, 0u, 0u, 0u);
                } else {
                    const float *pp = p.points + (int64_t)b * p.sb + (int64_t)n * p.sn;
                    const float *cc = p.centres + (int64_t)g * C;
                    float x[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float a0 = 0.f, a1 = 0.f;
                        if (valid && i < C) { a0 = __ldg(pp); a1 = __ldg(cc + i); }
                        pp += p.sc;
                        x[i] = __fsub_rn(a0, a1);
                    }
                    *reinterpret_cast<uint4 *>(X_row) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                    *reinterpret_cast<uint4 *>(X_row + A_LBO) = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], x[15]));
                }
                const unsigned vm = __ballot_sync(0xffffffffu, valid);
                if (lane == 0) sm.vmask[8 * u + 4 * (it & 1) + pw] = vm;
                fence_proxy_async();                       // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(full + u);
                tr.mark(0x200 + u);
            }
        }
    } else {
        // ================================ worker warps ================================================================
        const int q = warp & 3, hb = warp >> 2, row = q * 32 + lane;
        const int gpt = TC_ROWS >> logK;
        const int G = (int)p.G;
        int wl[T2_MAX_UNITS], wit[T2_MAX_UNITS], nit[T2_MAX_UNITS];
        uint32_t acc_par = 0;
        int remaining = 0;
#pragma unroll
        for (int u = 0; u < T2_MAX_UNITS; ++u) { wl[u] = 0; wit[u] = 0; nit[u] = u < U ? n_it_of(u) : 0; remaining += nit[u] * L; }
        bool w_ready = false;
        Tracer tr; tr.init(p.trace, 0, blockIdx.x == 0 && warp == 0 && lane == 0);
        const uint32_t row_off = (uint32_t)(row * 128 + ((row & 7) << 4));
        const bool contig = p.ob == (int64_t)S * p.os;
        while (remaining > 0) {
#pragma unroll
            for (int u = 0; u < T2_MAX_UNITS; ++u) {
                if (wit[u] >= nit[u]) continue;
                if (static_sched) mbar_wait(acc_ready + u, (acc_par >> u) & 1u);
                else if (!mbar_probe(acc_ready + u, (acc_par >> u) & 1u, T4_POLL_NS)) continue;
                acc_par ^= 1u << u;
                tc_fence_after();
                const int l = wl[u], it = wit[u];
                const Tc2Layer &ly = p.ly[l];
                const uint32_t tmem_row = tmem_base + (uint32_t)(u * p.cols_per_unit) + ((uint32_t)(q * 32) << 16);
                tr.mark(0x400 + l * 16 + u);
                if (l < L - 1) {
                    // hidden layer: this warp's column share of its 32 rows -> ReLU -> bf16 -> next operand (in place)
                    const uint32_t a_rowx = smem_u32(sm.A_base) + (uint32_t)(u * p.a_unit_bytes) + row_off;
                    epi_share(ly.nb32, hb, tmem_row, a_rowx);
                    tc_fence_before();           // TMEM reads ordered before the next MMA's write to this region
                    fence_proxy_async();
                } else {
                    // last layer (transposed): this thread = channel mh*128 + row; columns = the tile's rows.
                    // 4-way split of the (128-channel slice, 32-row block) items of this lane quarter: whole groups stay
                    // in one warp
                    if (!w_ready) { mbar_wait(sm.w_bar, 0); w_ready = true; }     // the fp32 biases arrive with the weight image
                    const int tile = first_tile + u + it * stride;
                    const float *bs = sm.bias_s + ly.b_off;
                    const uint32_t *vm = sm.vmask + 8 * u + 4 * (it & 1);
                    PoolOut o;
                    o.all_valid = (vm[0] & vm[1] & vm[2] & vm[3]) == 0xffffffffu;
                    o.has_bf = p.out_bf16 != nullptr;
                    o.os = p.os; o.ob = p.ob; o.contig = contig; o.S = S; o.G = G;
                    const int g0 = tile * gpt;
                    const int n_mh = ly.n_mh;
                    const int bpg = K >= 32 ? K / 32 : 1;                   // 32-row blocks per group
                    int ipw = n_mh; if (ipw < bpg) ipw = bpg;               // items per warp
                    const int item0 = hb * ipw;
                    if (item0 < n_mh * 4) {
                        const int mh = item0 >> 2, blk_lo = item0 & 3, blk_hi = min(4, blk_lo + ipw);
                        for (int mhh = mh; mhh < n_mh && mhh * 4 < item0 + ipw; ++mhh) {
                            const int c = mhh * TC_ROWS + row;
                            o.bv = bs[c];
                            o.c_ok = c < ly.Cout;
                            const int b_lo = (mhh == mh) ? blk_lo : 0;
                            const int b_hi = (mhh == mh) ? blk_hi : min(4, item0 + ipw - mhh * 4);
                            o.gg = g0 + ((b_lo * 32) >> logK);
                            o.po = p.out + p.co_off + c;
                            o.pb = p.out_bf16 + p.co_off + c;
                            pool_slice_any<T4_POOL_PAIR>(logK, tmem_row + (uint32_t)(mhh * TC_ROWS), b_lo, b_hi, vm, o);
                        }
                    }
                    tc_fence_before();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(w_done + u);
                tr.mark(0x500 + l * 16 + u);
                if (++wl[u] == L) { wl[u] = 0; ++wit[u]; }
                --remaining;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == T4_MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

int launch_sa_tc4(const Tc2Params &p_in, size_t smem, cudaStream_t st)
{
    Tc2Params p = p_in;
    static const int static_env = getenv("PN2_TC4_STATIC") ? atoi(getenv("PN2_TC4_STATIC")) : 1;
    p.static_sched = static_env;
    int64_t grid = (p.tiles + p.U - 1) / p.U;
    if (grid > 148) grid = 148;
    auto run = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("sa_tc4: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        kern<<<(unsigned)grid, T4_THREADS, smem, st>>>(p);
        return launch_status("pn2_sa_fused(bf16, warp-specialised)");
    };
    return p.C <= 4 ? run(sa_tc4_kernel<true>) : run(sa_tc4_kernel<false>);
}

}  // namespace pn2
