// TMA tensor maps for row gathers (cp.async.bulk.tensor ... tile::gather4, sm_100) and their device self-test.
// The pipelined fused kernel gathers neighbour feature rows (bf16, 128 bytes = 64 channels per block) with
// gather4: one instruction moves 4 arbitrary rows of a 2-D tensor into 4 consecutive 128-byte-swizzled rows of
// the MMA operand in shared memory, no SM registers, completion on an mbarrier.
#include <cuda.h>
#include <cudaTypedefs.h>
#include "pn2_tc.cuh"

namespace pn2 {

static PFN_cuTensorMapEncodeTiled get_encode()
{
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled)p;
    }
    return fn;
}

// 2-D view [rows][cols] of a contiguous bf16 tensor; box = 64 columns x 1 row, 128-byte swizzle; rows outside
// the tensor read as zero.  Returns 0 or a negative error (message set).
int make_row_gather_map(const void *base, int64_t rows, int cols, void *out_map /* CUtensorMap, 64-byte aligned */)
{
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) { set_error("tensor map: cuTensorMapEncodeTiled entry point not available"); return -2; }
    if (cols % 8 != 0 || ((uintptr_t)base & 15)) { set_error("tensor map: rows must be multiples of 16 bytes, 16-byte aligned"); return -1; }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)(cols < 64 ? cols : 64), 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc((CUtensorMap *)out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tensor map: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

// 2-D row-major tensor [rows][cols] of bf16 (elem_bytes 2) or fp32 (4) elements with a row pitch of `pitch_bytes` (multiple of
// 16); box = 128 bytes of columns x box_rows rows, 128-byte swizzle: the shared-memory image is the K-major SWIZZLE_128B MMA
// operand block (loads) / the staging tile of the GEMM epilogue (stores).  Elements outside the tensor read as zero and are
// not written.  atom32: the 128-byte swizzle with 32-byte atoms (32-byte chunk c of row r at position c ^ (r & 3)) -- the only
// layout the tensor core accepts for MN-major 32-bit operands (UMMA SWIZZLE_128B_BASE32B).
int make_tile_map(const void *base, int elem_bytes, int64_t rows, int64_t cols, int64_t pitch_bytes, int box_rows, void *out_map, int atom32)
{
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) { set_error("tensor map: cuTensorMapEncodeTiled entry point not available"); return -2; }
    if ((elem_bytes != 2 && elem_bytes != 4) || pitch_bytes % 16 != 0 || ((uintptr_t)base & 15) || box_rows < 1 || box_rows > 256) {
        set_error("tensor map: rows must be 16-byte aligned with a pitch that is a multiple of 16 bytes");
        return -1;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc((CUtensorMap *)out_map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                           const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tensor map: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

__global__ void __launch_bounds__(32, 1) tma_gather_selftest_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t *rows4, int n_inst,
                                                                   int col0, unsigned char *out)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    unsigned char *dst = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int lane = threadIdx.x;
    for (int i = lane; i < n_inst * 512 / 4; i += 32) reinterpret_cast<uint32_t *>(dst)[i] = 0xdeadbeefu;
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_expect_tx(&bar, (uint32_t)n_inst * 512u);
    __syncwarp();
    if (lane < n_inst)
        tma_gather4(dst + lane * 512, &tmap, col0, rows4[4 * lane], rows4[4 * lane + 1], rows4[4 * lane + 2], rows4[4 * lane + 3], &bar);
    mbar_wait(&bar, 0);
    for (int i = lane; i < n_inst * 512 / 4; i += 32) reinterpret_cast<uint32_t *>(out)[i] = reinterpret_cast<uint32_t *>(dst)[i];
}

int launch_selftest_tma_gather(const void *src_bf16, int64_t rows, int cols, const int32_t *rows4, int n_inst, int col0, void *out, cudaStream_t st)
{
    PN2_REQUIRE(src_bf16 && rows4 && out && n_inst >= 1 && n_inst <= 32, "pn2_selftest_tma_gather: bad arguments");
    alignas(64) CUtensorMap map;
    const int rc = make_row_gather_map(src_bf16, rows, cols, &map);
    if (rc) return rc;
    tma_gather_selftest_kernel<<<1, 32, n_inst * 512 + 1024, st>>>(map, rows4, n_inst, col0, (unsigned char *)out);
    return launch_status("pn2_selftest_tma_gather");
}

}  // namespace pn2
