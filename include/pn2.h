/*
 * pn2.h -- C ABI of libpn2_sm100.so, the B200 (sm_100a) implementation of the PointNet++
 * set-abstraction path of Regen2001/PointNet-Like-Pose-Estimation (model/pointnet2_utils.py).
 *
 * The reference has no native code and therefore no FFI of its own: its "operator interface" for this
 * path is the Python surface of model/pointnet2_utils.py.  Each entry point below names the reference
 * function (file:line) whose arithmetic it replaces; the Python binding that the reference side adds is
 * shown in INTEGRATION.md and shipped as pointnet-like-pose-estimation_b200/_pn2_native.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never allocates,
 *     frees or synchronises; all launches go to `stream` (a cudaStream_t passed as void*).
 *   - coordinates ("points" in this fork) are addressed with int64 ELEMENT strides so that both
 *     layouts the models produce ([B,C,N] permuted, and contiguous [B,N,C]) are read in place; the
 *     stride of the last dim also selects the fp32 summation order that ATen would use.
 *   - return value: 0 ok; <0 argument error found on the host (pn2_last_error_string() explains);
 *     >0 a cudaError_t from the launch.
 *   - re-entrant and thread-safe: no global mutable state besides one-time function attributes and a
 *     thread-local error string.
 */
#ifndef PN2_H_
#define PN2_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_ABI_VERSION 15
#define PN2_MAX_COORD 16 /* exact summation-order rules are verified against ATen for C <= 16 */
#define PN2_MAX_LAYERS 8

/* channel order of a grouped row */
#define PN2_ORDER_SSG 0 /* [xyz - centre, feat]   pointnet2_utils.py:113-114 */
#define PN2_ORDER_MSG 1 /* [feat, xyz - centre]   pointnet2_utils.py:208-209 */
#define PN2_ORDER_ALL 2 /* [xyz (raw), feat]      pointnet2_utils.py:134-138 */

/* arithmetic of the shared MLP */
#define PN2_PREC_F32 0  /* fp32 FFMA, fp32 accumulate           (<= 1e-5 of the reference) */
#define PN2_PREC_BF16 1 /* bf16 tcgen05.mma, fp32 accumulate in TMEM (<= 2e-2)             */

int pn2_version(void);
const char *pn2_last_error_string(void);

/* farthest_point_sample -- pointnet2_utils.py:47-68.
 * points [B,N,C] with element strides (sb,sn,sc); start[B] = the torch.randint draw of :59 (int64, device,
 * values in [0,N); anything else is clamped into the cloud rather than read out of bounds).
 * Outputs (any may be NULL): idx64 [B,S] (the function's return value), idx32 [B,S], new_points [B,S,C]
 * contiguous (= index_points(points, idx), :107 / :201, fused because the cloud is already on chip). */
int pn2_fps_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                const int64_t *start, int S, int64_t *idx64, int32_t *idx32, float *new_points, void *stream);

/* index_points -- pointnet2_utils.py:28-45.  src [B,N,C] strided; idx [B,M] int64 (M = prod of idx dims 1..);
 * out [B,M,C] contiguous.  err_flag (int32, device, may be NULL) is set to 1 on an out-of-range index. */
int pn2_index_points_f32(const float *src, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                         const int64_t *idx, int64_t M, float *out, int32_t *err_flag, void *stream);

/* square_distance -- pointnet2_utils.py:5-26, the expanded form with the reference's operation order.
 * src [B,S,C] strided, dst [B,N,C] strided, out [B,S,N] contiguous. */
int pn2_square_distance_f32(const float *src, int64_t ssb, int64_t ssn, int64_t ssc,
                            const float *dst, int64_t dsb, int64_t dsn, int64_t dsc,
                            int B, int S, int N, int C, float *out, void *stream);

/* query_ball_point -- pointnet2_utils.py:70-90.  centres [B,S,C] strided (the reference always passes a
 * contiguous tensor), points [B,N,C] strided; r2 = (float)(radius**2) computed by the caller in double.
 * Outputs (either may be NULL): idx64 [B,S,K], idx32 [B,S,K].  empty_flag (int32, device, may be NULL) is
 * set to 1 if some centre has no neighbour (the reference then fails inside index_points); those rows are
 * filled with N. */
int pn2_ball_query_f32(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                       const float *points, int64_t sb, int64_t sn, int64_t sc,
                       int B, int S, int N, int C, float r2, int K,
                       int64_t *idx64, int32_t *idx32, int32_t *empty_flag, void *stream);

/* Materialised grouping -- sample_and_group / sample_and_group_all, pointnet2_utils.py:109-116, :134-140,
 * :205-211.  Only used by the stand-alone functions and by training mode: the fused kernels never write
 * this tensor.  feat may be NULL (D = 0).  idx32 [B,S,K] (ignored for PN2_ORDER_ALL, where S=1, K=N).
 * out [B,S,K,C+D] contiguous. */
int pn2_group_f32(const float *points, int64_t sb, int64_t sn, int64_t sc,
                  const float *feat, int64_t fb, int64_t fn, int64_t fc,
                  const int32_t *idx32, const float *centres,
                  int B, int N, int S, int K, int C, int D, int order, float *out, void *stream);

/* Differentiable feature gather of training mode -- index_points(feature, idx) of pointnet2_utils.py:112 / :208 and its
 * backward.  feat [B,N,D] with element strides (fb, fn, 1); idx32 [B,M] (M = S*K); out [B,M,D] contiguous.
 * pn2_scatter_add_rows_f32 adds grad [B,M,D] into grad_feat [B,N,D] (contiguous, zero-filled by the caller) with fp32
 * reductions; an index outside [0,N) (the empty-ball marker N) gathers a zero row and scatters nothing. */
int pn2_gather_rows_f32(const float *feat, int64_t fb, int64_t fn, int64_t fc, const int32_t *idx32, int B, int N, int64_t M, int D,
                        float *out, void *stream);
int pn2_scatter_add_rows_f32(const float *grad, const int32_t *idx32, int B, int N, int64_t M, int D, float *grad_feat, void *stream);

/* Upload of the FPS start indices (the torch.randint draw of pointnet2_utils.py:59, made on the host like the reference's):
 * `host_mapped` = n int64 in PINNED, device-mapped host memory (cudaHostAlloc under unified addressing: the host pointer is
 * valid on the device); a kernel on `stream` copies them to dst (device).  Unlike a cudaMemcpyAsync this does not queue on the
 * host-to-device copy engine behind a large input upload of the same process, and unlike a pageable copy it does not stall the
 * enqueueing thread.  The caller keeps the host buffer unchanged until the stream has passed this launch. */
int pn2_upload_i64(const int64_t *host_mapped, int64_t *dst, int64_t n, void *stream);

/* Fold eval-mode BatchNorm into the 1x1 convolutions and lay the weights out for the fused kernels
 * (pointnet2_utils.py:154-155, :168-170).  L layers; dims[L+1] = {Cin, C1..CL}.  For layer l the six
 * device pointers ptrs[6*l + {0..5}] = conv.weight [Cout,Cin], conv.bias, bn.weight, bn.bias,
 * bn.running_mean, bn.running_var.  pn2_packed_bytes gives the size of `packed` for `prec`.  C (coordinate
 * channels inside dims[0]) and `order` (PN2_ORDER_*) describe the reference's input-channel order of the
 * first layer; the tensor-core layout stores features first so that gathered rows are 16-byte aligned. */
int64_t pn2_packed_bytes(int L, const int *dims, int prec);
int pn2_pack_weights(int L, const int *dims, const void *const *ptrs, float eps, int prec, int C, int order,
                     void *packed, void *stream);

/* Fused set abstraction: gather -> shared MLP (folded BN, ReLU) -> max over the K neighbours, the
 * [B,S,K,Cin] tensor staying on chip -- PointNetSetAbstraction.forward :166-172 and one radius branch of
 * PointNetSetAbstractionMsg.forward :204-218.
 *   points [B,N,C] strided, feat [B,N,D] strided or NULL, centres [B,S,C] contiguous (unused for ALL),
 *   idx32 [B,S,K] (unused for ALL: S=1, K=N).
 *   out: pooled features, element (b,s,co) at out[b*ob + s*os + co_off + co]  (channels-last; co_off places
 *   an MSG branch inside the concatenated tensor of :222).  When K rows span several tiles the kernel
 *   combines with atomicMax and the caller must have zero-filled `out` (see pn2_sa_needs_zero_init). */
int pn2_sa_needs_zero_init(int S, int K, int L, const int *dims, int prec);
/* bf16 side tensors (PN2_PREC_BF16 only): in bf16 mode consecutive set-abstraction layers can hand the pooled
 * features over in bf16 as well -- out_bf16 (same addressing as out, element type bf16) is written next to the
 * fp32 result and a later launch may pass it as feat_bf16 ([B,N,D] contiguous, D % 8 == 0), which the gather
 * then copies straight into the MMA operand with cp.async.  Both may be NULL.  pn2_sa_bf16_side_io tells the
 * caller whether the kernel that serves this configuration WRITES out_bf16 (1) or not (0: pass NULL); feat_bf16 is
 * read by every bf16 kernel (the streaming kernel of the wide stacks included), so pass it whenever it exists. */
int pn2_sa_bf16_side_io(int K, int C, int order, int L, const int *dims, int prec);
int pn2_sa_fused(const float *points, int64_t sb, int64_t sn, int64_t sc,
                 const float *feat, int64_t fb, int64_t fn, int64_t fc,
                 const float *centres, const int32_t *idx32,
                 int B, int N, int S, int K, int C, int D, int order,
                 int L, const int *dims, const void *packed, int prec,
                 float *out, int64_t ob, int64_t os, int co_off,
                 const void *feat_bf16, void *out_bf16, void *stream);

/* Input preparation of the pose heads, fused -- provider.py:5-21 (normalization: subtract the centroid, divide by the
 * largest norm, per cloud) followed by provider.py:166-180 (splice_torch: append the one-hot category vector) and the
 * transpose of train_rotation.py:116.  xyz [B,N,3] with element strides; category[B] int64 (device) in [0,num_category);
 * out [B, 3+num_category, N] contiguous = the tensor the models consume.  num_category = 0 gives plain normalisation.
 * fp32 like the reference's numpy path; sums are tree reductions, so results agree to rounding (<= 1e-5 on coordinates in [-1, 1]), not bitwise. */
int pn2_normalize_splice_f32(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                             const int64_t *category, int num_category, float *out, void *stream);

/* query_ball_point for the n_radii balls of one PointNetSetAbstractionMsg layer in one scan -- the loop of
 * pointnet2_utils.py:203-204 calls :70-90 once per radius on the same centres and cloud; the distances do not depend on
 * the radius.  r2[n_radii], K[n_radii] are HOST arrays (r2 as in pn2_ball_query_f32); idx64[n_radii] / idx32[n_radii] are
 * host arrays of device pointers [B,S,K[r]] (either array, or single entries, may be NULL).  Each list is bit-identical
 * to a separate pn2_ball_query_f32 call; n_radii in 1..3.  Radii that are not ascending, widths other than 3 / 10 and
 * clouds beyond the one-pass shared-memory budget are served by one query per radius internally. */
int pn2_ball_query_multi_f32(const float *centres, int64_t csb, int64_t csn, int64_t csc,
                             const float *points, int64_t sb, int64_t sn, int64_t sc,
                             int B, int S, int N, int C, int n_radii, const float *r2, const int *K,
                             int64_t *const *idx64, int32_t *const *idx32, int32_t *empty_flag, void *stream);

/* The reference's numpy copies of farthest point sampling -- provider.py:183-204, data_utils/ModelDataLoader.py:10-31,
 * point_collect/transformer.py:120-141 (called per file by the data generators and loaders, and per DBSCAN cluster by
 * point_collect/collect.py:64-68).  They sample on the FIRST THREE columns of an [N,C] array in the array's dtype:
 * float64 for np.loadtxt / generated clouds, float32 where the loader casts first (ModelNetDataLoader.py:102-105).
 * points [B,N,>=3] with element strides (B clouds of equal N; the reference handles one per call); start[B] = the
 * np.random.randint(0, N) draw of :193 (int64, device); idx64 [B,S] = `centroids`.  Any N >= 1: clouds whose four
 * N-element arrays do not fit in shared memory need `scratch` of pn2_fps_xyz_scratch_bytes(B, N, sizeof(element))
 * bytes (0 = none needed).  Bit-exact with numpy: sub, square and the sequential 3-term sum rounded separately. */
int64_t pn2_fps_xyz_scratch_bytes(int B, int N, int elem_bytes);
int pn2_fps_xyz_f64(const double *points, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                    const int64_t *start, int S, int64_t *idx64, void *scratch, void *stream);
int pn2_fps_xyz_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                    const int64_t *start, int S, int64_t *idx64, void *scratch, void *stream);

/* Training-mode BatchNorm2d + ReLU of the shared MLP -- relu(bn(conv(x))) of pointnet2_utils.py:170 / :217 with batch
 * statistics (nn.BatchNorm2d in train(): biased variance over all B*K*S positions) -- on channels-last rows
 * y [M, C] = the 1x1 convolution's output WITHOUT its bias (a bias in front of a batch-statistics BatchNorm cancels).
 *   forward : stats [5][C] <- mean, biased var, invstd, a = gamma*invstd, b = beta - mean*a;  z = max(0, y*a + b).
 *             K = 0: z [M, C].  K > 0: the max over each group of K consecutive rows (:172 / :218) is fused:
 *             z [M/K, C] pooled maximum, arg [M/K, C] int32 = its first row, and the full z never reaches HBM.
 *             running_mean / running_var [C] (both or neither; may be NULL): nn.BatchNorm2d's bookkeeping in the same
 *             launch, running <- (1 - factor) * running + factor * batch, with the mean shifted by conv_bias [C] (the
 *             bias left out of y; may be NULL) and the variance unbiased (M / (M - 1)).
 *   backward: dz = gradient of z (K = 0: [M, C]; K > 0: [M/K, C] with the forward's arg); dy [M, C] = gradient of y;
 *             sums [2][C] <- dbeta, dgamma.  The ReLU mask is recomputed from y and stats.
 * ws = pn2_bn_train_ws_bytes(C) bytes of scratch (per-CTA fp64 partial sums, added in a fixed order: deterministic). */
int64_t pn2_bn_train_ws_bytes(int C);
int pn2_bn_relu_train_fwd_f32(const float *y, int64_t M, int C, int K, const float *gamma, const float *beta, float eps,
                              void *ws, float *stats, float *z, int32_t *arg,
                              float *running_mean, float *running_var, const float *conv_bias, float factor, void *stream);
int pn2_bn_relu_train_bwd_f32(const float *dz, const float *y, const int32_t *arg, int64_t M, int C, int K,
                              const float *stats, void *ws, float *sums, float *dy, void *stream);

/* Fused FC head, eval mode -- the tail of every model after the last set abstraction: pointnet2_cls_ssg.py:31-38
 * (fc1/bn1/relu/drop -> fc2/bn2/relu/drop -> fc3 -> log_softmax, argmax), rotation_ssg.py:32-38, translation_ssg.py:31-33,39-44
 * (+ mean, optionally through mean_fc1/mean_bn1/relu/mean_fc2), sign_ssg.py:29-36 (sigmoid, sign(. - 0.5)); the PointNet-v1
 * heads (width.py:39-44 ...) have the same shape.  One launch instead of ~14 library launches.
 *   x [B,K0] fp32 with row stride x_stride (the pooled features of the last set abstraction), K0 -> H1 -> H2 -> OUT.
 *   BatchNorm1d is folded by the caller: w1p / w2p = the folded weights W' [H][K] packed as [K/4][H][4] floats, b1 / b2 the
 *   folded biases; w3 [OUT][H2] row-major, b3 [OUT].  kind = PN2_HEAD_*:
 *     PLAIN  out [B,OUT] = fc3(.)                         CLS   out = log_softmax, aux = int64 [B] argmax (may be NULL)
 *     TRANS  out = fc3(.) + m, m = mean [B,3] (row stride mean_stride) or, when mean_mlp != NULL, W2 relu(W1' mean + b1') + b2
 *            with mean_mlp = 45 floats {W1' [6][3], b1' [6], W2 [3][6], b2 [3]} (BatchNorm1d(6) folded into W1', b1')
 *     SIGN   out = sigmoid(fc3(.)), aux = float [B,OUT] sign(out - 0.5) (may be NULL)
 *   Widths: H1, H2 = 8 x a divisor of 256; K0 >= 256; OUT <= 16 (the reference: 1024 -> 512 -> 256 -> 7 | 3 | 1). */
#define PN2_HEAD_PLAIN 0
#define PN2_HEAD_CLS 1
#define PN2_HEAD_TRANS 2
#define PN2_HEAD_SIGN 3
int pn2_fc_head_f32(const float *x, int64_t x_stride, int B, int K0, int H1, int H2, int OUT,
                    const float *w1p, const float *b1, const float *w2p, const float *b2,
                    const float *w3, const float *b3, int kind,
                    const float *mean, int64_t mean_stride, const float *mean_mlp,
                    float *out, void *aux, void *stream);

/* Row GEMM on the tensor cores: C[M,N] = A[M,K] . W[N,K]^T for tall operands (M = point or grouped rows) -- the 1x1
 * convolutions of the path where they are not fused into pn2_sa_fused: the per-point product of an MSG branch's factorised
 * first layer (pointnet2_utils.py:208-217 evaluated once per point) and the training-mode convolutions / input gradients
 * (pointnet2_utils.py:168-170, :214-217 under train_rotation.py:121-124).
 *   A [M,K] row stride lda, W [N,K] row stride ldw, C [M,N] row stride ldc (all in elements), dtypes PN2_DT_*.
 *   w_dtype = BF16: bf16 x bf16 -> fp32 accumulate (fp32 A rows are converted on load); K % 8 == 0.
 *   w_dtype = F32 : TF32 product of fp32 operands (the precision policy of the reference's cuDNN convolutions); K % 4 == 0.
 *   Rows of A, W and C must start on 16-byte boundaries.  N is arbitrary (<= 256 output columns per CTA, tiled beyond). */
#define PN2_DT_F32 0
#define PN2_DT_BF16 1
int pn2_rows_gemm(const void *A, int64_t lda, int a_dtype, const void *W, int64_t ldw, int w_dtype,
                  void *C, int64_t ldc, int c_dtype, int64_t M, int N, int K, void *stream);

/* Weight gradient of a 1x1 convolution on rows (training mode, backward of pointnet2_utils.py:170 / :217):
 *   dW[Cout,Cin] = dy[M,Cout]^T . x[M,Cin], TF32 product, fp32 accumulation in TMEM per row range, partials added in a fixed
 *   order (deterministic).  dy / x rows (strides ldy / ldx elements) start on 16-byte boundaries; Cout, Cin multiples of 4.
 *   ws = pn2_rows_wgrad_ws_bytes(M, Cout, Cin) bytes of scratch; dW [Cout][Cin] contiguous. */
int64_t pn2_rows_wgrad_ws_bytes(int64_t M, int Cout, int Cin);
int pn2_rows_wgrad_tf32(const float *dy, int64_t ldy, const float *x, int64_t ldx, int64_t M, int Cout, int Cin,
                        void *ws, float *dW, void *stream);

/* Device self-test of the tcgen05 building blocks (one 128xN tile against a scalar loop in the same
 * launch).  Returns 0 and writes max|err| to *max_err (device float). */
int pn2_selftest_umma(int n_cols, int k_dim, int variant, float *max_err, float *dump, void *stream);

/* Device self-test of the TMA row gather (cp.async.bulk.tensor tile::gather4 with a 128-byte-swizzled tensor map)
 * used for the bf16 feature gather: src [rows][cols] bf16 contiguous; rows4 (device) holds 4 row indices per
 * instruction, n_inst <= 32 instructions; out (device) receives n_inst * 512 bytes: the shared-memory image. */
int pn2_selftest_tma_gather(const void *src_bf16, int64_t rows, int cols, const int32_t *rows4, int n_inst, int col0,
                            void *out, void *stream);

/* Debug: stage timeline of the pipelined fused kernel.  buf = device buffer of 5 * 1024 uint64 (roles: unit 0..3,
 * MMA warp), zero-filled by the caller, or NULL to switch tracing off.  CTA 0 appends (event << 48) | clock64 per
 * pipeline stage (events: tools/tc2_trace.py).  Process-wide, not thread-safe; not for production use. */
int pn2_debug_trace(void *buf);

#ifdef __cplusplus
}
#endif
#endif /* PN2_H_ */
