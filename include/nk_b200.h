/*
 * nk_b200.h -- C ABI of the B200-native dense forward/backward hot path of neuronika.
 *
 * This is the drop-in boundary: one C function per
 * (operator, direction), taking device pointers and sizes, launched asynchronously on
 * the context's CUDA stream.  A Rust `Forward`/`Backward` node (reference trait objects,
 * neuronika-variable/src/autograd.rs:7-12, 20-25) keeps `Shared<CuArray>` handles exactly
 * like the reference's only device node does (cuda/cunode/binary_op/mod.rs:11-82) and
 * calls one of these functions from `forward()` / `backward()`.
 *
 * Conventions
 *   - every function returns 0 (NK_OK) or a negative nk_status; the message is available
 *     through nk_last_error().  Nothing aborts or throws across the ABI (the reference
 *     `.unwrap()`s every CUDA result, cuda/device.rs:36-45; the Rust wrapper does the same
 *     on our status codes).
 *   - all tensors are dense, C-order (row-major); images are NCHW (reference layout).
 *   - `beta` on a backward entry point selects the reference's accumulate protocol:
 *     beta = 1 -> `grad += ...` (what every reference Backward node does), beta = 0 ->
 *     overwrite (used by the host for a gradient buffer that is known to be all-zero,
 *     which gives identical results without the read).
 *   - element types: NK_F32 (reference type) and NK_BF16 (tensor-core operand type).
 *   - a context is single-threaded (like the reference's Rc graph, utils.rs:9); use one
 *     context per host thread / per GPU.
 *   - there is NO CPU fallback: every entry point fails with NK_ERR_CUDA when no device
 *     is present.
 */
#ifndef NK_B200_H
#define NK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nk_ctx nk_ctx;

typedef enum {
  NK_OK = 0,
  NK_ERR_INVALID_ARG = -1,
  NK_ERR_CUDA = -2,
  NK_ERR_NCCL = -3,
  NK_ERR_OOM = -4,
  NK_ERR_UNSUPPORTED = -5
} nk_status;

typedef enum { NK_F32 = 0, NK_BF16 = 1 } nk_dtype;

/* GEMM engine selection (nk_gemm_config): AUTO picks tcgen05 whenever the operands are
 * bf16 and TMA-addressable, else the SIMT kernel. */
typedef enum { NK_GEMM_AUTO = 0, NK_GEMM_SIMT = 1, NK_GEMM_TCGEN05 = 2 } nk_gemm_engine;
/* Convolution engine selection (nk_conv_config): AUTO = tensor-core kernels wherever they apply; DIRECT = the CUDA-core
 * kernels only (the parity path the reference's goldens run on); UNFUSED = AUTO without the one-pass dX + dW backward. */
typedef enum { NK_CONV_AUTO = 0, NK_CONV_DIRECT = 1, NK_CONV_UNFUSED = 2 } nk_conv_engine;

/* ---- context (replaces cuda::Device, neuronika-variable/src/cuda/device.rs:11-75) ---- */
int nk_ctx_create(int device, nk_ctx** out);
int nk_ctx_destroy(nk_ctx* ctx);
/* adopt an external cudaStream_t (e.g. torch's) so events / NCCL order with our kernels */
int nk_ctx_set_stream(nk_ctx* ctx, void* cuda_stream);
void* nk_ctx_stream(nk_ctx* ctx);
const char* nk_last_error(nk_ctx* ctx);
const char* nk_version(void);
int nk_sync(nk_ctx* ctx);
/* number of kernels this library launched on ctx since creation */
uint64_t nk_launch_count(nk_ctx* ctx);
int nk_sm_count(nk_ctx* ctx);
int nk_gemm_config(nk_ctx* ctx, int engine /* nk_gemm_engine */);
int nk_conv_config(nk_ctx* ctx, int engine /* nk_conv_engine */);
/* Optional tail split of the CTA-pair GEMM: the tiles of a last, at most half full wave are cut in two k halves on two
 * pairs each (the first half publishes its f32 accumulator rows, the second adds them in its epilogue), so that 256
 * tiles on 74 pairs take 3.5 tile times instead of 4.  Default: OFF -- at 4096^3 the hand-over (9 MB of partial sums
 * through L2 plus the wait) costs more than the half tile it saves (0.100 vs 0.094 ms), and the balanced grid leaves
 * 20 SMs to kernels of another stream (the data-parallel exchange).  Kept for shapes with long k loops. */
int nk_gemm_tail_split(nk_ctx* ctx, int enable);
/* name of the kernel variant the last nk_gemm call used ("tcgen05_nt_128x256", "simt", ...) */
const char* nk_last_gemm_kernel(nk_ctx* ctx);

/* ---- buffers (replaces cuda::CuArray, cuda/cuarray.rs:10-171) ---- */
int nk_alloc(nk_ctx* ctx, size_t bytes, void** dptr);      /* zero-filled, like CuArray::zeroed :35 */
int nk_alloc_uninit(nk_ctx* ctx, size_t bytes, void** dptr); /* for buffers the next kernel fully overwrites */
int nk_free(nk_ctx* ctx, void* dptr);
int nk_h2d(nk_ctx* ctx, void* dst, const void* src, size_t bytes);   /* from_ndarray :114 */
int nk_d2h(nk_ctx* ctx, void* dst, const void* src, size_t bytes);   /* as_ndarray :101 (blocks) */
int nk_d2d(nk_ctx* ctx, void* dst, const void* src, size_t bytes);
int nk_memset0(nk_ctx* ctx, void* dptr, size_t bytes);
int nk_host_alloc(nk_ctx* ctx, size_t bytes, void** hptr);           /* pinned host memory */
int nk_host_free(nk_ctx* ctx, void* hptr);
int nk_fill(nk_ctx* ctx, void* dptr, int dtype, size_t n, float value);
int nk_cast(nk_ctx* ctx, void* dst, int dst_dtype, const void* src, int src_dtype, size_t n);
/* ---- whole-step capture.  A training step launches the same kernels every iteration (the reference rebuilds the
 * same define-by-run graph each time, examples/quickstart.rs:216-227), so the step can be recorded once and replayed
 * with one driver call: between nk_capture_begin and nk_capture_end every kernel / copy / memset this library
 * enqueues on the context stream -- and on streams that join it through events, e.g. the exchange side stream -- is
 * recorded into a CUDA graph instead of running.  While capturing, nk_alloc / nk_alloc_uninit take memory from an
 * arena of `arena_bytes` owned by the graph (fixed addresses on every replay; blocks freed inside the capture are
 * recycled) and nk_free of arena memory is a no-op for as long as the graph lives.  Things that cannot be captured
 * (nk_d2h, nk_sync, growth of the scratch workspace) fail: run the step once eagerly first.  nk_graph_launch replays
 * the step on the context stream; results are in the same buffers every time. */
typedef struct nk_graph nk_graph;
int nk_capture_begin(nk_ctx* ctx, size_t arena_bytes);
int nk_capture_end(nk_ctx* ctx, nk_graph** out);
int nk_graph_launch(nk_ctx* ctx, nk_graph* graph);
int nk_graph_destroy(nk_ctx* ctx, nk_graph* graph);
int64_t nk_graph_kernel_count(nk_graph* graph);   /* kernel nodes per replay (counted into nk_launch_count) */
size_t nk_graph_arena_used(nk_graph* graph);
/* CUDA-event stopwatch on the context stream */
int nk_timer_start(nk_ctx* ctx);
int nk_timer_stop(nk_ctx* ctx, float* ms);

/* ---- matrix multiply: C = alpha * op(A) . op(B) + beta * C  (row-major, like ndarray's
 * general_mat_mul call sites: matrix_matrix_mul/mod.rs:33,65,97; matrix_matrix_mul_t/mod.rs:33,65,97)
 *   op(A) is M x K: transA=0 -> A stored (M,K) lda>=K ; transA=1 -> A stored (K,M) lda>=M
 *   op(B) is K x N: transB=0 -> B stored (K,N) ldb>=N ; transB=1 -> B stored (N,K) ldb>=K
 *   mm fwd: NN; mm dA: NT; mm dB: TN; mm_t fwd: NT; mm_t dX: NN; mm_t dW: TN.
 *   Optional fused epilogue: + bias[N] (row broadcast, Linear fwd) and ReLU. */
int nk_gemm(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
            const void* A, int64_t lda, const void* B, int64_t ldb, float beta, void* C,
            int64_t ldc, int ab_dtype, int c_dtype);
int nk_gemm_bias_act(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K,
                     float alpha, const void* A, int64_t lda, const void* B, int64_t ldb,
                     float beta, void* C, int64_t ldc, int ab_dtype, int c_dtype,
                     const void* bias /* N elements, c_dtype or f32 */, int bias_dtype,
                     int relu);

/* C = beta*C + relu'(relu_operand) (.) (op(A).op(B)): the input gradient of a matmul whose left operand is the output
 * of a ReLU, with that ReLU's backward (relu/mod.rs:71-78: dx += (x > 0) * g) applied in the GEMM epilogue instead of a
 * separate pass over the (M, N) gradient.  relu_operand is (M, N) with C's element type and leading dimension; either
 * the ReLU's input or its output (y = max(x, 0) > 0  <=>  x > 0). */
int nk_gemm_relu_bwd(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A,
                     int64_t lda, const void* B, int64_t ldb, float beta, void* C, int64_t ldc, int ab_dtype,
                     int c_dtype, const void* relu_operand);
/* The same with beta = 0 and, in the same epilogue, colsum[n] += sum_m C[m][n] (N floats, f32 atomics; the caller zeroes or
 * keeps them): the bias gradient of the layer below -- the un-broadcast of its Addition's right operand,
 * addition/mod.rs:81-135 -- without a separate pass over the (M, N) gradient.  The sums are those of the values as stored
 * (after rounding to C's element type).  NK_ERR_UNSUPPORTED (nothing done) where no engine has the fused epilogue.
 * The graph uses it at fusion level 3 (nkg_set_fusion); measured equal to the separate column-sum pass at config 4. */
int nk_gemm_relu_bwd_colsum(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A,
                            int64_t lda, const void* B, int64_t ldb, float beta, void* C, int64_t ldc, int ab_dtype,
                            int c_dtype, const void* relu_operand, float* colsum);

/* ---- broadcasting add (addition/mod.rs:39-50, 81-135; utils.rs:97-125, 152-192) ----
 * shapes are right-aligned, up to NK_MAX_DIMS dims. */
#define NK_MAX_DIMS 6
int nk_add_bcast_fwd(nk_ctx* ctx, void* y, const void* l, const void* r, int dtype,
                     int y_ndim, const int64_t* y_shape, int l_ndim, const int64_t* l_shape,
                     int r_ndim, const int64_t* r_shape);
/* dst = beta*dst + unbroadcast(g -> dst_shape) */
int nk_unbroadcast_acc(nk_ctx* ctx, void* dst, int dst_dtype, int dst_ndim,
                       const int64_t* dst_shape, const void* g, int g_dtype, int g_ndim,
                       const int64_t* g_shape, float beta);

/* ---- relu (relu/mod.rs:29-38, 67-79) ---- */
int nk_relu_fwd(nk_ctx* ctx, void* y, const void* x, size_t n, int dtype);
int nk_relu_bwd(nk_ctx* ctx, void* dx, const void* x, const void* g, size_t n, int dtype, float beta);

/* ---- softmax / log-softmax along one axis of an (outer, len, inner) view
 * (softmax/mod.rs:37-53, 84-104; logsoftmax/mod.rs:37-53, 84-102) ---- */
int nk_softmax_fwd(nk_ctx* ctx, void* y, const void* x, int64_t outer, int64_t len, int64_t inner, int dtype);
int nk_softmax_bwd(nk_ctx* ctx, void* dx, const void* y, const void* g, int64_t outer, int64_t len,
                   int64_t inner, int dtype, float beta);
int nk_log_softmax_fwd(nk_ctx* ctx, void* y, const void* x, int64_t outer, int64_t len, int64_t inner, int dtype);
int nk_log_softmax_bwd(nk_ctx* ctx, void* dx, const void* y, const void* g, int64_t outer, int64_t len,
                       int64_t inner, int dtype, float beta);

/* ---- losses and scalar reductions; scalar outputs / seeds are device f32 ----
 * (squared_error/mod.rs:46-58, 98-122; nll/mod.rs:42-68, 100-133; sum/mod.rs, mean/mod.rs) */
int nk_mse_fwd(nk_ctx* ctx, float* loss, const void* x, const void* t, size_t n, int dtype, int mean);
int nk_mse_bwd(nk_ctx* ctx, void* dx, const void* x, const void* t, const float* g, size_t n,
               int dtype, int mean, float beta);
/* NLL: `target` holds the class ids as floats (`target as usize`, nll/mod.rs:55) in element type
 * target_dtype -- NK_F32 whatever the input's type, or NK_BF16 (exact only for ids <= 256, so
 * rejected when c > 256). */
int nk_nll_fwd(nk_ctx* ctx, float* loss, const void* logp, const void* target, int target_dtype,
               int64_t n, int64_t c, int dtype, int mean);
int nk_nll_bwd(nk_ctx* ctx, void* dlogp, const void* target, int target_dtype, const float* g,
               int64_t n, int64_t c, int dtype, int mean, float beta);
int nk_sum_fwd(nk_ctx* ctx, float* out, const void* x, size_t n, int dtype, int mean);
int nk_sum_bwd(nk_ctx* ctx, void* dx, const float* g, size_t n, int dtype, int mean, float beta);

/* ---- the rest of the elementwise family (csrc/nk_pointwise.cu) ----
 * binary ops broadcast like nk_add_bcast_fwd (utils.rs:97-125):
 *   subtraction/mod.rs:44-49, multiplication/mod.rs:44-49, division/mod.rs:44-49.
 * nk_binary_bcast_bwd accumulates ONE operand's gradient (side 0 = left, 1 = right):
 *   dst = beta*dst + unbroadcast(factor(g, l, r) -> shape of that operand), factor =
 *   SUB: g | -g (subtraction/mod.rs:87-92,130-135); MUL: g*r | g*l (multiplication/mod.rs:90-148);
 *   DIV: g/r | -g*l/r^2 (division/mod.rs:90-151).  l / r may be NULL where the factor ignores them. */
typedef enum { NK_BIN_ADD = 0, NK_BIN_SUB = 1, NK_BIN_MUL = 2, NK_BIN_DIV = 3 } nk_binary_op;
int nk_binary_bcast_fwd(nk_ctx* ctx, int op, void* y, const void* l, const void* r, int dtype,
                        int y_ndim, const int64_t* y_shape, int l_ndim, const int64_t* l_shape,
                        int r_ndim, const int64_t* r_shape);
int nk_binary_bcast_bwd(nk_ctx* ctx, int op, int side, void* dst, int dst_dtype, const void* g,
                        const void* l, const void* r, int dtype, int l_ndim, const int64_t* l_shape,
                        int r_ndim, const int64_t* r_shape, float beta);
/* unary ops: y = f(x); dx = beta*dx + g * f'(saved), where `saved` is the tensor the reference's
 * Backward node keeps -- the OUTPUT y for EXP, SQRT, SIGMOID, TANH; the INPUT x for LN, SOFTPLUS,
 * LEAKY_RELU (slope 0.01), POWI; ignored for NEG.  iparam = the integer exponent of POWI
 * (negation/mod.rs:32-68, exp/mod.rs:32-74, logn/mod.rs:32-74, sqrt/mod.rs:32-74, sigmoid/mod.rs:32-76,
 * tanh/mod.rs:32-76, softplus/mod.rs:32-76, leaky_relu/mod.rs:33-81, power/mod.rs:41-88). */
typedef enum { NK_UN_NEG = 0, NK_UN_EXP = 1, NK_UN_LN = 2, NK_UN_SQRT = 3, NK_UN_SIGMOID = 4,
               NK_UN_TANH = 5, NK_UN_SOFTPLUS = 6, NK_UN_LEAKY_RELU = 7, NK_UN_POWI = 8 } nk_unary_op;
int nk_unary_fwd(nk_ctx* ctx, int op, void* y, const void* x, size_t n, int dtype, int iparam);
int nk_unary_bwd(nk_ctx* ctx, int op, void* dx, const void* saved, const void* g, size_t n, int dtype,
                 int iparam, float beta);
/* dst (reversed shape) = beta*dst + src^T : ndarray's `.t()` reverses every axis
 * (transpose/mod.rs:32-36 forward with beta = 0; :66-68 backward `dX += G^T` with beta = 1). Bit exact. */
int nk_transpose(nk_ctx* ctx, void* dst, int dst_dtype, const void* src, int src_dtype, int ndim,
                 const int64_t* src_shape, float beta);
/* padding of the last nsp (1..3) dims of (planes, s...) with a mode; backward is the interior slice
 * for every mode, as in the reference (pad/mod.rs:157-182). */
typedef enum { NK_PAD_CONSTANT = 0, NK_PAD_REFLECTIVE = 1, NK_PAD_REPLICATIVE = 2 } nk_pad_mode;
int nk_padnd_fwd(nk_ctx* ctx, void* y, const void* x, int64_t planes, int nsp, const int64_t* in_sp,
                 const int64_t* pad, int mode, float value, int dtype);
int nk_padnd_bwd(nk_ctx* ctx, void* dx, const void* g, int64_t planes, int nsp, const int64_t* in_sp,
                 const int64_t* pad, int dtype, float beta);

/* ---- matrix-vector / vector-matrix / vector-vector products (8-f rank 3; csrc/nk_gemv.cu) ----
 * A is (rows, cols) row-major.  trans = 0: y[rows] = beta*y + A.x[cols] (MatrixVectorMul::forward,
 * matrix_vector_mul/mod.rs:32-40; vm dv, vector_matrix_mul/mod.rs:64-72); trans = 1: y[cols] = beta*y +
 * A^T.x[rows] (VectorMatrixMul::forward :32-40; mv dv, matrix_vector_mul/mod.rs:93-101).
 * nk_outer_acc: A = beta*A + u (x) v (mv dA :64-69, vm dA vector_matrix_mul/mod.rs:96-101).
 * nk_dot: *out = <a, b> (vector_vector_mul/mod.rs:32-34); nk_scale_acc: dst = beta*dst + x * (*scalar)
 * with the scalar on the device (the 0-d gradient of vv, :58-63). */
int nk_gemv(nk_ctx* ctx, int trans, int64_t rows, int64_t cols, const void* A, const void* x, float beta,
            void* y, int ax_dtype, int y_dtype);
int nk_outer_acc(nk_ctx* ctx, void* A, int a_dtype, const void* u, const void* v, int64_t rows,
                 int64_t cols, int uv_dtype, float beta);
int nk_dot(nk_ctx* ctx, float* out, const void* a, const void* b, size_t n, int dtype);
int nk_scale_acc(nk_ctx* ctx, void* dst, int dst_dtype, const void* x, int x_dtype, const float* scalar,
                 size_t n, float beta);

/* ---- 2-D constant/zero padding of (planes, H, W) (pad/mod.rs:97-129, 157-182) ---- */
int nk_pad2d_fwd(nk_ctx* ctx, void* y, const void* x, int64_t planes, int64_t h, int64_t w,
                 int64_t ph, int64_t pw, float value, int dtype);
int nk_pad2d_bwd(nk_ctx* ctx, void* dx, const void* g, int64_t planes, int64_t h, int64_t w,
                 int64_t ph, int64_t pw, int dtype, float beta);

/* ---- 2-D convolution (cross-correlation, NCHW, no implicit padding)
 * (convolution/mod.rs:85-123 fwd, 146-189 dX, 191-226 dW; arg checks utils.rs:427-496).
 *   x (N,Cin,H,W)  w (Cout,Cin/groups,kh,kw)  y (N,Cout,Ho,Wo),
 *   Ho = (H - dh*(kh-1) - 1)/sh + 1.
 *   fwd optionally fuses + bias[Cout] (the Conv2d layer's (Cout,1,1) bias,
 *   neuronika-nn/src/lib.rs:774) and ReLU; bwd_kernel optionally also accumulates
 *   dbias[Cout] += sum_{n,p,q} g. */
int nk_conv2d_fwd(nk_ctx* ctx, void* y, const void* x, const void* w, const void* bias, int relu,
                  int64_t n, int64_t cin, int64_t h, int64_t wd, int64_t cout, int64_t kh,
                  int64_t kw, int64_t sh, int64_t sw, int64_t dh, int64_t dw, int64_t groups,
                  int dtype);
int nk_conv2d_bwd_input(nk_ctx* ctx, void* dx, const void* g, const void* w, int64_t n,
                        int64_t cin, int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw,
                        int64_t sh, int64_t sw, int64_t dh, int64_t dw, int64_t groups, int dtype,
                        float beta);
int nk_conv2d_bwd_kernel(nk_ctx* ctx, void* dwt, int dw_dtype, void* dbias, const void* g,
                         const void* x, int64_t n, int64_t cin, int64_t h, int64_t wd,
                         int64_t cout, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t dh,
                         int64_t dw, int64_t groups, int dtype, float beta);
/* Both halves of ConvolutionBackward::backward (convolution/mod.rs:146-226, which runs the input and
 * the kernel half back to back) in one call: dx = beta_dx*dx + ..., dw = beta_dw*dw + ...,
 * dbias likewise (or NULL).  Where the tensor-core path applies the output gradient is streamed
 * ONCE for both; otherwise it is the two calls above in sequence.  Same results either way. */
int nk_conv2d_bwd(nk_ctx* ctx, void* dx, float beta_dx, void* dwt, int dw_dtype, void* dbias,
                  float beta_dw, const void* g, const void* x, const void* w, int64_t n,
                  int64_t cin, int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw,
                  int64_t sh, int64_t sw, int64_t dh, int64_t dw, int64_t groups, int dtype);
/* The same when the output gradient is ONE value everywhere -- VarDiff::backward(seed) on the convolution's own
 * output fills the root gradient with the seed (vardiff.rs:125-141): the kernel synthesises its G tiles from g_value
 * (rounded to the gradient's element type first, as the fill would) instead of reading 2|G| bytes that a fill kernel
 * would have had to write first.  Same arithmetic, same results as fill + nk_conv2d_bwd.  Returns NK_ERR_UNSUPPORTED
 * (nothing done) for shapes outside the tensor-core engine: the caller then materialises the gradient. */
int nk_conv2d_bwd_uniform(nk_ctx* ctx, void* dx, float beta_dx, void* dwt, int dw_dtype, void* dbias,
                          float beta_dw, float g_value, const void* x, const void* w, int64_t n, int64_t cin,
                          int64_t h, int64_t wd, int64_t cout, int64_t kh, int64_t kw, int64_t sh, int64_t sw,
                          int64_t dh, int64_t dw, int64_t groups, int dtype);
/* 1-D / 3-D convolution, x (N, Cin, s[0..nsp)), w (Cout, Cin/groups, k[0..nsp)), nsp = 1..3 sample dims
 * (the reference's convolution is generic over them: convolution/mod.rs:85-226, goldens
 * convolution/test.rs:144-239, 306-444 and the strided / dilated / grouped siblings).  CUDA-core gather
 * kernels; the 2-D entry points above are the tensor-core path. */
int nk_convnd_fwd(nk_ctx* ctx, void* y, const void* x, const void* w, int nsp, int64_t n, int64_t cin,
                  const int64_t* in_sp, int64_t cout, const int64_t* k, const int64_t* stride,
                  const int64_t* dilation, int64_t groups, int dtype);
int nk_convnd_bwd_input(nk_ctx* ctx, void* dx, const void* g, const void* w, int nsp, int64_t n,
                        int64_t cin, const int64_t* in_sp, int64_t cout, const int64_t* k,
                        const int64_t* stride, const int64_t* dilation, int64_t groups, int dtype,
                        float beta);
int nk_convnd_bwd_kernel(nk_ctx* ctx, void* dwt, int dw_dtype, const void* g, const void* x, int nsp,
                         int64_t n, int64_t cin, const int64_t* in_sp, int64_t cout, const int64_t* k,
                         const int64_t* stride, const int64_t* dilation, int64_t groups, int dtype,
                         float beta);
/* name of the kernel variant the last conv call used */
const char* nk_last_conv_kernel(nk_ctx* ctx);

/* ---- SGD (neuronika-optim/src/sgd/mod.rs:191-231, penalty.rs:63-67) ----
 *   g' = grad_scale*g + 2*l2*w ; no momentum: w -= lr*g' ;
 *   momentum: buf = mu*buf + (1-damp)*g' ; w -= lr*(nesterov ? g' + mu*buf : buf).
 *   `buf` (f32, n elements) may be NULL when momentum <= FLT_EPSILON.
 *   `master` (f32, optional) keeps an f32 copy of bf16 weights: the update is applied to
 *   master and w receives its rounding.  grad_scale = 1/world_size for data parallel. */
int nk_sgd_step(nk_ctx* ctx, void* w, int w_dtype, void* g, int g_dtype, float* buf, float* master,
                size_t n, float lr, float l2, float momentum, float dampening, int nesterov,
                float grad_scale, int write_back_grad);

/* ---- Adam / AMSGrad / RMSProp / Adagrad as single fused passes (8-f rank 2; csrc/nk_optim.cu) ----
 * Common to all: g' = grad_scale*g + l1*signum(w) + 2*l2*w (penalty.rs:63-79: L1, L2, ElasticNet), written
 * back into g when write_back_grad (the reference adds the penalty into the gradient); state arrays are
 * f32, n elements, zero before the first step; `master` as in nk_sgd_step.
 *   adam     m = b1*m + (1-b1)g'; v = b2*v + (1-b2)g'^2; w -= m / (sqrt(v)/sqrt(1-b2^t) + eps) * lr/(1-b1^t)
 *            (adam/mod.rs:131-169); max_exp_avg_sq != NULL -> AMSGrad: v^ = max(v^, v) replaces v in the
 *            denominator (amsgrad/mod.rs:159-204).  `step` = t, counted from 1.
 *   rmsprop  s = a*s + (1-a)g'^2; centered (grad_avg != NULL): ga = a*ga + (1-a)g', denom = sqrt(s - ga^2)+eps
 *            else sqrt(s)+eps; momentum (> f32::EPSILON, buffer != NULL): b = mu*b + g'/denom, w -= lr*b;
 *            else w -= g'/denom*lr (rmsprop/mod.rs:193-300).
 *   adagrad  s += g'^2; w -= g' / (sqrt(s) + eps) * lr / (1 + (t-1)*lr_decay)   (adagrad/mod.rs:113-140). */
int nk_adam_step(nk_ctx* ctx, void* w, int w_dtype, void* g, int g_dtype, float* exp_avg,
                 float* exp_avg_sq, float* max_exp_avg_sq, float* master, size_t n, int64_t step, float lr,
                 float beta1, float beta2, float eps, float l1, float l2, float grad_scale,
                 int write_back_grad);
int nk_rmsprop_step(nk_ctx* ctx, void* w, int w_dtype, void* g, int g_dtype, float* square_avg,
                    float* grad_avg, float* momentum_buf, float* master, size_t n, float lr, float alpha,
                    float eps, float momentum, float l1, float l2, float grad_scale, int write_back_grad);
int nk_adagrad_step(nk_ctx* ctx, void* w, int w_dtype, void* g, int g_dtype, float* grad_sq, float* master,
                    size_t n, int64_t step, float lr, float lr_decay, float eps, float l1, float l2,
                    float grad_scale, int write_back_grad);

/* ---- NCCL all-reduce behind the ABI (csrc/nk_comm.cu) ----
 * The context owns the communicator; libnccl.so.2 is bound at run time (dlopen), so the library loads
 * without it.  Rank 0 creates the 128-byte id and ships it to the others by any means; every rank then
 * calls nk_comm_init_rank (collective).  nk_allreduce_sum sums `n` elements in place over the replicas,
 * enqueued on the context stream (ordered with the kernels that produced the gradients and with the
 * nk_sgd_step that follows).  Errors: NK_ERR_NCCL. */
int nk_comm_unique_id(nk_ctx* ctx, void* id128);
int nk_comm_init_rank(nk_ctx* ctx, int world, int rank, const void* id128);
int nk_comm_destroy(nk_ctx* ctx);
int nk_comm_world(nk_ctx* ctx);
int nk_comm_rank(nk_ctx* ctx);
int nk_allreduce_sum(nk_ctx* ctx, void* ptr, size_t n, int dtype);

/* ---- data-parallel gradient exchange over NVLink peer memory ----
 * The reference has no multi-device path; under data parallel the only exchange on the hot path is
 * the sum of the weight gradients over the replicas before the SGD step above (which then runs
 * identically on every replica).  Rather than a library all-reduce after the dW GEMM, the exchange
 * is fused into the kernels around it (neuronika_b200/csrc/nk_peer.cu):
 *   nk_gemm_rs       the tcgen05 GEMM whose epilogue stores row shard o of the (M,N) f32 product
 *                    into rank o's slot buffer `slots[o]` (world*M/world*N floats, slot index = the
 *                    calling rank) over NVLink, tile by tile while the MMAs run (reduce-scatter);
 *   nk_peer_barrier  flag exchange through peer memory: returns (on the stream) once every rank has
 *                    reached the same `epoch`; `flags[r]` is rank r's flag array (>= world words);
 *   nk_reduce_bcast  the owner sums its `world` slots in rank order and stores the result into every
 *                    replica's gradient `grads[r]` at element rank*shard_elems (all-gather half).
 * Buffers that peers touch come from nk_ipc_alloc (cudaMalloc; pool memory cannot be exported) and
 * are mapped by the other processes with nk_ipc_export (64-byte handle) / nk_ipc_open. */
int nk_ipc_alloc(nk_ctx* ctx, size_t bytes, void** out);
int nk_ipc_free(nk_ctx* ctx, void* ptr);
int nk_ipc_export(nk_ctx* ctx, void* ptr, void* handle64);
int nk_ipc_open(nk_ctx* ctx, const void* handle64, void** peer_ptr);
int nk_ipc_close(nk_ctx* ctx, void* peer_ptr);
int nk_peer_barrier(nk_ctx* ctx, void* const* flags, int world, int rank, uint32_t epoch);
int nk_gemm_rs(nk_ctx* ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, float alpha,
               const void* A, int64_t lda, const void* B, int64_t ldb, void* const* slots, int world,
               int rank, int ab_dtype);
int nk_reduce_bcast(nk_ctx* ctx, const float* slots, void* const* grads, int world, int rank,
                    int64_t shard_elems, int max_ctas);
/* barrier -> owner reduce + broadcast -> barrier as ONE kernel whose epoch lives on the device (`state`: 16 bytes of
 * zero-initialised LOCAL device memory per exchange sequence), so the launch is identical every step and can be
 * captured in a CUDA graph.  flags[r] = rank r's flag words (>= 2*world, zero-initialised, peer-mapped).  Work is
 * handed out in 32 KB chunks, so max_ctas may be the SM count (0 = that): CTAs that do not fit beside a running
 * GEMM start when it retires.  All ranks must issue the same sequence of calls on the same (flags, state). */
int nk_reduce_exchange(nk_ctx* ctx, const float* slots, void* const* grads, void* const* flags, int world,
                       int rank, int64_t shard_elems, void* state, int max_ctas);
/* all-reduce (sum, rank order: bit-identical on every replica) of a small local vector `grad` (n <= 2^20 floats: biases,
 * the 10-wide layer) through peer memory in one single-CTA kernel: slots[r] = rank r's receive buffer (world*n floats),
 * flags / state as above but separate from nk_reduce_exchange's. */
int nk_peer_allreduce_small(nk_ctx* ctx, float* grad, void* const* slots, void* const* flags, int world, int rank,
                            int64_t n, void* state);

#ifdef __cplusplus
}
#endif
#endif /* NK_B200_H */
