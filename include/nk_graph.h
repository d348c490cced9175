/*
 * nk_graph.h -- C handle API of the host-side graph (C++), the mirror of the reference's
 * Var / VarDiff op surface over device tensors.
 *
 * The reference's host code is Rust (neuronika-variable/src/{var,vardiff,history,gradient}.rs);
 * no Rust toolchain exists in this environment, so the define-by-run graph above the kernel ABI
 * (nk_b200.h) is written in C++ (neuronika_b200/csrc/nk_graph.cpp) with the same semantics:
 *   - a variable is a handle to (data, tape); differentiable variables add (grad, backward tape)
 *     (var.rs:34-61, vardiff.rs:35-65);
 *   - op methods only record nodes; nothing is computed until forward() (lazy), which runs every
 *     node of the tape in creation order (var.rs:110-128); backward(seed) fills the root gradient
 *     with `seed` and runs the backward nodes in reverse order (vardiff.rs:125-141);
 *   - every backward node ACCUMULATES into its operands' gradients (beta = 1); leaf gradients
 *     persist until zero_grad() (vardiff.rs:100-102);
 *   - differentiability is sticky: Var (x) VarDiff -> VarDiff, and only the needed backward
 *     halves are built (var.rs:1048-1061).
 * This header exists so that Python (ctypes) and any other FFI can drive that C++ graph; a Rust
 * binding would not use it (it would implement Forward/Backward over nk_b200.h directly, see
 * INTEGRATION.md).
 *
 * All functions return 0 or a negative nk_status; nkg_last_error() gives the message (shape
 * errors carry the reference's own panic text where it has one).
 */
#ifndef NK_GRAPH_H
#define NK_GRAPH_H

#include "nk_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nkg_var nkg_var; /* Var or VarDiff */

typedef enum { NKG_MEAN = 0, NKG_SUM = 1 } nkg_reduction; /* neuronika-variable/src/lib.rs:29-36 */

const char* nkg_last_error(void);

/* ---- leaves (neuronika-variable/src/lib.rs:51-240 constructors; data zero-filled) ---- */
int nkg_leaf(nk_ctx* ctx, int ndim, const int64_t* shape, int dtype, nkg_var** out);
/* wrap caller-owned device memory as a leaf (flat parameter / gradient buckets for data parallel) */
int nkg_leaf_external(nk_ctx* ctx, int ndim, const int64_t* shape, int dtype, void* data_ptr, nkg_var** out);
/* Var::requires_grad (var.rs:104-107): returns a NEW differentiable handle sharing the data.
 * grad_dtype < 0 -> same as data; grad_ptr may supply caller-owned gradient storage (or NULL). */
int nkg_requires_grad(nkg_var* v, int grad_dtype, void* grad_ptr, nkg_var** out);
int nkg_clone(nkg_var* v, nkg_var** out);
int nkg_release(nkg_var* v);

/* ---- introspection ---- */
int nkg_is_diff(nkg_var* v);
int nkg_ndim(nkg_var* v);
int nkg_shape(nkg_var* v, int64_t* shape_out);
int nkg_dtype(nkg_var* v);
int nkg_grad_dtype(nkg_var* v);
void* nkg_data_ptr(nkg_var* v);  /* device pointer (allocates the buffer if still lazy) */
void* nkg_grad_ptr(nkg_var* v);  /* NULL for Var or after no_grad() */
int nkg_history_len(nkg_var* v); /* number of forward nodes on the tape (test.rs:748-806 checks) */
int nkg_backward_history_len(nkg_var* v);

/* ---- execution (var.rs:110-128, vardiff.rs:100-165) ---- */
int nkg_forward(nkg_var* v);
int nkg_backward(nkg_var* v, float seed);
int nkg_zero_grad(nkg_var* v);
int nkg_no_grad(nkg_var* v);
int nkg_with_grad(nkg_var* v);
/* host-side peephole fusion over the tape.  level 0: off.  level 1 (default): mm_t + bias add (+ ReLU) -> one GEMM
 * epilogue, conv + bias -> one kernel, gradient aliasing through single-consumer adds; invisible to results for ANY
 * use of the tape (a repeated backward() un-aliases first).  level 2: additionally the ReLU backward of a layer is
 * applied in the epilogue of the dX GEMM above it (nk_gemm_relu_bwd), which never stores the intermediate gradient --
 * exact for one backward() per tape (what a training loop does); a second backward() on such a tape fails loudly. */
int nkg_set_fusion(int level);

/* ---- operators (names follow the reference's methods) ---- */
int nkg_mm(nkg_var* a, nkg_var* b, nkg_var** out);      /* var.rs:1034-1061, vardiff.rs:1073-1106 */
int nkg_mm_t(nkg_var* a, nkg_var* b, nkg_var** out);    /* var.rs:1065-1094 */
int nkg_add(nkg_var* a, nkg_var* b, nkg_var** out);     /* vardiff.rs:902-924, broadcasting */
int nkg_relu(nkg_var* a, nkg_var** out);
int nkg_softmax(nkg_var* a, int axis, nkg_var** out);
int nkg_log_softmax(nkg_var* a, int axis, nkg_var** out);
int nkg_sum(nkg_var* a, nkg_var** out);
int nkg_mean(nkg_var* a, nkg_var** out);
int nkg_mse_loss(nkg_var* input, nkg_var* target, int reduction, nkg_var** out);
int nkg_nll_loss(nkg_var* input, nkg_var* target, int reduction, nkg_var** out);
int nkg_pad(nkg_var* a, int64_t ph, int64_t pw, float value, nkg_var** out); /* Zero = Constant(0) */
/* receiver is the KERNEL, argument the input, as in the reference (var.rs:704-716) */
int nkg_convolution(nkg_var* kernel, nkg_var* input, int64_t sh, int64_t sw, int64_t dh, int64_t dw,
                    int64_t groups, nkg_var** out);
/* (N, C, H, W) -> (N, C*H*W): bit-exact view; not in the reference (SURVEY.md 2.2 "missing") */
int nkg_flatten(nkg_var* a, nkg_var** out);

/* ---- the rest of the op surface: broadcasting arithmetic (vardiff.rs Sub/Mul/Div impls over
 * subtraction/, multiplication/, division/), unary maths (var.rs `exp`, `ln`, `sqrt`, `sigmoid`, `tanh`, `softplus`,
 * `leaky_relu`, `pow(i32)`, `Neg`), `t()` (reverses every axis), n-d padding with a mode, mv / vm / vv and 1-d / 3-d
 * convolution.  nkg_unary takes an nk_unary_op; the named functions are the reference's method names. */
int nkg_sub(nkg_var* a, nkg_var* b, nkg_var** out);
int nkg_mul(nkg_var* a, nkg_var* b, nkg_var** out);
int nkg_div(nkg_var* a, nkg_var* b, nkg_var** out);
int nkg_unary(nkg_var* a, int op, int iparam, nkg_var** out);
int nkg_neg(nkg_var* a, nkg_var** out);
int nkg_exp(nkg_var* a, nkg_var** out);
int nkg_ln(nkg_var* a, nkg_var** out);
int nkg_sqrt(nkg_var* a, nkg_var** out);
int nkg_sigmoid(nkg_var* a, nkg_var** out);
int nkg_tanh(nkg_var* a, nkg_var** out);
int nkg_softplus(nkg_var* a, nkg_var** out);
int nkg_leaky_relu(nkg_var* a, nkg_var** out);
int nkg_pow(nkg_var* a, int exp, nkg_var** out);
int nkg_transpose(nkg_var* a, nkg_var** out);
/* pad the nsp (1..3) sample dims of (N, C, ...) with an nk_pad_mode (pad/mod.rs:20-182) */
int nkg_pad_mode(nkg_var* a, int nsp, const int64_t* padding, int mode, float value, nkg_var** out);
int nkg_mv(nkg_var* matrix, nkg_var* vector, nkg_var** out);   /* matrix_vector_mul/mod.rs */
int nkg_vm(nkg_var* vector, nkg_var* matrix, nkg_var** out);   /* vector_matrix_mul/mod.rs */
int nkg_vv(nkg_var* a, nkg_var* b, nkg_var** out);             /* vector_vector_mul/mod.rs: 0-d result */
/* 1-d (N,C,L) / 3-d (N,C,D,H,W) convolution; the receiver is the kernel, as in nkg_convolution */
int nkg_convolution_nd(nkg_var* kernel, nkg_var* input, int nsp, const int64_t* stride, const int64_t* dilation,
                       int64_t groups, nkg_var** out);

/* ---- gradient-ready hook (data parallel overlap): `cb(user, begin, end)` is called from inside nkg_backward(), on
 * the calling thread, right after the LAST kernel that accumulates into elements [begin, end) of this leaf's gradient
 * in the running backward pass has been launched -- so the caller can start the all-reduce of that range while the
 * rest of backward runs.  Normally one call with [0, numel).  With row_chunks > 1, a matmul backward node that is the
 * last writer of the gradient computes it in that many row blocks (one GEMM each, same arithmetic per element) and
 * reports every block as soon as it is launched, so a single large layer's exchange overlaps its own dW GEMMs. */
typedef void (*nkg_grad_hook)(void* user, int64_t elem_begin, int64_t elem_end);
int nkg_set_grad_hook(nkg_var* leaf, nkg_grad_hook cb, void* user, int row_chunks);

/* ---- fused exchange (nk_b200.h "data-parallel gradient exchange"): when the matmul backward node that writes this
 * leaf's gradient finds it all-zero (beta = 0), it runs nk_gemm_rs into `slots` instead of the plain dW GEMM and calls
 * `cb(user, 1)`; otherwise (accumulating into an existing gradient, shape not shardable) it computes the gradient
 * locally as usual and calls `cb(user, 0)` so that the caller can fall back to an all-reduce.  world <= 8. */
typedef void (*nkg_grad_rs_hook)(void* user, int pushed);
int nkg_set_grad_rs(nkg_var* leaf, int world, int rank, void* const* slots, nkg_grad_rs_hook cb, void* user);

/* ---- SGD on a leaf (neuronika-optim/src/sgd/mod.rs:191-231) ---- */
int nkg_sgd_step(nkg_var* param, float* momentum_buf, float* master, float lr, float l2, float momentum,
                 float dampening, int nesterov, float grad_scale);

/* ---- Adam / AMSGrad / RMSProp / Adagrad on a leaf (neuronika-optim/src/{adam,amsgrad,rmsprop,adagrad}/mod.rs);
 * state arrays are caller-owned f32 device buffers of the parameter's size (see nk_b200.h nk_adam_step ...) */
int nkg_adam_step(nkg_var* param, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, float* master,
                  int64_t step, float lr, float beta1, float beta2, float eps, float l1, float l2, float grad_scale);
int nkg_rmsprop_step(nkg_var* param, float* square_avg, float* grad_avg, float* momentum_buf, float* master, float lr,
                     float alpha, float eps, float momentum, float l1, float l2, float grad_scale);
int nkg_adagrad_step(nkg_var* param, float* grad_sq, float* master, int64_t step, float lr, float lr_decay, float eps,
                     float l1, float l2, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif
