/* morec_b200 C ABI  --  the drop-in boundary of the B200-native MoRec training step.
 *
 * The reference (westlake-repl/IDvs.MoRec) has no FFI: its hot path is Python calling eager PyTorch ops.  The
 * entry points below are what a maintainer's `model/` package binds (via ctypes) in place of
 * those eager op sequences; each one cites the reference site it replaces (paths relative to
 * /root/reference/inbatch_sasrec_e2e_text unless noted).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers + sizes, no torch types; every function returns 0 on success and a negative
 *     code on failure, with a message retrievable through morec_last_error() (thread-local).
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).  Calls are
 *     asynchronous on that stream, re-entrant per stream, and allocate nothing persistent.
 *   - outputs are pre-allocated by the caller.  All matrices are row-major; `ld*` are leading dimensions in
 *     elements.
 *   - dtype: 0 = fp32 storage, tensor-core math in TF32 (tcgen05 kind::tf32), fp32 accumulate  ("parity mode")
 *            1 = bf16 storage, tcgen05 kind::f16, fp32 accumulate                               ("fast mode")
 */
#ifndef MOREC_B200_H
#define MOREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOREC_ABI_VERSION 1

/* ---- library ------------------------------------------------------------------------------- */
int morec_abi_version(void);
const char* morec_last_error(void);
/* number of SMs of the current device (148 on B200); <0 when no device/driver is present */
int morec_device_sms(void);

/* ---- tcgen05 GEMM with fused epilogues -------------------------------------------------------
 * C[M,N] = epi( alpha * A . B^T (+ bias) )
 *   a_mn_major = 0: A is [M,K] row-major;  1: A is stored [K,M] row-major (i.e. A^T), read in place
 *   b_mn_major = 0: B is [N,K] row-major;  1: B is stored [K,N] row-major
 * so  forward  y  = x W^T + b   : A=x (0), B=W (0)            model/modules.py:14-17,52-63; HF BertModel linears
 *     dgrad    dx = dy W        : A=dy (0), B=W stored [N_out,K_in] = "[K,N]" of this GEMM (1)
 *     wgrad    dW += dy^T x     : A=dy stored [M,N_out] (1), B=x stored [M,K_in] (1), accumulate=1 (split-K,
 *                                 TMA reduce-add into fp32 dW)
 * epilogue: see MOREC_EPI_*.  C2 receives the pre-activation for MOREC_EPI_GELU (needed by the backward).
 * aux ([M,ldaux], same dtype as the operands) feeds the activation-gradient epilogues.
 * out_bf16 selects the element type of C/C2 (fp32 or bf16).  accumulate requires fp32 C and EPI_LINEAR.
 */
enum {
    MOREC_EPI_LINEAR = 0,        /* C = alpha*acc + bias                                                   */
    MOREC_EPI_GELU = 1,          /* C2 = alpha*acc + bias ; C = gelu_erf(C2)      (HF BertIntermediate)    */
    MOREC_EPI_GELU_NOSAVE = 2,   /* C = gelu_erf(alpha*acc + bias)                (encoders.py:69-70 eval) */
    MOREC_EPI_RELU = 3,          /* C = relu(alpha*acc + bias)                    (modules.py:16)          */
    MOREC_EPI_MUL_GELU_GRAD = 4, /* C = alpha*acc * gelu_erf'(aux)   aux = saved pre-activation            */
    MOREC_EPI_MUL_RELU_GRAD = 5  /* C = alpha*acc * (aux > 0)        aux = saved relu output               */
};
int morec_gemm(const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux, int M, int N,
               int K, int lda, int ldb, int ldc, int ldaux, int a_mn_major, int b_mn_major, int dtype, int out_bf16,
               int epilogue, float alpha, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOREC_B200_H */
