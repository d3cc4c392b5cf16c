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
 *   - dtype (GEMM-bearing calls): 0 = fp32 storage, one tcgen05 kind::tf32 pass, fp32 accumulate     ("tf32")
 *            1 = bf16 storage, tcgen05 kind::f16, fp32 accumulate                                   ("bf16")
 *            2 = fp32 storage, error-compensated 3xTF32 on tcgen05 (hi*hi + hi*lo + lo*hi, operands split in
 *                shared memory), fp32-grade results: the PARITY mode checked against the CPU oracle  ("fp32")
 *            3 = fp16 storage, tcgen05 kind::f16 on IEEE-half operands, fp32 accumulate              ("fp16":
 *                the arithmetic of the reference's own CUDA path, torch.cuda.amp.autocast, run.py:242)
 *     element-wise / LayerNorm kernels only distinguish the storage type (0 or 2 = fp32, 1 = bf16, 3 = fp16).  The
 *     attention entry points take the same codes: 2 runs the exact-fp32 SIMT kernels (parity mode), 0 / 1 / 3 run
 *     the tensor-core kernels (TF32 resp. bf16 / fp16 mma, fp32 softmax) when the shape is covered (<= 64 tokens
 *     with 32- or 64-wide heads, <= 32 tokens with 256-wide heads) and fall back to the SIMT kernels otherwise.
 */
#ifndef MOREC_B200_H
#define MOREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOREC_ABI_VERSION 2

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
 * epilogue: see MOREC_EPI_*.  C2 receives the pre-activation for MOREC_EPI_GELU, the activation derivative for
 * MOREC_EPI_GELU_DGELU (needed by the backward).
 * aux ([M,ldaux], same dtype as the operands) feeds the activation-gradient epilogues.
 * out_dtype selects the element type of C/C2 (0 fp32, 1 bf16, 3 fp16).  accumulate requires fp32 C and EPI_LINEAR.
 */
enum {
    MOREC_EPI_LINEAR = 0,        /* C = alpha*acc + bias                                                   */
    MOREC_EPI_GELU = 1,          /* C2 = alpha*acc + bias ; C = gelu_erf(C2)      (HF BertIntermediate)    */
    MOREC_EPI_GELU_NOSAVE = 2,   /* C = gelu_erf(alpha*acc + bias)                (encoders.py:69-70 eval) */
    MOREC_EPI_RELU = 3,          /* C = relu(alpha*acc + bias)                    (modules.py:16)          */
    MOREC_EPI_MUL_GELU_GRAD = 4, /* C = alpha*acc * gelu_erf'(aux)   aux = saved pre-activation            */
    MOREC_EPI_MUL_RELU_GRAD = 5, /* C = alpha*acc * (aux > 0)        aux = saved relu output               */
    MOREC_EPI_GELU_DGELU = 6,    /* z = alpha*acc + bias ; C = gelu_erf(z) ; C2 = gelu_erf'(z): the forward saves
                                    the activation DERIVATIVE instead of the pre-activation (they share the
                                    exponential), so the backward epilogue is a plain multiply             */
    MOREC_EPI_MUL_AUX = 7        /* C = alpha*acc * aux              aux = C2 of MOREC_EPI_GELU_DGELU      */
};
int morec_gemm(const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux, int M, int N,
               int K, int lda, int ldb, int ldc, int ldaux, int a_mn_major, int b_mn_major, int dtype, int out_dtype,
               int epilogue, float alpha, int accumulate, void* stream);

/* ---- LayerNorm with fused residual / position add and dropout ---------------------------------
 * forward : z = dropout_pre(x) (+ residual) (+ pos[row % pos_period]);  y_pre = LN(z)*gamma + beta;
 *           y = dropout_post(y_pre)          rstd[M] saved for the backward; y_pre only needed if p_post > 0
 *   replaces  LayerNorm(residual + dropout(sublayer(x)))  model/modules.py:16-17,62-63 and HF BertSelfOutput/
 *             BertOutput;  dropout(LayerNorm(x + position_embedding))  model/modules.py:89-93; HF BertEmbeddings
 * backward: dy(+dy2) -> dz (grad of z; feeds the residual stream) and dx_branch (= dropout_pre mask applied to
 *           dz; may be null when p_pre == 0, then dz is also the branch gradient).  `y` is the tensor LN produced
 *           BEFORE post-dropout (y_pre when p_post > 0).  dgamma/dbeta/dbias/dpos are ACCUMULATED (atomicAdd):
 *           dbias[H] += colsum(dx_branch) (bias grad of the producing linear), dpos[pos_period,H] += dz rows.
 * dropout masks are regenerated from (seed, offset) with Philox4x32-10; H % 4 == 0, H <= 2048.
 */
int morec_layernorm_fwd(const void* x, const void* residual, const float* pos, int pos_period, const float* gamma,
                        const float* beta, void* y, void* y_pre, float* rstd, int M, int H, float eps, int dtype,
                        float p_pre, float p_post, uint64_t seed, uint64_t off_pre, uint64_t off_post, void* stream);
int morec_layernorm_bwd(const void* dy, const void* dy2, const void* y, const float* gamma, const float* beta,
                        const float* rstd, void* dz, void* dx_branch, float* dgamma, float* dbeta, float* dbias,
                        float* dpos, int pos_period, int M, int H, int dtype, float p_pre, float p_post, uint64_t seed,
                        uint64_t off_pre, uint64_t off_post, void* stream);

/* ---- short-sequence multi-head attention (<= 32 tokens per sequence) ---------------------------
 * q/k/v (and dq/dk/dv): [n_rows, ld], o (and dO): [n_rows, ld_o], head h in columns [h*head_dim, (h+1)*head_dim).  Sequences are either packed
 * (cu_seqlens[n_seq+1], rows cu[s]..cu[s+1]) or fixed length `seqlen` (rows s*seqlen..).  Score mask:
 * causal (k <= q) and/or key_mask[n_seq, seqlen] != 0, applied ADDITIVELY with `masked_add` (-1e9 in the reference,
 * model/encoders.py:27) so fully-masked rows reproduce the reference's uniform softmax.  Dropout acts on the
 * probabilities (model/modules.py:30).  head_dim % 4 == 0.
 *   replaces HF BertSelfAttention (call site model/encoders.py:68) and SelfAttention.forward model/modules.py:27-31
 */
int morec_attn_fwd(const void* q, const void* k, const void* v, void* o, const int32_t* cu_seqlens,
                   const float* key_mask, int causal, int n_seq, int seqlen, int n_heads, int head_dim, int ld,
                   int ld_o, float scale, float masked_add, int dtype, float dropout_p, uint64_t seed,
                   uint64_t offset, void* stream);
int morec_attn_bwd(const void* q, const void* k, const void* v, const void* d_o, void* dq, void* dk, void* dv,
                   const int32_t* cu_seqlens, const float* key_mask, int causal, int n_seq, int seqlen, int n_heads,
                   int head_dim, int ld, int ld_o, float scale, float masked_add, int dtype, float dropout_p,
                   uint64_t seed, uint64_t offset, void* stream);

/* ---- general attention, sequences / windows of up to 128 tokens, optional additive bias and mask ----------------
 * Swin window attention (HF SwinSelfAttention; call site inbatch_sasrec_e2e_vision/model/encoders.py:31): L = 49,
 * bias[n_heads, L, L] = relative-position bias gathered by the caller, mask[n_mask, L, L] = shifted-window mask
 * (0 / -100; window s uses mask s % n_mask); and BERT with long titles (cfg-2, T = 128) packed by cu_seqlens.
 * score = q.k * scale + bias + mask; dropout on the probabilities; backward also accumulates dbias (fp32, atomics).
 */
int morec_attn_gen_fwd(const void* q, const void* k, const void* v, void* o, const int32_t* cu_seqlens,
                       const float* bias, const float* mask, int n_mask, int n_seq, int seqlen, int n_heads,
                       int head_dim, int ld, int ld_o, float scale, int dtype, float dropout_p, uint64_t seed,
                       uint64_t offset, void* stream);
int morec_attn_gen_bwd(const void* q, const void* k, const void* v, const void* d_o, void* dq, void* dk, void* dv,
                       float* dbias, const int32_t* cu_seqlens, const float* bias, const float* mask, int n_mask,
                       int n_seq, int seqlen, int n_heads, int head_dim, int ld, int ld_o, float scale, int dtype,
                       float dropout_p, uint64_t seed, uint64_t offset, void* stream);

/* ---- in-batch debiased softmax cross-entropy (model/model.py:45-67) ---------------------------
 * morec_inbatch_mask   : integer pre-pass.  member[B, ceil(C/32)] bit c = (col_ids[c] in row_ids[b, 0..L]);
 *                        pad[ceil(C/32)] bit c = (col_ids[c] == 0).  Bit-exact restatement of model.py:51-63.
 * morec_inbatch_ce_fwd : scoring GEMM P[B*L,D] . E[C,D]^T on tcgen05 with the CE epilogue (logits stay in TMEM):
 *                        S = P.E^T - log_pop[c]; masked(r,c) = pad(c) | (member(b,c) & c != target(r)) -> -1e4;
 *                        target(r) = col_offset + b*(L+1) + j + 1.  Workspaces part_m/part_l: [B*L, NT] with
 *                        NT = morec_inbatch_ce_num_tiles(C).  Outputs row_lse[B*L], row_loss[B*L] (0 for invalid
 *                        rows), sum_cnt[2] = {sum of valid row losses, number of valid rows}, loss = mean (may be
 *                        null, e.g. when the caller all-reduces sum_cnt across ranks first).
 * morec_inbatch_ce_dlogits : dS[B*L, ldds] = (softmax(S) - onehot(target)) * valid(r) * grad_out / n_valid
 *                        (grad_out, n_valid: device scalars).  dP = dS.E and dE = dS^T.P then run on morec_gemm.
 */
int morec_inbatch_mask(const int64_t* row_ids, const int64_t* col_ids, uint32_t* member, uint32_t* pad, int B, int L,
                       int C, void* stream);
int morec_inbatch_ce_num_tiles(int C, int dtype);
int morec_inbatch_ce_fwd(const void* P, const void* E, const uint32_t* member, const uint32_t* pad,
                         const float* log_pop, const float* log_mask, int B, int L, int D, int C, int col_offset,
                         int dtype, float* part_m, float* part_l, float* tgt_logit, float* row_lse, float* row_loss,
                         float* sum_cnt, float* loss, void* stream);
int morec_inbatch_ce_dlogits(const void* P, const void* E, const uint32_t* member, const uint32_t* pad,
                             const float* log_pop, const float* log_mask, const float* row_lse, const float* grad_out,
                             const float* n_valid, int B, int L, int D, int C, int col_offset, int dtype, void* dS,
                             int ldds, void* stream);

/* ---- row gathers / scatters / reductions ------------------------------------------------------
 * bert_embed : out[t] = word[ids[t]] + posemb[pos[t]] + type0      (HF BertEmbeddings; LN follows separately)
 * gather_rows: dst[i] = idx[i] >= 0 ? src[idx[i]] : 0              (CLS pooling encoders.py:69, nn.Embedding
 *              model.py:37, unique-item -> slot expansion, input_embs[:, :-1] model.py:39-41)
 * scatter_add_rows: dst[idx[i]] += src[i] (fp32 dst, atomic)        (the matching backward)
 * colsum     : out[n] += sum_m x[m,n]                               (bias gradients)
 */
int morec_bert_embed_fwd(const int64_t* ids, const int32_t* pos, const float* word, const float* posemb,
                         const float* type0, void* out, int n_tok, int H, int dtype, void* stream);
int morec_bert_embed_bwd(const void* dz, const int64_t* ids, const int32_t* pos, float* dword, float* dposemb,
                         int n_tok, int H, int dtype, void* stream);
int morec_gather_rows(const void* src, const int32_t* idx, void* dst, int n, int H, int ld_src, int ld_dst,
                      int src_dtype, int dst_dtype, void* stream);
int morec_scatter_add_rows(const void* src, const int32_t* idx, float* dst, int n, int H, int ld_src, int ld_dst,
                           int src_dtype, void* stream);
int morec_colsum(const void* x, float* out, int M, int N, int ld, int dtype, void* stream);
/* Token packing plan of the text tower (encoders.py:107-117 feeds [n, 2T] rows: T word-piece ids || T attention-mask
 * entries, run.py:93-98).  mask_row_lens: lens[r] = #(mask entries != 0) of row r -- the only per-item number the host
 * needs (prefix sum -> cu_seqlens).  pack_tokens: for encoded item s (row enc_rows[s]) the kept word pieces, in column
 * order, are written to tok_ids / tok_pos [cu[s], cu[s+1]) (tok_pos = original column). */
int morec_mask_row_lens(const int64_t* text, int ld, int T, int n, int32_t* lens, void* stream);
int morec_pack_tokens(const int64_t* text, int ld, int T, const int32_t* enc_rows, const int32_t* cu, int n_enc,
                      int64_t* tok_ids, int32_t* tok_pos, void* stream);
/* out[r] = (x ? x[r] : 0) + alpha * (group_scale ? group_scale[r / rows_per_group] : 1) * y[idx ? idx[r] : r]
 * (residual add of a window-permuted branch with per-image drop-path scale: HF SwinLayer.forward; its backward;
 * broadcast backward of the mean pool).  idx[r] < 0 contributes 0. */
int morec_scale_add_rows(const void* x, const void* y, const int32_t* idx, const float* group_scale, int rows_per_group,
                         float alpha, void* out, int n, int H, int ld_y, int dtype, void* stream);
/* out[g] = mean of rows [g*rows_per_group, (g+1)*rows_per_group)   (HF SwinModel AdaptiveAvgPool1d over 49 tokens) */
int morec_mean_rows(const void* x, void* out, int n_groups, int rows_per_group, int H, int dtype, void* stream);
/* out = dy * act'(aux): mode 0 = erf-GELU with aux = pre-activation (encoders.py:70), 1 = ReLU with aux = output */
int morec_act_bwd(const void* dy, const void* aux, void* out, int64_t n, int mode, int dtype, void* stream);
/* fp32 -> 16-bit storage (dst_dtype 1 = bf16, 3 = fp16) */
int morec_cast_f32_to_16(const float* src, void* dst, int64_t n, int dst_dtype, void* stream);
/* the same for many tensors in ONE launch (the 16-bit weight shadows of a whole tower): tensors = DEVICE array of
 * n_tensors MorecCastTensor, chunk_start = DEVICE int32[n_tensors+1] exclusive prefix sum of
 * ceil(n / morec_cast_chunk_elems()); src / dst 16-byte aligned. */
typedef struct MorecCastTensor {
    const float* src;
    void* dst;
    long long n;
} MorecCastTensor;
int morec_cast_chunk_elems(void);
int morec_cast_f32_to_16_multi(const void* tensors, const int32_t* chunk_start, int n_tensors, int n_chunks,
                               int dst_dtype, void* stream);

/* ---- multi-tensor AdamW with fused unscale + found-inf (run.py:159-162, 245-247) ---------------
 * tensors: DEVICE array of n_tensors MorecAdamTensor (one per parameter); chunk_start: DEVICE int32[n_tensors+1],
 * exclusive prefix sum of ceil(n / morec_adamw_chunk_elems()) per tensor; n_chunks = chunk_start[n_tensors].
 * torch.optim.AdamW semantics (decoupled weight decay, bias correction), hyper-parameters PER TENSOR (param groups
 * may differ in lr / weight decay / betas / eps).  The step count lives on the DEVICE (`step`, float32 scalar like
 * torch's state['step']): it is advanced by one -- and the update applied -- only when no overflow was found, which
 * is GradScaler's rule (run.py:245-247) without a host round trip.  grad_scale (device scalar S, may be null):
 * every gradient is divided by S (GradScaler.unscale_); found_inf (device scalar, may be null): non-zero skips the
 * update; when check_finite != 0 found_inf (zeroed by the caller) is first set to 1 if any gradient is inf/nan.  p16 (optional) receives a 16-bit copy (p16_dtype 1 = bf16,
 * 3 = fp16) of the updated parameter: the compute-dtype weight shadow the next forward reads.
 */
typedef struct MorecAdamTensor {
    float* p;
    const float* g;
    float* m;
    float* v;
    void* p16;
    int n;
    float lr;
    float wd;
    float beta1;
    float beta2;
    float eps;
} MorecAdamTensor;
int morec_adamw_chunk_elems(void);
int morec_adamw_multi(const void* tensors, const int32_t* chunk_start, int n_tensors, int n_chunks, float* step,
                      const float* grad_scale, float* found_inf, int check_finite, int p16_dtype, void* stream);

/* ---- full-catalogue evaluation: rank of the held-out item (data_utils/metrics.py:49-57, 77-107) ------------
 * Replaces, per eval batch of U users,  scores = prec_emb . item_embeddings^T ; scores[history] = -inf ;
 * scores[1:] ; argsort ; rank of the target  by ONE tcgen05 GEMM whose epilogue counts, per user, the catalogue items
 * that precede the target:   rank(u) = 1 + count[u],
 *   count[u] = #{ c in [1, n_cols) : c not in history(u), c != tgt[u],
 *                 s[u,c] > tgt_score[u]  or  (s[u,c] == tgt_score[u] and c < tgt[u]) }      (= stable descending sort)
 * The [U, n_cols] score matrix is never written.  morec_eval_hist_bits builds the history bit matrix
 * bits[U, ceil(n_cols/32)] from a CSR list (hist_ptr int32 [U+1], hist_items int64 [n_hist]); column 0 (the pad item)
 * is always masked.  tgt_score[u] = <P[u], E[tgt[u]]> must come from the same GEMM path (morec_gemm over the gathered
 * target rows); tgt_seen (optional, [U]) receives the score this pass computed in the target column, for verification.
 * P [U, D], E [n_cols, D] row-major, dtype as morec_gemm (2 = fp32-grade 3xTF32, the reference's eval arithmetic).
 */
int morec_eval_hist_bits(const int32_t* hist_ptr, const int64_t* hist_items, int U, int n_hist, int n_cols,
                         uint32_t* bits, void* stream);
int morec_eval_rank(const void* P, const void* E, const uint32_t* hist_bits, const float* tgt_score, const int32_t* tgt,
                    int U, int n_cols, int D, int dtype, int32_t* count, float* tgt_seen, void* stream);

/* ---- BCE head of the bce_* packages (bce_text/main-end2end/model/model.py:44-51) ---------------------------
 * pos[r] = <P[r], Epos[r]>, neg[r] = <P[r], Eneg[r]>; sum_cnt = {sum over valid rows of softplus(-pos) + softplus(neg),
 * number of valid rows}; loss = sum / cnt (two BCEWithLogits means over the same rows).  Backward: dP, dEpos, dEneg
 * (same dtype as the inputs: 0 fp32, 1 bf16, 3 fp16) scaled by grad_out / cnt (device scalars). */
int morec_bce_fwd(const void* P, const void* Epos, const void* Eneg, const float* log_mask, int R, int D, int dtype,
                  float* pos, float* neg, float* sum_cnt, void* stream);
int morec_bce_bwd(const void* P, const void* Epos, const void* Eneg, const float* log_mask, const float* pos,
                  const float* neg, const float* grad_out, const float* sum_cnt, int R, int D, int dtype, void* dP,
                  void* dEpos, void* dEneg, void* stream);

/* ---- SM clock probe: out[0] = SM cycles, out[1] = nanoseconds elapsed over a ~20 us spin of one thread; lets
 * bench.py report the SM clock under load without NVML queries inside the timed region (they stall launches). */
int morec_clock_probe(uint64_t* out_cycles_ns, void* stream);

/* ---- one BERT encoder layer per call (host-side sequencing in C++: 7 launches forward, 17 backward) -----------
 * Replaces HF BertLayer.forward and its autograd backward (call site model/encoders.py:68) for PACKED tokens.
 * All pointers are caller-allocated device buffers.  Weight matrices are in the compute dtype (fp32 for dtype 0/2,
 * bf16 shadows for dtype 1); biases / LayerNorm parameters / rstd and all parameter gradients are fp32.
 *   forward : x -> qkv[n_tok,3H] -> ctx[n_tok,H] -> x1 = LN(x + drop(ctx Wo^T + bo)) -> pre/act[n_tok,I] ->
 *             x2 = LN(x1 + drop(act Wo2^T + bo2));   tmp_h[n_tok,H] is scratch.
 *   backward: (dy + dy2) = grad of x2  ->  grad of x returned as TWO addends dz1 and dxq (the layer below feeds them
 *             to its own output LayerNorm backward as dy / dy2, so the sum is never materialised).
 *             Parameter gradients are ACCUMULATED into zero-initialised fp32 buffers (split-K TMA reduce-add).
 *             scratch: dz2, dbr, dx1b, dctx [n_tok,H]; dpre [n_tok,I]; dqkv [n_tok,3H].
 */
typedef struct MorecBertLayerFwd {
    int n_tok, n_seq, H, I, n_heads, max_len, dtype, _pad;
    float eps, p_hidden, p_attn, _padf;
    uint64_t seed, off_attn, off_ln1, off_ln2;
    const int32_t* cu_seqlens;
    const void* wqkv; const float* bqkv;
    const void* w_ao; const float* b_ao; const float* g1; const float* b1;
    const void* w_i; const float* b_i;
    const void* w_o; const float* b_o; const float* g2; const float* b2;
    const void* x; void* qkv; void* ctx; void* tmp_h; void* x1; float* rstd1; void* pre; void* act; void* x2; float* rstd2;
} MorecBertLayerFwd;
int morec_bert_layer_fwd(const MorecBertLayerFwd* args, void* stream);

typedef struct MorecBertLayerBwd {
    MorecBertLayerFwd fwd;            /* the forward's parameters and saved activations */
    const void* dy; const void* dy2;  /* gradient of x2 (dy2 may be null) */
    void* dz1; void* dxq;             /* outputs: the two addends of the gradient of x */
    void* dz2; void* dbr; void* dx1b; void* dctx; void* dpre; void* dqkv;   /* scratch */
    float* dwqkv; float* dbqkv; float* dw_ao; float* db_ao; float* dg1; float* db1;
    float* dw_i; float* db_i; float* dw_o; float* db_o; float* dg2; float* db2;
} MorecBertLayerBwd;
int morec_bert_layer_bwd(const MorecBertLayerBwd* args, void* stream);
/* several layers per call: HOST arrays of records in execution order (backward: last layer first) */
int morec_bert_layers_fwd(const MorecBertLayerFwd* layers, int n_layers, void* stream);
int morec_bert_layers_bwd(const MorecBertLayerBwd* layers, int n_layers, void* stream);
/* The same, with control over the final join: join = 0 leaves the weight-gradient work of the LAST record in flight on
 * the library's side stream when the call returns.  `stream` is ordered after that work by the time the NEXT record
 * (of a following call) has been enqueued, or by a following call with join = 1 (n_layers may be 0: join only).  Lets
 * a multi-GPU caller interleave per-layer gradient collectives with the layers without serialising the two streams
 * at every layer (hot path of inbatch_sasrec_e2e_text/run.py:148 + model/encoders.py:68 under DistributedDataParallel). */
int morec_bert_layers_bwd_ex(const MorecBertLayerBwd* layers, int n_layers, int join, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOREC_B200_H */
