/*
 * b200vae.h -- C ABI of libb200vae.so, the sm_100a engine underneath
 * rectorch_b200.{nets,models,samplers,evaluation,metrics}.
 *
 * The reference (makgyver/rectorch) is pure Python and has NO FFI of its own
 * (SURVEY.md section 8b): the drop-in boundary is its Python class surface, which
 * rectorch_b200 mirrors.  This header is the layer directly beneath that
 * surface; every entry point names the reference code it replaces (file:line
 * under /root/reference/rectorch) so a maintainer can see what a ctypes stub in
 * rectorch itself would bind (INTEGRATION.md shows that stub).
 *
 * Conventions
 *   - plain C types only; all `*_dev` / unqualified data pointers are DEVICE pointers
 *     unless the name ends in `_host`.
 *   - the caller owns every buffer it passes in (torch CUDA tensors in practice);
 *     the library owns only the opaque context (workspaces, TMA descriptors).
 *   - every call returns 0 on success or a negative B200VAE_E* code;
 *     b200vae_last_error() returns a thread-local message.  There is no CPU
 *     fallback anywhere: without a Blackwell GPU the calls fail.
 *   - calls are asynchronous on the `stream` argument (a cudaStream_t passed as
 *     void*; NULL = legacy default stream) and are not thread-safe per context.
 */
#ifndef B200VAE_H
#define B200VAE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VAE_MAX_LAYERS 8

#define B200VAE_OK            0
#define B200VAE_EINVAL       -1   /* bad argument / unsupported shape            */
#define B200VAE_ECUDA        -2   /* CUDA runtime / driver error                 */
#define B200VAE_ESTATE       -3   /* call sequence error (params/CSR not bound)  */
#define B200VAE_ECAPACITY    -4   /* batch or nnz exceeds the context capacity   */

typedef struct b200vae_ctx b200vae_ctx;

/* Network description == constructor arguments of
 * MultiVAE_net(dec_dims, enc_dims, dropout) / MultiDAE_net(...)  (nets.py:208, 390).
 * dec_dims[n_dec] == n_items, enc_dims[0] == n_items + cond_dim.  For is_vae the LAST encoder layer
 * has 2*enc_dims[n_enc] outputs (nets.py:264). */
typedef struct {
    int32_t device;                          /* CUDA ordinal                                  */
    int32_t is_vae;                          /* 1 = MultiVAE_net, 0 = MultiDAE_net            */
    int32_t n_enc;                           /* number of encoder Linear layers               */
    int32_t n_dec;                           /* number of decoder Linear layers               */
    int32_t enc_dims[B200VAE_MAX_LAYERS + 1];
    int32_t dec_dims[B200VAE_MAX_LAYERS + 1];
    int32_t max_batch;                       /* rows per call, capacity                       */
    int64_t max_batch_nnz;                   /* non-zeros per batch, capacity (input+target)  */
    int32_t use_tensor_cores;                /* 1 = tcgen05 path for the item-sized GEMMs when
                                                shapes allow (last hidden width % 8 == 0, n_items >= 1024;
                                                default), 0 = fp32 SIMT kernels */
    int32_t cond_dim;                        /* CMultiVAE_net(cond_dim, ...) (nets.py:455-480): the encoder input is
                                                [n_items ratings | cond_dim condition flags], enc_dims[0] ==
                                                n_items + cond_dim; the condition columns bypass F.normalize and
                                                nn.Dropout.  0 for MultiVAE_net / MultiDAE_net */
} b200vae_config;

const char* b200vae_last_error(void);
int  b200vae_version(void);

int  b200vae_ctx_create(b200vae_ctx** out, const b200vae_config* cfg);
int  b200vae_ctx_destroy(b200vae_ctx* ctx);

/* Parameter / gradient / Adam-state arenas: four fp32 arrays of n_elems floats with
 * the same internal layout.  Layer l (encoder layers first, then decoder) has its
 * weight at w_off[l] and bias at b_off[l].  Weight layout is nn.Linear's (out,in)
 * row-major EXCEPT encoder layer 0, which is stored item-major (in,out) = (n_items,H1)
 * so that a user's history is a gather of contiguous rows; the Python side exposes it
 * as a transposed nn.Parameter view (state_dict / checkpoints keep reference shapes,
 * models.py:485-488).  Replaces: nn.Linear storage + torch.optim.Adam state. */
int  b200vae_bind_params(b200vae_ctx* ctx, float* w, float* g, float* m, float* v,
                         int64_t n_elems, const int64_t* w_off, const int64_t* b_off);

/* Refresh the copies the engine derives from the weight arena (the fp16 image of
 * the decoder output weight that the tensor cores read).  b200vae_adam_step keeps them in
 * step; call this after anything else wrote the arena (init_weights, load_state_dict,
 * nets.py:235-247 / models.py:513). */
int  b200vae_sync_weights(b200vae_ctx* ctx, void* stream);

/* Data parallelism with a sharded optimizer: the fp16 image of W_d lives in a CALLER-owned
 * device buffer of n_halfs >= n_items * H halfs (a torch tensor, so that torch.distributed can all-gather the
 * ranks' shards into it).  The context re-derives the image into it at once.
 * Replaces: nothing in the reference (single process); this is the storage of nets.py:417's weight as the tensor
 * cores read it. */
int  b200vae_bind_shadow(b200vae_ctx* ctx, void* wd16, int64_t n_halfs);

/* Data parallelism with a sharded optimizer for ENCODER layer 0 (the other item-sized tensor): this rank owns the
 * item rows j with j % mod_n == mod_r of the [n_items x H1] weight.  With sharding on,
 *   - b200vae_enc0_grad scatters only into the rank's own rows (1/mod_n of the global batch's scatter work),
 *   - the Adam launches update only those rows and also write their fp16 image, packed, into block mod_r of
 *     w1_gathered[mod_n][n_items/mod_n x H1] (a caller-owned device buffer of fp16) -- the caller then all-gathers
 *     that buffer in place over the ranks (half the bytes of the fp32 rows),
 *   - the forward pass gathers encoder-0 rows from w1_gathered (row j at block j % mod_n, index j / mod_n): under
 *     this mode the encoder's first layer reads fp16 images of its weights like the decoder's last layer does.
 * The call (re)builds w1_gathered from the complete weight arena.  w1_gathered == NULL or mod_n <= 1 turns it off.
 * Replaces: nothing in the reference (single process); cf. torch.optim.Adam over nn.Linear(n_items, H1), models.py:768. */
int  b200vae_set_w1_sharding(b200vae_ctx* ctx, void* w1_gathered, int32_t mod_n, int32_t mod_r);
/* Move the rank's rows of the encoder-0 tensor between an arena (w, m or v: pass the arena's base pointer) and a
 * packed buffer: direction 0 packs the OWN rows into packed[n_items/mod_n x H1]; direction 1 rewrites EVERY row of
 * the tensor from an all-gathered packed_all[mod_n][n_items/mod_n x H1].  Used to rebuild complete fp32 tensors
 * (state_dict / checkpoints) from the shards. */
int  b200vae_w1_rows(b200vae_ctx* ctx, float* arena_base, float* packed, int direction, void* stream);

/* Data-parallel small exchange: instead of four small collectives per step (all-gather of the encoder-0 deltas,
 * all-reduces of the hidden-layer gradients, of the b_d gradient and of the loss components) every rank all-gathers ONE
 * packed record [delta | a | b | c].  b200vae_dp_pack writes dst[n_delta ..] = [a | b | c] (the caller lets
 * b200vae_forward_backward write the deltas straight into dst[0 .. n_delta)); b200vae_dp_unpack reads the gathered
 * records recv[n_ranks][stride]: delta_all[r * n_delta + i] = record r's delta, a / b / c = the sums over the ranks
 * in rank order (bit-identical on every rank).  Replaces: the gradient averaging loss.backward() would need under
 * torch DistributedDataParallel for these tensors (models.py:832). */
int  b200vae_dp_pack(float* dst, const float* a, int64_t na, const float* b, int64_t nb, const float* c, int64_t nc,
                     void* stream);
int  b200vae_dp_unpack(const float* recv, int32_t n_ranks, int64_t stride, int64_t n_delta, float* delta_all, float* a,
                       int64_t na, float* b, int64_t nb, float* c, int64_t nc, void* stream);

/* The next engine call that reads the fp16 image of W_d (forward_backward / train_step / predict / decode) makes its
 * stream wait for `event` (a cudaEvent_t recorded by the caller on the stream that refreshes the image) right before
 * the first GEMM that needs it -- the encoder part of the step overlaps the refresh.  One-shot. */
int  b200vae_defer_wait_event(b200vae_ctx* ctx, void* event);

/* Device-side capacity-overflow flag (a batch had more non-zeros than max_batch_nnz).
 * Synchronises the device; returns B200VAE_ECAPACITY once and clears the flag. */
int  b200vae_check_error_flag(b200vae_ctx* ctx);

/* AE_net.decode(z) (nets.py:227-233, 413-417): decoder layers on a caller-provided latent
 * batch z [B x latent] -> scores [B x n_items]. */
int  b200vae_decode(b200vae_ctx* ctx, const float* z, int32_t B, float* scores, void* stream);

/* Bind a device-resident CSR user x item matrix (slot 0 = training/input matrix,
 * slot 1 = target / held-out matrix).  values may be NULL (all ones).  Replaces the
 * scipy matrices held by DataSampler (samplers.py:77-81). */
int  b200vae_bind_csr(b200vae_ctx* ctx, int slot, const int64_t* indptr, const int32_t* indices,
                      const float* values, int64_t n_rows);

/* Compress a dense fp32 [B x n_items] device batch into the context's internal batch
 * CSR for `slot` (used when a caller hands dense tensors to train_batch/predict like
 * the reference does, models.py:441, 818-822).  After this call `row_ids == NULL`
 * in the step functions means "the internal batch".  */
int  b200vae_dense_to_csr(b200vae_ctx* ctx, int slot, const float* dense, int32_t B, void* stream);

/* K1: CSR -> dense batch expander (DataSampler.__iter__, samplers.py:99-105).
 * out[B x n_items] fp32; row_ids are rows of the bound CSR in `slot`. */
int  b200vae_expand_batch(b200vae_ctx* ctx, int slot, const int32_t* row_ids, int32_t B,
                          float* out, void* stream);

/* One training step minus the optimizer: forward, loss, backward into the gradient
 * arena (MultiVAE.train_batch models.py:817-832 / AETrainer.train_batch 441-445).
 *   row_ids      rows of CSR slot 0 (and slot 1 when use_target) or NULL = internal batch
 *   B_global     divisor of the batch mean (== B unless the batch is sharded over ranks)
 *   beta         annealed KL weight (models.py:824-827); ignored for DAE
 *   lam          MultiDAE norm-regulariser weight (models.py:702-706); 0 for VAE
 *   dropout_p    nn.Dropout p (train mode)
 *   seed/step    Philox key / counter for dropout + eps draws (production RNG)
 *   keep_tape    optional uint8 per non-zero of the batch (1 = kept): parity mode, replaces Philox
 *   eps_tape     optional [B x latent] N(0,1) draws: parity mode
 *   loss_out     device float[4]: {loss, nll(BCE term), kld, reg}
 *   enc0_delta_out  NULL, or device [B x H1]: receives d(loss)/d(pre-activation of encoder layer 0) and the
 *                gradient of that layer (weight + bias) is NOT written -- a data-parallel caller gathers the
 *                deltas of all ranks and calls b200vae_enc0_grad instead of all-reducing the dense matrix
 */
int  b200vae_forward_backward(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B, int32_t B_global,
                              int use_target, float beta, float lam, float dropout_p,
                              uint64_t seed, uint64_t step, int64_t row_offset,
                              const uint8_t* keep_tape, const float* eps_tape,
                              float* loss_out, float* enc0_delta_out, void* stream);

/* Encoder-0 gradient of a GLOBAL batch from its factors: dW1[j,:] = sum_u xt[u,j] * delta[u,:], db1 = colsum(delta)
 * over the B_total rows `row_ids` of CSR slot 0 (which must hold the rows of every rank); xt is recomputed
 * from the CSR with the Philox keys (seed, step, row + row_offset, item) the owning rank's forward pass used.
 * Replaces the all-reduce of the dense [n_items x H1] gradient that loss.backward() + DistributedDataParallel
 * style training would need (models.py:832; SURVEY.md section 8f N1 "sparse-aware encoder-gradient exchange"). */
int  b200vae_enc0_grad(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B_total, const float* delta,
                       float dropout_p, uint64_t seed, uint64_t step, int64_t row_offset, void* stream);

/* Make `stream` wait until the most recent b200vae_forward_backward has finished writing the
 * gradients of the decoder output layer (the arena range starting at that layer's w_off, i.e. the
 * tail of the gradient arena).  Data-parallel callers start the all-reduce of that half (~50 % of the
 * bytes) on a side stream while the encoder half of the backward pass is still running. */
int  b200vae_wait_wd_ready(b200vae_ctx* ctx, void* stream);

/* K8: fused Adam over the whole arena (torch.optim.Adam.step as configured at
 * models.py:657-659 / 768-770), with the MultiDAE extras folded into the gradient read:
 * g += lam * w/||w||_2 (per tensor) + weight_decay * w.  `step` is the 1-based count. */
int  b200vae_adam_step(b200vae_ctx* ctx, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float lam, int64_t step, void* stream);

/* Same update restricted to the arena elements [elem_lo, elem_hi) (tensor boundaries).  Lets a data-parallel
 * caller update the half of the parameters whose gradient all-reduce has finished while the other half's
 * all-reduce is still in flight.  All ranges of one optimisation step use the same `step`; the range that
 * starts at element 0 must be issued last.  narrow != 0 launches a grid of only a few CTAs per SM, for a
 * range that runs on a second stream beside other kernels (see b200vae_adam_step_split). */
int  b200vae_adam_step_range(b200vae_ctx* ctx, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float lam, int64_t step, int64_t elem_lo, int64_t elem_hi,
                             int narrow, void* stream);

/* The Adam schedule of the fused single-GPU step, on gradients that are already in the gradient arena:
 * rows of the encoder-0 weight NOT listed in touched_items have an exactly-zero gradient and are updated by a
 * launch of their own on the context's side stream (overlap_bits & 2), the decoder-output tensors by another
 * (overlap_bits & 1), the rest on `stream`; the call joins both streams before it returns control of `stream`.
 * Every element gets the arithmetic of b200vae_adam_step bit for bit (tests/test_gpu_overlap.py).
 * Replaces: torch.optim.Adam.step as configured in models.py:657-659, 768-770. */
int  b200vae_adam_step_split(b200vae_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay,
                             float lam, int64_t step, const int32_t* touched_items, int32_t n_touched,
                             int overlap_bits, void* stream);

/* forward_backward + adam_step in one call (single-GPU fast path). */
int  b200vae_train_step(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B, int use_target,
                        float beta, float lam, float dropout_p, uint64_t seed, int64_t step,
                        const uint8_t* keep_tape, const float* eps_tape,
                        float lr, float beta1, float beta2, float eps, float weight_decay,
                        float* loss_out, void* stream);

/* End-to-end variant with HOST inputs (pinned or pageable): copies the batch CSR
 * (indptr[B+1] int64 rebased to 0, indices int32, values fp32 or NULL) host->device,
 * runs train_step, copies loss back to loss_host[4] and synchronises the stream. */
int  b200vae_train_step_host(b200vae_ctx* ctx, const int64_t* indptr_host, const int32_t* indices_host,
                             const float* values_host, int32_t B, float beta, float lam,
                             float dropout_p, uint64_t seed, int64_t step, float lr,
                             float weight_decay, float* loss_host, void* stream);

/* K9: eval-mode forward (VAE.predict models.py:619-625 / AETrainer.predict 467-473).
 * scores[B x n_items]; mu/logvar [B x latent] may be NULL.  remove_train sets the
 * scores of the input's non-zeros to -inf.  `train_mode` != 0 runs the stochastic
 * train-mode forward instead (net.forward()/encode() in training mode, nets.py:394-411). */
int  b200vae_predict(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B, int remove_train,
                     int train_mode, float dropout_p, uint64_t seed, uint64_t step,
                     float* scores, float* mu, float* logvar, void* stream);

/* K10: per-row top-K by radix select + ranking metrics against the held-out CSR
 * (Metrics.recall_at_k/ndcg_at_k/hit_at_k/mrr_at_k, metrics.py:136-285).
 *   scores   [B x n_items] device fp32
 *   gt_row_ids rows of CSR slot 1, or NULL = internal batch of slot 1
 *   kinds[i] 0=recall 1=ndcg 2=hit 3=mrr ; ks[i] = k ; out[n_metrics x B] fp32
 *   topk_idx optional [B x kmax] int32 (sorted by score desc, ties by item id asc). */
int  b200vae_topk_metrics(b200vae_ctx* ctx, const float* scores, const int32_t* gt_row_ids, int32_t B,
                          const int32_t* kinds, const int32_t* ks, int32_t n_metrics,
                          float* out, int32_t* topk_idx, void* stream);

/* ---- context-free helpers (dense-tensor API surface of the reference) ------------------ */

/* Same kernel as b200vae_topk_metrics on caller-provided device arrays: scores [B x n_items],
 * ground truth as a batch-local CSR (indptr[B+1] starting at 0).  kinds/ks are DEVICE arrays;
 * kmax = max_i min(ks[i], n_items).  Backs Metrics.compute(pred_scores, ground_truth, ...)
 * (metrics.py:31-85). */
int  b200vae_topk_metrics_csr(const float* scores, int32_t B, int32_t n_items, const int64_t* gt_indptr,
                              const int32_t* gt_indices, const float* gt_values, const int32_t* kinds_dev,
                              const int32_t* ks_dev, int32_t n_metrics, int32_t kmax, float* out,
                              int32_t* topk_idx, void* stream);

/* K1 without a context: out[B x n_items] = rows `row_ids` (NULL = 0..B-1) of a device CSR
 * (DataSampler.__iter__, samplers.py:99-105). */
int  b200vae_expand_rows_raw(const int64_t* indptr, const int32_t* indices, const float* values,
                             const int32_t* row_ids, int32_t B, int32_t n_items, float* out, void* stream);

/* dense [B x n_items] -> CSR in two phases: indices == NULL counts and scans into indptr
 * (indptr[B] = nnz), a second call with buffers of >= cap entries fills them. */
int  b200vae_dense_to_csr_raw(const float* dense, int32_t B, int32_t n_items, int64_t* lens_tmp,
                              int64_t* indptr, int32_t* indices, float* values, int64_t cap, void* stream);

/* out_rows[r] = -sum_j log_softmax(logits[r,:])_j * target[r,j]   (models.py:701, 813) */
int  b200vae_multinomial_nll_rows(const float* logits, const float* target, int32_t B, int32_t n_items,
                                  float* out_rows, void* stream);
/* out_rows[r] = -0.5 * sum_l (1 + logvar - mu^2 - exp(logvar))      (models.py:814) */
int  b200vae_kl_rows(const float* mu, const float* logvar, int32_t B, int32_t L, float* out_rows, void* stream);

/* ---- per-kernel entry points (unit parity tests, ncu captures) ---------------------- */

/* C[M x N] = A[M x K] * B^T  with tcgen05 kind::f16 (fp16 operands, fp32 accumulate in TMEM).
 * A, B are IEEE fp16 arrays.
 * a_mn_major = 0: A is [M x K] row-major (K contiguous); 1: A is given as [K x M] (M contiguous)
 * b_mn_major = 0: B is [N x K] row-major (K contiguous); 1: B is given as [K x N] (N contiguous)
 * Leading dimensions in elements; must be multiples of 8; pointers 16-byte aligned.
 * Replaces: the torch.nn.functional.linear call sites of the item-sized layers (nets.py:417, their autograd
 * backward at models.py:832), whose operands the engine keeps as fp16 images (10-bit mantissa, like tf32). */
int  b200vae_gemm_f16(b200vae_ctx* ctx, const void* A, int64_t lda, int a_mn_major,
                      const void* B, int64_t ldb, int b_mn_major,
                      float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, void* stream);

/* K4 standalone: lse[r] = log sum_j exp(h[r,:].W[j,:] + b[j]) for r < B without
 * materialising the [B x n_items] logits (F.log_softmax over the decoder output,
 * models.py:813 / nets.py:417).  h16 [B x H] and W16 [n_items x H] are row-major fp16 arrays (H % 8 == 0),
 * bias is fp32.
 * lse == NULL launches only the fused GEMM + log-sum-exp kernel (the per-tile partials stay in
 * the context's workspace): used to time that kernel back to back. */
int  b200vae_dec_fwd_lse(b200vae_ctx* ctx, const void* h16, const void* W16, const float* bias,
                         int32_t B, int32_t n_items, int32_t H, float* lse, void* stream);

/* Measurement aid (scripts/launch_probe.py): launches an EMPTY kernel with the given grid, block size, dynamic shared
 * memory and cluster size -- the launch configuration of the tcgen05 kernels -- so that their fixed launch cost can
 * be separated from their work. */
int  b200vae_probe_launch(int grid, int threads, int smem_bytes, int cluster, void* stream);

/* Deterministic mode (what torch.use_deterministic_algorithms(True) asks of the reference's torch ops): the
 * sparse encoder-0 product and its gradient (the only floating-point atomics of the step) run as ordered
 * per-row / per-item reductions, so two runs of the same steps give bit-identical weights.  Slower; off by default. */
int b200vae_set_deterministic(b200vae_ctx* ctx, int enable);

/* Introspection for bench.py: kernels launched by this context since the last reset. */
int64_t b200vae_launch_count(b200vae_ctx* ctx, int reset);

/* Time (ms) of the most recent execution of an instrumented kernel group, measured
 * with CUDA events on the launching stream; which: 0 = decoder fwd (K4), 1 = Adam (K8),
 * 2 = decoder bwd recompute (K5), 3 = dW_d GEMM, 4 = dh GEMM.  Enable with
 * b200vae_set_timing(ctx, 1); values are valid after the stream is synchronised.
 * enable = 2 keeps the production two-stream schedule and makes b200vae_timing_report give completion times since the
 * start of the step (side-stream launches are suffixed with '+') instead of per-launch intervals. */
int  b200vae_set_timing(b200vae_ctx* ctx, int enable);
/* "launcher_name milliseconds\n" for every launch of the most recent step run with timing enabled
 * (CUDA events recorded after each launch on the launching stream; sync the stream first).
 * Returns the number of events. */
int  b200vae_timing_report(b200vae_ctx* ctx, char* buf, int cap);
float b200vae_kernel_ms(b200vae_ctx* ctx, int which);

/* Conditioned batch builder (ConditionedDataSampler.__iter__, samplers.py:187-232) on the device: for example i
 * = (row ex_rows[i] of the bound CSRs, condition ex_conds[i] or -1) the internal batch of slot 0 becomes
 * [CSR-0 row | one-hot(cond)] (the condition as column n_items + cond) and the internal batch of slot 1 the
 * CSR-1 row restricted to the items that satisfy the condition (any condition when -1).  item_cond_mask[j] has
 * bit c set iff item j satisfies condition c (n_cond <= 64).  Afterwards row_ids == NULL in the step functions
 * means these batches.  Examples whose filtered target is empty must have been dropped by the caller
 * (samplers.py:227-229 drops them after the fact). */
int  b200vae_build_cond_batch(b200vae_ctx* ctx, const int32_t* ex_rows, const int32_t* ex_conds, int32_t B,
                              const uint64_t* item_cond_mask, void* stream);

/* ---- EASE closed form (EASE.train / EASE.predict, models.py:1006-1026, 1051-1054) ----------
 * All pointers are DEVICE pointers; the three calls synchronise `stream` before returning.
 *
 * G32[n_items x n_items] = X^T X for the CSR matrix X (values NULL = all ones): tcgen05 GEMMs over dense fp16 images
 * of user chunks (exact for 0/1 and small-integer ratings; counts are exact in fp32 below 2^24).
 * Replaces: X = train_data.toarray(); G = np.dot(X.T, X)                                   (models.py:1008-1010) */
int  b200vae_ease_gram(const int64_t* indptr, const int32_t* indices, const float* values, int64_t n_users,
                       int32_t n_items, float* G32, void* stream);
/* Bm[n_items x n_items] (fp32) = P / (-diag P) with diag(Bm) = 0 and P = (G + lam I)^-1, inverted in fp64 by a blocked
 * in-place Gauss-Jordan elimination (the matrix is symmetric positive definite: no pivoting).
 * Replaces: G[diag] += lam; P = np.linalg.inv(G); B = P / (-np.diag(P)); B[diag] = 0       (models.py:1011-1017) */
int  b200vae_ease_solve(const float* G32, int32_t n_items, double lam, float* Bm, void* stream);
/* out[r, :] = X[row_ids[r], :] * Bm for r < n_rows (row_ids NULL = rows 0..n_rows-1; n_rows <= 65535): the rows of
 * the reference's score matrix np.dot(X, B), computed on demand instead of stored    (models.py:1019, 1051) */
int  b200vae_ease_scores(const int64_t* indptr, const int32_t* indices, const float* values, const int32_t* row_ids,
                         int32_t n_rows, int32_t n_items, const float* Bm, float* out, void* stream);

/* ---- data ingest (host side): pre-processed rating files -> canonical CSR ------------------------------
 * Replaces pd.read_csv + scipy.sparse.csr_matrix((values, (rows, cols))) in DataReader._load_train_data /
 * _load_train_test_data (data.py:363-409).  The file has a header line "uid,iid[,<value>,...]" and one
 * rating per line; it is parsed by n_threads host threads (0 = all).  All pointers below are HOST pointers. */
typedef struct b200vae_csv b200vae_csv;
int  b200vae_csv_open(b200vae_csv** out, const char* path, char sep, int n_threads);
/* record count, column count of the header, min / max of the uid column, max of the iid column */
int  b200vae_csv_info(const b200vae_csv* csv, int64_t* n_records, int32_t* n_cols, int64_t* uid_min,
                      int64_t* uid_max, int64_t* iid_max);
/* name of the third column (the one DataReader takes the values from when cfg.topn is false, data.py:373) */
const char* b200vae_csv_value_column(const b200vae_csv* csv);
/* CSR of shape [n_rows x n_cols] with row = uid - uid_base: indices sorted inside a row, duplicate
 * (row, col) records summed (scipy's canonical form).  use_values = 0 -> every record counts 1.0 (cfg.topn).
 * indptr_host [n_rows+1]; indices_host / values_host sized for the record count; *nnz_out = entries written. */
int  b200vae_csv_to_csr(const b200vae_csv* csv, int64_t uid_base, int64_t n_rows, int32_t n_cols, int use_values,
                        int64_t* indptr_host, int32_t* indices_host, double* values_host, int64_t* nnz_out);
int  b200vae_csv_close(b200vae_csv* csv);

#ifdef __cplusplus
}
#endif
#endif /* B200VAE_H */
