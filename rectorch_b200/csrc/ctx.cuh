// ctx.cuh -- the opaque engine context (workspaces + bound buffers).
#pragma once
#include "common.cuh"
#include <vector>

namespace b200 {

struct CsrSlot {
    const int64_t* indptr = nullptr;
    const int32_t* indices = nullptr;
    const float*   values = nullptr;
    int64_t n_rows = 0;
    // internal batch CSR (dense_to_csr / host staging); used when row_ids == NULL
    int64_t* int_indptr = nullptr;   // [max_batch+1]
    int32_t* int_indices = nullptr;  // [max_batch_nnz]
    float*   int_values = nullptr;   // [max_batch_nnz]
    bool     int_has_values = false;
    int64_t* bp = nullptr;           // [max_batch+1] compact scan for the current batch
    int32_t* sp = nullptr;           // [max_batch+1] segment pointer for the current batch
};

// Makes the context's device current for the duration of an API call and restores the caller's afterwards (a process
// that drives several GPUs must not find its current device changed behind its back).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (dev < 0) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct Ctx {
    b200vae_config cfg;
    int num_sms = 0;
    int64_t launches = 0;
    int n_items = 0, latent = 0;      // decoder outputs (items), latent width
    int enc_in = 0;                   // encoder input width = n_items + cond_dim
    std::vector<Layer> enc, dec;      // forward order
    int n_tensors = 0;                // 2 * (n_enc + n_dec)
    int max_width = 0;                // widest hidden activation

    // arenas
    float *w = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;
    int64_t n_elems = 0;
    bool params_bound = false;

    CsrSlot slot[2];

    // workspaces (device)
    float* xt = nullptr;              // [max_batch_nnz] scaled+dropped input values
    float* T = nullptr;               // [B] target row sums
    float* loss_row = nullptr;        // [B]
    float* kl_row = nullptr;          // [B]
    float* lse = nullptr;             // [B]
    float* rowscale = nullptr;        // [B] T_u / B_global
    bool norms_valid = false;         // c->norms hold the per-tensor norms of the CURRENT weights
    bool dw1_clean = false;           // encoder-0 gradient is all zero (the Adam kernel re-zeroes what it consumed)
    std::vector<float*> act_enc;      // per encoder layer output [B x out]
    std::vector<float*> act_dec;      // per decoder layer output, except the last
    float* z = nullptr;               // [B x latent]
    float* eps = nullptr;             // [B x latent]
    float* gvec = nullptr;            // [B x H_last]   sum_j t_uj W_d[j,:]
    float* P = nullptr;               // SIMT path only: [B x n_items] softmax * T/B
    // tensor-core path: fp16 operand images (round to nearest; 10-bit mantissa like tf32, half the bytes)
    __half* P16 = nullptr;            // [n_items x Bpad] P~^T = (softmax - t/T) * 2^14, users contiguous
    __half* hsT = nullptr;            // [(H_last+8) x Bpad] (h * T_u/(B*R) * 2^8)^T + the row of T_u/(B*R) * 2^8
    __half* h16 = nullptr;            // [B x H_last] last hidden activation
    __half* wd16 = nullptr;           // [n_items x H_last] W_d, maintained by Adam
    // data parallelism with a sharded optimizer for encoder layer 0 (b200vae_set_w1_sharding): this rank owns the item
    // rows j with j % w1_mod_n == w1_mod_r; w1g is the caller-owned gathered fp16 image [w1_mod_n][w1_rows_per x H1] that
    // the forward gather reads (row j at block j % n, index j / n) and the ranks all-gather after every Adam step
    __half* w1g = nullptr;
    int w1_mod_n = 1, w1_mod_r = 0;
    int64_t w1_rows_per = 0;
    bool wd16_external = false;       // wd16 is a caller-owned buffer (b200vae_bind_shadow)
    cudaEvent_t wd16_pending = nullptr;   // the next reader of wd16 waits for this event first (b200vae_defer_wait_event)
    float* dw_scale = nullptr;        // [1] R = power of two >= max_u T_u/B of the current batch (un-scales dW_d)
    // hidden layers on the tensor cores (tc_hidden): fp16 images of their weights (arena range [ws_lo, ws_hi),
    // maintained by Adam like wd16) and of every activation / activation gradient they read.  Activation images have
    // pitch round_up(width + 1, 8) and a column of ones at index `width` (set once): used as the B operand of the
    // weight-gradient GEMM it makes the bias gradient one more output column (no separate column-sum kernel).
    bool tc_hidden = false;
    __half* ws16 = nullptr;
    int64_t ws_lo = 0, ws_hi = 0;
    std::vector<__half*> act_enc16, act_dec16;
    std::vector<int> ld_enc16, ld_dec16;
    __half* z16 = nullptr;
    int ldz16 = 0;
    __half* dbuf16[2] = {nullptr, nullptr};   // fp16 images of the ping-pong activation gradients, pitch ld_d16
    int ld_d16 = 0;
    int32_t* d_specs = nullptr;       // [128] metric specs for topk
    // deterministic mode (b200vae_set_deterministic): sparse gather / scatter without floating-point atomics
    bool deterministic = false;
    int det_rows = 0;                 // max(encoder-0 input width, n_items)
    int* det_count = nullptr;         // [det_rows + 1] per-item non-zero counts of the batch
    int* det_off = nullptr;           // [det_rows + 1] their exclusive scan
    int* det_cursor = nullptr;        // [det_rows + 1]
    int64_t* det_ent = nullptr;       // [max_batch_nnz] (row << 32 | position) per item, fill order
    int64_t* det_sorted = nullptr;    // [max_batch_nnz] the same, ascending per item
    float* spmm_acc = nullptr;        // [B x max(width)] zeroed accumulator for multi-segment gathers
    int*   spmm_ticket = nullptr;     // [B + 1] zeroed per-row completion tickets; [B] = the loss-closing ticket of the fix-up
    float* dbuf[2] = {nullptr, nullptr};  // [B x max_width] ping-pong activation gradients
    float* part_max = nullptr;        // [n_tiles x B]
    float* part_sum = nullptr;
    int    n_lse_tiles = 0;
    float* splitk = nullptr;          // split-K partial products for dh
    int64_t splitk_elems = 0;
    float* norms = nullptr;           // [n_tensors]
    float* norm_partial = nullptr;    // [n_tensors x 64]
    int64_t* d_toff = nullptr;        // [n_tensors] tensor offsets (device copy)
    int64_t* d_tlen = nullptr;
    std::vector<int64_t> toff, tlen;  // host copies, parameters() order
    float* loss_dev = nullptr;        // [4] scratch for train_step_host
    int*   d_err = nullptr;           // device error flag (capacity overflow)
    int64_t* lens_tmp = nullptr;      // [max_batch+1]
    int64_t* lens_tmp2 = nullptr;     // [max_batch+1]

    // timing instrumentation
    std::vector<cudaEvent_t> tev;     // one event after every launch of the current step (timing mode)
    std::vector<const char*> tnames;
    int tcount = 0;
    int timing = 0;                   // 1: serial schedule, interval per launch; 2: production schedule, completion timeline
    std::vector<cudaStream_t> tstream;
    cudaEvent_t ev[5][2];
    bool ev_valid[5] = {false, false, false, false, false};

    bool use_tc = true;
    bool tc_dec = false;              // decoder-last shapes are tcgen05-eligible
    cudaEvent_t ev_wd = nullptr;      // recorded when dW_d / db_d (tail of the gradient arena) are final

    // Single-GPU fused step: Adam work that does not depend on the end of the backward pass runs on a second
    // stream, in narrow launches that share the SMs with the small kernels of the main stream (engine.cu,
    // train_step_fused).  `mark[j]` = last step whose batch read / wrote row j of the encoder-0 weight.
    int32_t* mark = nullptr;          // [n_items]
    cudaStream_t side = nullptr;
    cudaEvent_t ev_mark = nullptr;    // marks of the current step are written
    cudaEvent_t ev_side = nullptr;    // side-stream Adam of the current step is complete
    int overlap = 1;                  // bit 0: decoder-output Adam beside the encoder backward (default);
                                      // bit 1: untouched encoder-0 rows beside the forward pass -- measured slower
                                      // (the tcgen05 kernels wait for the narrow launch), kept for experiments;
                                      // bit 2 (with bit 1): those rows after the decoder-output Adam instead -- no gain
                                      // (B200VAE_OVERLAP; profiles/r2_schedule_experiments.txt)
    // scratch of b200vae_enc0_grad (a GLOBAL batch: rows of every data-parallel rank), grown on demand
    int64_t* gl_bp = nullptr;
    int32_t* gl_sp = nullptr;
    float*   gl_xt = nullptr;
    int64_t  gl_rows = 0, gl_nnz = 0;
    bool predict_transposed = true;   // score GEMM computed as W_d h^T with coalesced stores (B200VAE_PREDICT_T=0: h W_d^T)
    bool fuse_small = true;           // offset scans inside batch_prep, loss closed by the fix-up kernel (B200VAE_FUSE_SMALL=0: own launches)
    int wd_chunks = 0;                // > 1: dW_d GEMM + decoder-output Adam in item chunks on the side stream (B200VAE_WD_CHUNKS)
    int wd_chunk_ctas = 8;            // CTAs per SM of the chunked Adam launches (B200VAE_WD_CHUNK_CTAS)
    int wd_discard = 1;               // discard the consumed gradient lines from L2 (B200VAE_WD_DISCARD)
    int overlap_host = 1;             // schedule of the host-synchronous entry point (B200VAE_HOST_OVERLAP)
    int side_ctas[2] = {2, 2};        // CTAs per SM of the two side launches (B200VAE_SIDE_CTAS="a,b")
    int side_threads = 256;
};

// bookkeeping after every kernel launch: launch counter + (timing mode) an event named after the launcher
inline void note(Ctx* c, const char* name, cudaStream_t s) {
    c->launches++;
    if (c->timing) {
        if (c->tcount == (int)c->tev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            c->tev.push_back(e);
            c->tnames.push_back(name);
            c->tstream.push_back(s);
        }
        c->tnames[c->tcount] = name;
        c->tstream[c->tcount] = s;
        cudaEventRecord(c->tev[c->tcount], s);
        c->tcount++;
    }
}

inline void tick(Ctx* c, int which, int edge, cudaStream_t s) {
    if (c->timing) {
        cudaEventRecord(c->ev[which][edge], s);
        if (edge == 1) c->ev_valid[which] = true;
    }
}

}  // namespace b200
