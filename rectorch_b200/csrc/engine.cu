// engine.cu -- context management, step orchestration and the exported C ABI.
//
// Forward / backward structure follows the reference (file:line under
// /root/reference/rectorch): nets.py:219-233 (MultiDAE_net), 394-417 (MultiVAE_net),
// models.py:813-815 / 701-706 (losses), 441-446 / 817-835 (train_batch).  What differs is
// the decomposition of the item-sized ends so that no [B x n_items] input or logits
// tensor is ever materialised in the forward pass:
//
//   x~ . W1^T                  -> gather-sum of item-major W1 rows          (spmm_gather)
//   sum_j log_softmax(l)_j t_j -> T*lse - h.(sum_j t_j W_d[j,:]) - sum_j t_j b_j
//   dlogits = softmax*T/B - t/B: dense part through the GEMMs, sparse part by scatters
#include <stdarg.h>
#include <algorithm>
#include <cmath>
#include <new>
#include "ctx.cuh"

namespace b200 {

static thread_local char g_err[512] = "";

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200VAE_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

template <typename T>
static int dmalloc(T** p, int64_t n) {
    *p = nullptr;
    if (n <= 0) n = 1;
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), (size_t)n * sizeof(T)));
    return 0;
}

// activation image [rows x ld] fp16: zeros with a column of ones at `col` (bias-gradient column, see ctx.cuh)
__global__ void k_init_image(__half* __restrict__ p, int64_t rows, int64_t ld, int col) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * ld) p[i] = __float2half_rn((i % ld) == col ? 1.f : 0.f);
}
static int alloc_image(__half** p, int64_t rows, int width, int* ld_out) {
    const int64_t ld = round_up((int64_t)width + 1, 8);
    *p = nullptr;
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(p), (size_t)(rows * ld) * sizeof(__half)));
    k_init_image<<<(unsigned)cdiv(rows * ld, 256), 256>>>(*p, rows, ld, width);
    B200_CUDA_OK(cudaGetLastError());
    *ld_out = (int)ld;
    return 0;
}

// rank-interleaved layout of the encoder-0 weight rows: row j of [rows x H] <-> block j % n, index j / n.
// only_r >= 0: pack only the rows of that residue into a [rows_per x H] buffer (block offset dropped).
__global__ void k_w1_pack(const float* __restrict__ src, float* __restrict__ dst, int rows, int H, int n, int64_t rows_per,
                          int only_r) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
    const int h4 = H / 4;
    if (i >= (int64_t)rows * h4) return;
    const int64_t j = i / h4;
    const int q = (int)(i - j * h4);
    const int r = (int)(j % n);
    if (only_r >= 0 && r != only_r) return;
    const int64_t blk = only_r >= 0 ? 0 : (int64_t)r * rows_per;
    reinterpret_cast<float4*>(dst)[(blk + j / n) * h4 + q] = reinterpret_cast<const float4*>(src)[i];
}
// fp16 image of ALL rows in the rank-interleaved layout (what the forward gather reads under encoder-0 sharding)
__global__ void k_w1_image(const float* __restrict__ src, __half* __restrict__ dst, int rows, int H, int n, int64_t rows_per) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
    const int h4 = H / 4;
    if (i >= (int64_t)rows * h4) return;
    const int64_t j = i / h4;
    const int q = (int)(i - j * h4);
    const float4 t = reinterpret_cast<const float4*>(src)[i];
    const __half2 lo = __floats2half2_rn(f16_clamp(t.x), f16_clamp(t.y));
    const __half2 hi = __floats2half2_rn(f16_clamp(t.z), f16_clamp(t.w));
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(dst)[((int64_t)(j % n) * rows_per + j / n) * h4 + q] = pk;
}
__global__ void k_w1_unpack(const float* __restrict__ src, float* __restrict__ dst, int rows, int H, int n, int64_t rows_per) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int h4 = H / 4;
    if (i >= (int64_t)rows * h4) return;
    const int64_t j = i / h4;
    const int q = (int)(i - j * h4);
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[((int64_t)(j % n) * rows_per + j / n) * h4 + q];
}

static void free_ctx(Ctx* c) {
    auto F = [](void* p) { if (p) cudaFree(p); };
    for (int s = 0; s < 2; ++s) {
        F(c->slot[s].int_indptr); F(c->slot[s].int_indices); F(c->slot[s].int_values); F(c->slot[s].bp); F(c->slot[s].sp);
    }
    F(c->xt); F(c->T); F(c->loss_row); F(c->kl_row); F(c->lse); F(c->rowscale);
    for (float* p : c->act_enc) F(p);
    for (float* p : c->act_dec) F(p);
    F(c->z); F(c->eps); F(c->gvec); F(c->P); F(c->P16); F(c->hsT); F(c->dbuf[0]); F(c->dbuf[1]);
    F(c->part_max); F(c->part_sum); F(c->splitk); F(c->norms); F(c->norm_partial);
    F(c->d_toff); F(c->d_tlen); F(c->loss_dev); F(c->d_err); F(c->lens_tmp); F(c->lens_tmp2);
    F(c->h16); if (!c->wd16_external) F(c->wd16); F(c->dw_scale); F(c->d_specs);
    F(c->ws16); F(c->z16); F(c->dbuf16[0]); F(c->dbuf16[1]);
    for (__half* p : c->act_enc16) F(p);
    for (__half* p : c->act_dec16) F(p); F(c->spmm_acc); F(c->spmm_ticket);
    F(c->det_count); F(c->det_off); F(c->det_cursor); F(c->det_ent); F(c->det_sorted);
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 2; ++j) cudaEventDestroy(c->ev[i][j]);
    for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
    if (c->ev_wd) cudaEventDestroy(c->ev_wd);
    if (c->ev_mark) cudaEventDestroy(c->ev_mark);
    if (c->ev_side) cudaEventDestroy(c->ev_side);
    if (c->side) cudaStreamDestroy(c->side);
    F(c->mark); F(c->gl_bp); F(c->gl_sp); F(c->gl_xt);
    delete c;
}

// defer_scan: the caller's next launch on this view is batch_prep(..., with_scan = true), which writes bp / sp itself
static int make_view(Ctx* c, int slot, const int32_t* row_ids, int B, BatchView* v, cudaStream_t s, bool defer_scan = false) {
    CsrSlot& S = c->slot[slot];
    v->B = B;
    v->row_ids = row_ids;
    if (row_ids) {
        B200_REQUIRE(S.indptr != nullptr, B200VAE_ESTATE, "CSR slot %d is not bound", slot);
        v->indptr = S.indptr;
        v->indices = S.indices;
        v->values = S.values;
    } else {
        v->indptr = S.int_indptr;
        v->indices = S.int_indices;
        v->values = S.int_has_values ? S.int_values : nullptr;
    }
    v->bp = S.bp;
    v->sp = S.sp;
    if (!defer_scan) B200_CHECK(launch_batch_scan(c, v->indptr, row_ids, B, c->cfg.max_batch_nnz, S.bp, S.sp, s));
    return 0;
}

// C = act(A * W^T + b) for an nn.Linear weight W [out x in] (row-major)
static int linear_fwd(Ctx* c, const float* A, int B, const Layer& L, float* out, cudaStream_t s) {
    GemmEpi e;
    e.bias = c->w + L.b_off;
    e.act = L.tanh_act ? 1 : 0;
    return launch_simt_gemm(c, EPI_STORE, A, L.in, 1, c->w + L.w_off, 1, L.in, out, L.out, B, L.out, L.in, e, s);
}

// gradients of an nn.Linear (not encoder layer 0): dW = dY^T inp, db = colsum(dY),
// dinp = (dY W) * tanh'(prev_out) if prev_out != NULL
static int linear_bwd(Ctx* c, const float* dY, const float* inp, int B, const Layer& L, float* dinp,
                      const float* prev_out, cudaStream_t s) {
    GemmEpi e;
    B200_CHECK(launch_simt_gemm(c, EPI_STORE, dY, 1, L.out, inp, L.in, 1, c->g + L.w_off, L.in, L.out, L.in, B, e, s));
    B200_CHECK(launch_colsum(c, dY, L.out, B, L.out, c->g + L.b_off, s));
    if (dinp) {
        GemmEpi e2;
        e2.mulY = prev_out;
        e2.ldy = L.in;
        B200_CHECK(launch_simt_gemm(c, EPI_STORE, dY, L.out, 1, c->w + L.w_off, L.in, 1, dinp, L.in, B, L.in, L.out, e2, s));
    }
    return 0;
}

// ---- the same two, on the tensor cores (tc_hidden) ----------------------------------------------------------
// out = act(A W^T + b): A16 [B x in] (pitch lda16), W16 = fp16 image of W [out x in]; writes fp32 `out` [B x out] and,
// when out16 != NULL, its fp16 image (pitch ld16) for the next GEMM.            (nets.py:398-404, 413-416)
static int linear_fwd_tc(Ctx* c, const __half* A16, int64_t lda16, int B, const Layer& L, float* out, __half* out16,
                         int64_t ld16, cudaStream_t s) {
    TcEpi e;
    e.bias = c->w + L.b_off;
    e.act = L.tanh_act ? 1 : 0;
    e.C16 = out16;
    e.ldc16 = ld16;
    if (!e.act && !out16) e.act = 0;
    return launch_tc_gemm(c, TC_EPI_STORE, A16, lda16, 0, c->ws16 + (L.w_off - c->ws_lo), L.in, 0, out, L.out, B, L.out, L.in, e, s);
}
// dW | db = dY^T [inp | 1]  (one GEMM: the ones column of the activation image yields the bias gradient) and
// dinp = (dY W) * tanh'(prev_out), fp32 + fp16 image.        (autograd of the same layers, models.py:832)
static int linear_bwd_tc(Ctx* c, const __half* dY16, const __half* inp16, int64_t ld_inp16, int B, const Layer& L,
                         float* dinp, __half* dinp16, const float* prev_out, cudaStream_t s) {
    TcEpi e;
    e.bias_col = L.in;
    e.bias_grad = c->g + L.b_off;
    B200_CHECK(launch_tc_gemm(c, TC_EPI_STORE, dY16, c->ld_d16, 1, inp16, ld_inp16, 1, c->g + L.w_off, L.in, L.out, L.in + 1, B,
                              e, s));
    if (dinp) {
        TcEpi e2;
        e2.mulY = prev_out;
        e2.ldy = L.in;
        e2.C16 = dinp16;
        e2.ldc16 = c->ld_d16;
        B200_CHECK(launch_tc_gemm(c, TC_EPI_STORE, dY16, c->ld_d16, 0, c->ws16 + (L.w_off - c->ws_lo), L.in, 1, dinp, L.in, B,
                                  L.in, L.out, e2, s));
    }
    return 0;
}

// a data-parallel caller refreshes the fp16 image of W_d on another stream (all-gather of the ranks' shards):
// whoever reads it next waits for that work first
static int wait_wd16(Ctx* c, cudaStream_t s) {
    if (c->wd16_pending) {
        B200_CUDA_OK(cudaStreamWaitEvent(s, c->wd16_pending, 0));
        c->wd16_pending = nullptr;
    }
    return 0;
}

struct FwdState {
    BatchView in, tgt;
    const float* h_last;      // input of the last decoder layer [B x H]
    const float* h_last_tanh; // same pointer if that activation came out of a tanh, else NULL
    int H;
    bool scan_deferred = false;   // `in` was made with defer_scan: batch_prep computes bp / sp
};

// encoder + reparameterisation + hidden decoder layers.  Leaves z and h_last in the ctx.
struct AdamHyper {
    float lr, beta1, beta2, eps, wd, lam;
    int64_t step;
    int ov = 0;           // effective overlap bits of the fused step this update belongs to (0 = serial)
};
static int adam_step(Ctx* c, const AdamHyper& h, cudaStream_t s, int64_t r_lo = 0, int64_t r_hi = -1,
                     int w1_filter = ADAM_ROWS_ALL, int ctas_per_sm = 8, bool stream_hints = false);

// `fused` (single-GPU fused step only): batch_prep stamps the encoder-0 rows this step touches, and the Adam
// update of all the OTHER rows -- zero gradient, not read by this step's gather -- starts right away on the side
// stream, in a narrow launch that shares the SMs with the forward pass.
static int forward_hidden(Ctx* c, FwdState* st, int B, bool train, float p, uint64_t seed, uint64_t step,
                          int64_t row_offset, const uint8_t* keep_tape, const float* eps_tape,
                          float* row_sum_out, cudaStream_t s, const AdamHyper* fused = nullptr) {
    const bool split_rows = fused && (fused->ov & 2);
    B200_CHECK(launch_batch_prep(c, st->in, p, seed, step, row_offset, keep_tape, train, c->xt, row_sum_out,
                                 split_rows ? c->mark : nullptr, split_rows ? (int32_t)fused->step : 0, s, st->scan_deferred));
    const Layer& e0 = c->enc[0];
    if (split_rows && !(fused->ov & 4)) {
        B200_CUDA_OK(cudaEventRecord(c->ev_mark, s));
        B200_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_mark, 0));
        B200_CHECK(adam_step(c, *fused, c->side, e0.w_off, e0.w_off + (int64_t)e0.in * e0.out, ADAM_ROWS_UNMARKED,
                             c->side_ctas[1]));
    }
    const bool tch = c->tc_hidden;
    const size_t n_enc = c->enc.size();
    // the last encoder output of a VAE is (mu | logvar): it feeds the reparameterisation, not a GEMM
    auto enc_img = [&](size_t i) -> __half* { return (tch && !(c->cfg.is_vae && i + 1 == n_enc)) ? c->act_enc16[i] : nullptr; };
    const bool w1s = c->w1_mod_n > 1;     // sharded encoder-0 optimizer: the gather reads the all-gathered copy
    B200_CHECK(launch_spmm_gather(c, st->in, c->xt, c->w + e0.w_off, e0.out, c->w + e0.b_off,
                                  e0.tanh_act ? 1 : 0, c->act_enc[0], s, enc_img(0), tch ? c->ld_enc16[0] : 0,
                                  w1s ? c->w1g : nullptr, c->w1_mod_n, c->w1_rows_per));
    for (size_t i = 1; i < n_enc; ++i) {
        if (tch) B200_CHECK(linear_fwd_tc(c, c->act_enc16[i - 1], c->ld_enc16[i - 1], B, c->enc[i], c->act_enc[i], enc_img(i),
                                          c->ld_enc16[i], s));
        else     B200_CHECK(linear_fwd(c, c->act_enc[i - 1], B, c->enc[i], c->act_enc[i], s));
    }
    const float* z;
    const float* z_tanh = nullptr;
    const __half* z16 = nullptr;
    int64_t ldz16 = 0;
    if (c->cfg.is_vae) {
        B200_CHECK(launch_reparam_kl(c, c->act_enc.back(), B, c->latent, train, eps_tape, seed, step,
                                     row_offset, st->in.row_ids, c->z, c->eps, c->kl_row, tch ? c->z16 : nullptr,
                                     c->ldz16, s));
        z = c->z;
        z16 = c->z16;
        ldz16 = c->ldz16;
    } else {
        z = c->act_enc.back();
        z_tanh = z;
        if (tch) { z16 = c->act_enc16.back(); ldz16 = c->ld_enc16.back(); }
    }
    const float* h = z;
    const float* h_tanh = z_tanh;
    const __half* h16 = z16;
    int64_t ldh16 = ldz16;
    for (size_t i = 0; i + 1 < c->dec.size(); ++i) {
        if (tch) {
            B200_CHECK(linear_fwd_tc(c, h16, ldh16, B, c->dec[i], c->act_dec[i], c->act_dec16[i], c->ld_dec16[i], s));
            h16 = c->act_dec16[i];
            ldh16 = c->ld_dec16[i];
        } else {
            B200_CHECK(linear_fwd(c, h, B, c->dec[i], c->act_dec[i], s));
        }
        h = c->act_dec[i];
        h_tanh = h;
    }
    st->h_last = h;
    st->h_last_tanh = h_tanh;
    st->H = c->dec.back().in;
    return 0;
}

// Operand prep of the decoder-output GEMMs: h [B x H] fp32 ->
//   h16 [B x H]          fp16 image (A operand of K4 / K5)
//   hsT [(H+8) x Bp]     rows 0..H-1 = (h * rs_u/R * 2^8)^T, row H = rs_u/R * 2^8 (bias-gradient row), rest 0,
//                        with rs_u = T_u / B_global and R = the power of two >= max_u rs_u, so that every entry is
//                        <= 256 in magnitude: the per-user factor of dlogits = rs_u * (softmax - t/T_u) rides on
//                        this operand of the dW_d GEMM and P~ stays a pure probability.  R goes to *dw_scale.
__global__ void __launch_bounds__(256)
k_prep_h16(const float* __restrict__ h, int B, int H, int Bp, const float* __restrict__ T, float inv_Bg,
           __half* __restrict__ h16, __half* __restrict__ hsT, float* __restrict__ dw_scale) {
    __shared__ float tile[32][33];
    __shared__ float red[8];
    pdl_sync();
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    const int tid = ty * 32 + tx;
    float R = 1.f;
    if (hsT) {
        float mx = 0.f;
        for (int i = tid; i < B; i += 256) mx = fmaxf(mx, fabsf(T[i]));
        mx = warp_max(mx);
        if (tx == 0) red[ty] = mx;
        __syncthreads();
        mx = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
        mx *= inv_Bg;
        if (mx > 0.f) {
            int e;
            frexpf(mx, &e);          // mx = f * 2^e, f in [0.5, 1)
            R = ldexpf(1.f, e);
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *dw_scale = R;
    }
    const float hs_mul = ldexpf(inv_Bg / R, (int)HS_LOG2_SCALE);
    const int b0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
    for (int i = ty; i < 32; i += 8) {
        const int b = b0 + i, hh = h0 + tx;
        float val = 0.f;
        if (b < B) {
            const float rs = hsT ? T[b] * hs_mul : 0.f;
            if (hh < H) {
                const float x = h[(int64_t)b * H + hh];
                h16[(int64_t)b * H + hh] = __float2half_rn(f16_clamp(x));
                val = x * rs;
            } else if (hh == H) {
                val = rs;
            }
        }
        tile[i][tx] = val;
    }
    if (!hsT) return;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int hh = h0 + i, b = b0 + tx;
        if (hh < H + 8 && b < Bp) hsT[(int64_t)hh * Bp + b] = __float2half_rn(f16_clamp(tile[tx][i]));
    }
}

// fused decoder GEMM + per-tile (max, sum exp) partials; the merge happens in row_loss.  With
// `for_backward` the operand prep also emits the scaled, transposed hsT the dW_d GEMM needs (T must be final).
static int dec_lse(Ctx* c, const float* h, int B, int H, int Bg, int* n_tiles, bool for_backward, cudaStream_t s) {
    const Layer& L = c->dec.back();
    const float* W = c->w + L.w_off;
    const float* b = c->w + L.b_off;
    int I = c->n_items;
    if (c->tc_dec) {
        const int Bp = (int)round_up(B, 8);
        dim3 tg((unsigned)cdiv(Bp, 32), (unsigned)cdiv(H + 8, 32));
        B200_CUDA_OK(launch_pdl(k_prep_h16, tg, dim3(32, 8), 0, s, h, B, H, Bp, (const float*)c->T, 1.0f / (float)Bg, c->h16,
                                for_backward ? c->hsT : (__half*)nullptr, c->dw_scale));
        note(c, "prep_h16", s);
        TcEpi e;
        e.bias = b;
        e.part_max = c->part_max;
        e.part_sum = c->part_sum;
        e.part_rows = c->n_lse_tiles;
        B200_CHECK(wait_wd16(c, s));
        tick(c, 0, 0, s);      // K4 proper: the tcgen05 GEMM + log-sum-exp kernel alone (operand prep is outside)
        B200_CHECK(launch_tc_gemm(c, TC_EPI_LSE, c->h16, H, 0, c->wd16, H, 0, nullptr, 0, B, I, H, e, s));
        tick(c, 0, 1, s);
        *n_tiles = tc_lse_parts(B, I, H, c->num_sms);
    } else {
        GemmEpi e;
        e.bias = b;
        e.part_max = c->part_max;
        e.part_sum = c->part_sum;
        tick(c, 0, 0, s);
        B200_CHECK(launch_simt_gemm(c, EPI_LSE, h, H, 1, W, 1, H, nullptr, 0, B, I, H, e, s));
        tick(c, 0, 1, s);
        *n_tiles = (int)cdiv(I, 64);
    }
    return 0;
}

// gradients of encoder layer 0 from its factors: dW1[j,:] += sum_u xt[u,j] * delta[u,:] (sparse scatter into the
// all-zero gradient rows), db1 = colsum(delta)
static int enc0_grad(Ctx* c, const BatchView& v, const float* xt, const float* delta, int B, cudaStream_t s) {
    const Layer& e0 = c->enc[0];
    if (!c->dw1_clean) {
        B200_CUDA_OK(cudaMemsetAsync(c->g + e0.w_off, 0, (size_t)e0.in * e0.out * sizeof(float), s));
        if (c->timing) { note(c, "memset_dW1", s); c->launches--; }
    }
    c->dw1_clean = false;
    B200_CHECK(launch_spmm_scatter(c, v, xt, 1.0f, delta, e0.out, c->g + e0.w_off, s, c->w1_mod_n, c->w1_mod_r));
    B200_CHECK(launch_colsum(c, delta, e0.out, B, e0.out, c->g + e0.b_off, s));
    return 0;
}

static int forward_backward(Ctx* c, const int32_t* row_ids, int B, int Bg, int use_target, float beta,
                            float lam, float p, uint64_t seed, uint64_t step, int64_t row_offset,
                            const uint8_t* keep_tape, const float* eps_tape, float* loss_out,
                            cudaStream_t s, const AdamHyper* fused = nullptr, float* enc0_delta_out = nullptr) {
    B200_REQUIRE(c->params_bound, B200VAE_ESTATE, "bind_params has not been called");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    B200_REQUIRE(Bg >= B, B200VAE_EINVAL, "B_global (%d) < B (%d)", Bg, B);
    const int I = c->n_items;
    const float inv_Bg = 1.0f / (float)Bg;
    if (c->timing) { c->tcount = 0; note(c, "step_start", s); c->launches--; }
    const bool dae_reg = (!c->cfg.is_vae) && lam != 0.f;
    if (dae_reg) {
        // before anything of this step can move a weight (the side stream updates untouched rows early)
        B200_CHECK(launch_tensor_norms(c, c->w, c->d_toff, c->d_tlen, c->n_tensors, c->norm_partial, c->norms, s));
        c->norms_valid = true;      // the weights do not change before this step's Adam: it reuses them
    }
    FwdState st;
    // batch_prep folds the offset scans in (every CTA re-sums the rows before its own: fine up to a few thousand rows)
    st.scan_deferred = (B <= 4096) && c->fuse_small;
    B200_CHECK(make_view(c, 0, row_ids, B, &st.in, s, st.scan_deferred));
    if (use_target) {
        B200_CHECK(make_view(c, 1, row_ids, B, &st.tgt, s));
    } else {
        st.tgt = st.in;
    }
    // ---------------- forward ----------------
    if (use_target) B200_CHECK(launch_row_sums(c, st.tgt, c->T, s));
    B200_CHECK(forward_hidden(c, &st, B, true, p, seed, step, row_offset, keep_tape, eps_tape,
                              use_target ? nullptr : c->T, s, fused));
    const Layer& DL = c->dec.back();
    const int H = st.H;
    const float* Wd = c->w + DL.w_off;
    int n_lse_tiles = 0;
    B200_CHECK(dec_lse(c, st.h_last, B, H, Bg, &n_lse_tiles, true, s));
    if (c->tc_dec) {
        // merge the LSE partials, T/B; the sparse loss term is taken from P^T after the recompute kernel (target_fixup)
        B200_CHECK(launch_row_loss(c, st.tgt, st.h_last, nullptr, H, c->w + DL.b_off, c->part_max, c->part_sum, n_lse_tiles,
                                   c->lse, c->T, inv_Bg, c->loss_row, c->rowscale, s));
    } else {
        B200_CHECK(launch_spmm_gather(c, st.tgt, nullptr, Wd, H, nullptr, 0, c->gvec, s));
        B200_CHECK(launch_row_loss(c, st.tgt, st.h_last, c->gvec, H, c->w + DL.b_off, c->part_max, c->part_sum, n_lse_tiles,
                                   c->lse, c->T, inv_Bg, c->loss_row, c->rowscale, s));
    }

    // ---------------- backward: decoder output layer ----------------
    float* rowscale = c->rowscale;   // T_u / B_global, written by row_loss
    float* dWd = c->g + DL.w_off;
    float* dbd = c->g + DL.b_off;
    float* d0 = c->dbuf[0];
    float* d1 = c->dbuf[1];
    // Fused single-GPU step, plain Adam: the dW_d GEMM moves to the side stream and runs in item chunks, each followed by
    // the Adam update of the same rows of W_d while that piece of the gradient is still in L2, and by a discard of its
    // lines -- the 2 x 120 MB round trip of dW_d through HBM disappears and the GEMM leaves the critical path (the main
    // stream goes straight to dh and the hidden layers).
    const bool wd_chunked = fused && (fused->ov & 1) && c->tc_dec && c->wd_chunks > 1 && fused->wd == 0.f && fused->lam == 0.f &&
                            I >= 64 * c->wd_chunks && (DL.w_off % 32) == 0;
    const int Bp_ = (int)round_up(B, 8);
    auto dwd_gemm = [&](int n0, int n1, cudaStream_t st) -> int {
        TcEpi e2;
        e2.bias_grad = dbd + n0;
        e2.bias_col = H;       // row H of the product (the rs row of hsT) is the bias gradient
        e2.transpose_out = 1;
        e2.out_scale = exp2f(-(HS_LOG2_SCALE + PROB_LOG2_SCALE));
        e2.out_scale_ptr = c->dw_scale;
        return launch_tc_gemm(c, TC_EPI_STORE, c->hsT, Bp_, 0, c->P16 + (int64_t)n0 * Bp_, Bp_, 0, dWd + (int64_t)n0 * H, H, H + 8,
                              n1 - n0, B, e2, st);
    };
    if (c->tc_dec) {
        // P~^T [I x Bp] (item-major, users contiguous, fp16) = softmax * 2^14 from the recompute kernel
        const int Bp = (int)round_up(B, 8);
        TcEpi e;
        e.bias = c->w + DL.b_off;
        e.lse = c->lse;
        tick(c, 2, 0, s);
        B200_CHECK(launch_tc_gemm(c, TC_EPI_PROB, c->h16, H, 0, c->wd16, H, 0, c->P16, Bp, B, I, H, e, s));
        tick(c, 2, 1, s);
        // sparse part of dlogits + the sparse loss term, straight on P~^T
        // ... and, from its last CTA, the loss itself (k_loss_final without its launch)
        LossTail lt = {};
        if (c->fuse_small) {
            lt.loss_out = loss_out;
            lt.kl_row = c->cfg.is_vae ? c->kl_row : nullptr;
            lt.norms = dae_reg ? c->norms : nullptr;
            lt.ticket = c->spmm_ticket + c->cfg.max_batch;
            lt.B = B;
            lt.n_tensors = c->n_tensors;
            lt.inv_Bg = inv_Bg;
            lt.beta = c->cfg.is_vae ? beta : 0.f;
            lt.lam = dae_reg ? lam : 0.f;
        }
        B200_CHECK(launch_target_fixup(c, st.tgt, c->P16, Bp, c->T, c->lse, c->h16, H, c->wd16, H, c->w + DL.b_off, H,
                                       c->loss_row, s, c->fuse_small ? &lt : nullptr));
        // (dW_d | db_d)^T = [hs | rs]^T [(H+8) x B] * P~ [B x I]: A = hsT (K-major, from dec_lse; stays resident in
        // shared memory), B = P~^T (K-major).  Computing the transpose puts the hidden index on the TMEM lanes, so each
        // epilogue store instruction writes 32 consecutive floats of a dW_d row (full 128 B lines).  The epilogue
        // un-scales by R * 2^-(8+14).
        if (!wd_chunked) {
            tick(c, 3, 0, s);
            B200_CHECK(dwd_gemm(0, I, s));
            tick(c, 3, 1, s);
        }
        // dh = rs_u * 2^-14 * (P~ W_d) : A = P~^T given as [K=I x M=Bp] (MN-major), B = W_d [K=I x N=H] (MN-major);
        // the per-user factor and tanh' are applied by the split-K reduction
        TcEpi e3;
        int split = std::max(1, std::min(64, tc_parallel_tiles(c->num_sms) / tc_output_tiles(B, H)));
        e3.split_k = split;
        e3.split_stride = (int64_t)B * H;
        B200_REQUIRE((int64_t)split * B * H <= c->splitk_elems, B200VAE_ECAPACITY, "split-K workspace too small");
        tick(c, 4, 0, s);
        B200_CHECK(launch_tc_gemm(c, TC_EPI_STORE, c->P16, Bp, 1, c->wd16, H, 1, c->splitk, H, B, H, I, e3, s));
        B200_CHECK(launch_splitk_reduce(c, c->splitk, split, e3.split_stride, d0, H, B, H, H, nullptr, 0,
                                        0.f, st.h_last_tanh, H, rowscale, exp2f(-PROB_LOG2_SCALE), s,
                                        c->tc_hidden ? c->dbuf16[0] : nullptr, c->ld_d16));
        tick(c, 4, 1, s);
    } else {
        GemmEpi e;
        e.bias = c->w + DL.b_off;
        e.lse = c->lse;
        e.rowscale = rowscale;
        tick(c, 2, 0, s);
        B200_CHECK(launch_simt_gemm(c, EPI_PROB, st.h_last, H, 1, Wd, 1, H, c->P, I, B, I, H, e, s));
        tick(c, 2, 1, s);
        GemmEpi e2;
        tick(c, 3, 0, s);
        B200_CHECK(launch_simt_gemm(c, EPI_STORE, c->P, 1, I, st.h_last, H, 1, dWd, H, I, H, B, e2, s));
        B200_CHECK(launch_colsum(c, c->P, I, B, I, dbd, s));
        tick(c, 3, 1, s);
        GemmEpi e3;
        e3.addend = c->gvec;
        e3.ld_addend = H;
        e3.addend_scale = -inv_Bg;
        e3.mulY = st.h_last_tanh;
        e3.ldy = H;
        tick(c, 4, 0, s);
        B200_CHECK(launch_simt_gemm(c, EPI_STORE, c->P, I, 1, Wd, H, 1, d0, H, B, H, I, e3, s));
        tick(c, 4, 1, s);
    }
    // sparse part of dlogits = -t/Bg (the tensor-core path already folded it into P^T)
    if (!c->tc_dec) B200_CHECK(launch_spmm_scatter_bias(c, st.tgt, nullptr, -inv_Bg, st.h_last, H, dWd, dbd, s));
    if (!(c->tc_dec && c->fuse_small))
        B200_CHECK(launch_loss_final(c, c->loss_row, c->cfg.is_vae ? c->kl_row : nullptr, B, inv_Bg,
                                     c->cfg.is_vae ? beta : 0.f, dae_reg ? lam : 0.f,
                                     dae_reg ? c->norms : nullptr, c->n_tensors, loss_out, s));
    // the decoder-output gradients (the tail of the gradient arena) are final from here on: a data-parallel
    // caller can start reducing them while the rest of the backward pass runs (b200vae_wait_wd_ready)
    B200_CUDA_OK(cudaEventRecord(c->ev_wd, s));
    if (fused && (fused->ov & 1)) {
        // fused single-GPU step: hand the decoder-output Adam to the side stream NOW, before the launches of the
        // encoder backward are issued -- a caller that synchronises every step (b200vae_train_step_host) is
        // host-bound here, and a side launch issued after them would start too late to overlap anything
        // (measured: 869 us/step end to end vs 740 us when the host runs ahead)
        B200_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_wd, 0));
        if (wd_chunked) {
            const int per = (int)round_up(cdiv(I, c->wd_chunks), 32);   // 32 rows: chunk boundaries on 128-byte lines
            for (int n0 = 0; n0 < I; n0 += per) {
                const int n1 = std::min(I, n0 + per);
                B200_CHECK(dwd_gemm(n0, n1, c->side));
                const int64_t lo = DL.w_off + (int64_t)n0 * H;
                const int64_t hi = (n1 == I) ? c->n_elems : DL.w_off + (int64_t)n1 * H;
                B200_CHECK(adam_step(c, *fused, c->side, lo, hi, ADAM_ROWS_ALL, c->wd_chunk_ctas, true));
                if (c->wd_discard) B200_CHECK(launch_discard_l2(c, c->g + lo, hi - lo, c->side));
            }
        } else {
            B200_CHECK(adam_step(c, *fused, c->side, c->dec.back().w_off, c->n_elems, ADAM_ROWS_ALL, c->side_ctas[0]));
        }
    }
    if (fused && (fused->ov & 6) == 6) {
        // schedule bit 2 ("late"): the untouched encoder-0 rows follow the decoder-output Adam on the side stream, beside the
        // hidden-layer backward and the sparse scatter -- the part of the step that leaves HBM idle -- instead of beside
        // the forward pass, so the closing Adam on the main stream only has the rows this batch touched
        const Layer& E0 = c->enc[0];
        B200_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_wd, 0));
        B200_CHECK(adam_step(c, *fused, c->side, E0.w_off, E0.w_off + (int64_t)E0.in * E0.out, ADAM_ROWS_UNMARKED,
                             c->side_ctas[1]));
    }

    // ---------------- backward: hidden decoder layers ----------------
    // invariant: `cur` holds d(loss)/d(pre-activation of the layer below the one being processed)
    float* cur = d0;
    float* nxt = d1;
    const bool tch = c->tc_hidden;
    __half* cur16 = c->dbuf16[0];     // fp16 images travel with the fp32 buffers (tc_hidden)
    __half* nxt16 = c->dbuf16[1];
    const float* z_tanh = c->cfg.is_vae ? nullptr : c->act_enc.back();
    const float* z = c->cfg.is_vae ? c->z : c->act_enc.back();
    const __half* z16 = c->cfg.is_vae ? c->z16 : (tch ? c->act_enc16.back() : nullptr);
    const int64_t ldz16 = c->cfg.is_vae ? c->ldz16 : (tch ? c->ld_enc16.back() : 0);
    for (int i = (int)c->dec.size() - 2; i >= 0; --i) {
        const float* inp = (i == 0) ? z : c->act_dec[i - 1];
        const float* prev_tanh = (i == 0) ? z_tanh : c->act_dec[i - 1];
        if (tch) B200_CHECK(linear_bwd_tc(c, cur16, (i == 0) ? z16 : c->act_dec16[i - 1], (i == 0) ? ldz16 : c->ld_dec16[i - 1], B,
                                          c->dec[i], nxt, nxt16, prev_tanh, s));
        else     B200_CHECK(linear_bwd(c, cur, inp, B, c->dec[i], nxt, prev_tanh, s));
        std::swap(cur, nxt);
        std::swap(cur16, nxt16);
    }
    if (c->cfg.is_vae) {
        B200_CHECK(launch_dz_to_denc(c, cur, c->act_enc.back(), c->eps, B, c->latent, beta * inv_Bg, true, nxt,
                                     tch ? nxt16 : nullptr, c->ld_d16, s));
        std::swap(cur, nxt);
        std::swap(cur16, nxt16);
    }
    // ---------------- backward: encoder ----------------
    for (int i = (int)c->enc.size() - 1; i >= 1; --i) {
        const float* inp = c->act_enc[i - 1];
        const float* prev_tanh = c->enc[i - 1].tanh_act ? c->act_enc[i - 1] : nullptr;
        if (tch) B200_CHECK(linear_bwd_tc(c, cur16, c->act_enc16[i - 1], c->ld_enc16[i - 1], B, c->enc[i], nxt, nxt16, prev_tanh, s));
        else     B200_CHECK(linear_bwd(c, cur, inp, B, c->enc[i], nxt, prev_tanh, s));
        std::swap(cur, nxt);
        std::swap(cur16, nxt16);
    }
    const Layer& e0 = c->enc[0];
    if (enc0_delta_out) {
        // data-parallel caller: the encoder-0 gradient is assembled from the factors of ALL ranks
        // (b200vae_enc0_grad) instead of being reduced as a dense [n_items x H1] matrix
        B200_CUDA_OK(cudaMemcpyAsync(enc0_delta_out, cur, (size_t)B * e0.out * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    return enc0_grad(c, st.in, c->xt, cur, B, s);
}

// Adam over the arena range [r_lo, r_hi) (whole arena: 0, n_elems).  Ranges must not cut a tensor; the ranges
// of one step are issued tail first and the range that starts at 0 closes the step.  `w1_filter` selects rows
// of the encoder-0 weight by their step mark (AdamOpt), `ctas_per_sm` the width of the launch.
static int adam_step(Ctx* c, const AdamHyper& h, cudaStream_t s, int64_t r_lo, int64_t r_hi, int w1_filter,
                     int ctas_per_sm, bool stream_hints) {
    if (r_hi < 0) r_hi = c->n_elems;
    B200_REQUIRE(r_lo >= 0 && r_lo < r_hi && r_hi <= c->n_elems && r_lo % 4 == 0, B200VAE_EINVAL, "bad Adam range");
    B200_REQUIRE(c->params_bound, B200VAE_ESTATE, "bind_params has not been called");
    B200_REQUIRE(h.step >= 1, B200VAE_EINVAL, "adam step must be >= 1");
    double bc1 = 1.0 - std::pow((double)h.beta1, (double)h.step);
    double bc2 = 1.0 - std::pow((double)h.beta2, (double)h.step);
    float step_size = (float)((double)h.lr / bc1);
    float bc2_sqrt = (float)std::sqrt(bc2);
    tick(c, 1, 0, s);
    const Layer& DL = c->dec.back();
    __half* shadow = c->tc_dec ? c->wd16 : nullptr;
    const int64_t sh_lo = DL.w_off, sh_hi = DL.w_off + (int64_t)DL.in * DL.out;
    // Adam re-zeroes the encoder-0 gradient rows it consumed (sparse writes), so forward_backward never memsets
    const Layer& E0 = c->enc[0];
    const int64_t z_lo = E0.w_off, z_hi = E0.w_off + (int64_t)E0.in * E0.out;
    AdamOpt opt;
    opt.mark = c->mark;
    opt.mark_step = (int32_t)h.step;
    opt.filter = w1_filter;
    opt.row_len = E0.out;
    opt.ctas_per_sm = ctas_per_sm;
    opt.threads = (ctas_per_sm < 8) ? c->side_threads : 256;
    opt.stream = stream_hints;
    if (c->w1_mod_n > 1 && w1_filter == ADAM_ROWS_ALL) {      // this rank updates only its rows of the encoder-0 weight
        opt.filter = ADAM_ROWS_MOD;
        opt.w1g = c->w1g;
        opt.mod_n = c->w1_mod_n;
        opt.mod_r = c->w1_mod_r;
        opt.rows_per = c->w1_rows_per;
    }
    if (c->tc_hidden) {     // window positions are relative to each launch's base pointers
        opt.shadow2 = c->ws16;
        opt.s2_lo = c->ws_lo - r_lo;
        opt.s2_hi = c->ws_hi - r_lo;
    }
    if (h.wd == 0.f && h.lam == 0.f) {
        // kernel indices are relative to the pointers it gets: shift the shadow / re-zero windows by r_lo
        // (the shadow pointer is pre-offset so that shadow[(e - r_lo) - (sh_lo - r_lo)] addresses element e - sh_lo)
        const int64_t zl = std::max(z_lo, r_lo) - r_lo, zh = std::min(z_hi, r_hi) - r_lo;
        B200_REQUIRE(w1_filter == ADAM_ROWS_ALL || zh <= zl || z_lo >= r_lo, B200VAE_EINVAL,
                     "a filtered Adam range must start at or before the encoder-0 weight");
        B200_CHECK(launch_adam(c, c->w + r_lo, c->g + r_lo, c->m + r_lo, c->v + r_lo, r_hi - r_lo, step_size, h.beta1,
                               h.beta2, bc2_sqrt, h.eps, 0.f, 0.f, nullptr, shadow, sh_lo - r_lo, sh_hi - r_lo,
                               z_lo - r_lo, z_hi - r_lo, opt, s));
    } else {
        if (h.lam != 0.f && !c->norms_valid)
            B200_CHECK(launch_tensor_norms(c, c->w, c->d_toff, c->d_tlen, c->n_tensors, c->norm_partial, c->norms, s));
        for (int t = 0; t < c->n_tensors; ++t) {
            int64_t o = c->toff[t];
            if (o < r_lo || o >= r_hi) continue;
            const bool is_wd = (o == DL.w_off);
            if (c->tc_hidden) { opt.s2_lo = c->ws_lo - o; opt.s2_hi = c->ws_hi - o; }
            B200_CHECK(launch_adam(c, c->w + o, c->g + o, c->m + o, c->v + o, c->tlen[t], step_size, h.beta1,
                                   h.beta2, bc2_sqrt, h.eps, h.wd, h.lam, h.lam != 0.f ? c->norms + t : nullptr,
                                   is_wd ? shadow : nullptr, 0, is_wd ? c->tlen[t] : 0,
                                   0, (o == z_lo) ? c->tlen[t] : 0, opt, s));
        }
    }
    tick(c, 1, 1, s);
    // the range that starts at 0 and reads gradients is the last (or only) one of a step: the weights have moved.
    // (The untouched-row launch also starts at 0 but runs first; invalidating there would make the closing launch
    // recompute the norms from half-updated weights.)
    if (r_lo == 0 && w1_filter != ADAM_ROWS_UNMARKED) c->norms_valid = false;
    if (w1_filter != ADAM_ROWS_UNMARKED && r_lo <= z_lo && r_hi >= z_hi) c->dw1_clean = true;
    return 0;
}

// One single-GPU optimisation step.  Timeline (main stream | side stream):
//   batch scan, batch_prep (stamps the touched encoder-0 rows)  | -
//   encoder, decoder, loss, decoder-output backward (tcgen05)   | Adam of the untouched encoder-0 rows
//   hidden-layer backward, sparse scatter                       | Adam of W_d, b_d (+ fp16 image), once dW_d is final
//   Adam of the touched encoder-0 rows and the small tensors    |
// The side launches are narrow grid-stride kernels (side_ctas CTAs per SM) so the small kernels of the main stream
// run beside them; the tcgen05 kernels need whole SMs and simply start when a side launch has drained.  Every
// element sees exactly the arithmetic of the one-launch Adam (same adam_one, gradient +0 for untouched rows).
// schedule bits that can actually be used out of the requested ones (B200VAE_OVERLAP semantics)
static int effective_overlap(const Ctx* c, int requested) {
    const Layer& E0 = c->enc[0];
    int ov = (c->timing == 1 || !c->side) ? 0 : (requested & 7);
    if (!(ov & 2)) ov &= ~4;
    if (E0.out % 4 != 0 || E0.w_off % 4 != 0) ov &= ~6;      // the row filter works on whole float4s
    return ov;
}

// Adam part of the fused step (`ov` != 0).  Expects: marks of step h.step written before ev_mark (bit 1; the
// untouched-row launch was already issued by forward_hidden), ev_wd recorded once dW_d / db_d are final;
// tail_issued: forward_backward already put the decoder-output launch on the side stream.
static int fused_adam_finish(Ctx* c, const AdamHyper& h, int ov, cudaStream_t s, bool tail_issued = false) {
    const int64_t cut = c->dec.back().w_off;
    int64_t head_hi = c->n_elems;
    if (ov & 1) {
        if (!tail_issued) {
            B200_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_wd, 0));
            B200_CHECK(adam_step(c, h, c->side, cut, c->n_elems, ADAM_ROWS_ALL, c->side_ctas[0]));
        }
        head_hi = cut;
    }
    B200_CUDA_OK(cudaEventRecord(c->ev_side, c->side));
    B200_CHECK(adam_step(c, h, s, 0, head_hi, (ov & 2) ? ADAM_ROWS_MARKED : ADAM_ROWS_ALL, 8));
    B200_CUDA_OK(cudaStreamWaitEvent(s, c->ev_side, 0));
    return 0;
}

static int train_step_fused(Ctx* c, const int32_t* row_ids, int B, int use_target, float beta, float p, uint64_t seed,
                            const uint8_t* keep_tape, const float* eps_tape, AdamHyper h, float* loss_out,
                            cudaStream_t s, bool host_sync = false) {
    // host_sync: the caller synchronises with the host after every step (b200vae_train_step_host), so the GPU
    // starts each step idle.  Round 1 (tf32 kernels, no programmatic dependent launch) had schedule 3 ahead for such
    // callers; with the round-2 kernels schedule 1 wins for both kinds of caller [B200, cfg2, 60 steps]:
    // 652 us/step and 710 K users/s end to end, against 700 us / 669 K for schedule 3 and 715 us / 662 K for 7
    // (profiles/r2_schedule_experiments.txt).
    h.ov = effective_overlap(c, host_sync ? c->overlap_host : c->overlap);
    if (!h.ov) {
        B200_CHECK(forward_backward(c, row_ids, B, B, use_target, beta, h.lam, p, seed, (uint64_t)h.step, 0, keep_tape,
                                    eps_tape, loss_out, s));
        return adam_step(c, h, s);
    }
    B200_CHECK(forward_backward(c, row_ids, B, B, use_target, beta, h.lam, p, seed, (uint64_t)h.step, 0, keep_tape,
                                eps_tape, loss_out, s, &h));
    return fused_adam_finish(c, h, h.ov, s, true);
}

__global__ void k_stamp_rows(const int32_t* __restrict__ items, int n, int n_items, int32_t* __restrict__ mark, int32_t step) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && items[i] >= 0 && items[i] < n_items) mark[items[i]] = step;
}

static int predict(Ctx* c, const int32_t* row_ids, int B, int remove_train, int train_mode, float p,
                   uint64_t seed, uint64_t step, float* scores, float* mu, float* logvar, cudaStream_t s) {
    B200_REQUIRE(c->params_bound, B200VAE_ESTATE, "bind_params has not been called");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    FwdState st;
    st.scan_deferred = (B <= 4096) && c->fuse_small;
    B200_CHECK(make_view(c, 0, row_ids, B, &st.in, s, st.scan_deferred));
    st.tgt = st.in;
    B200_CHECK(forward_hidden(c, &st, B, train_mode != 0, p, seed, step, 0, nullptr, nullptr, nullptr, s));
    const int L = c->latent, I = c->n_items;
    if (c->cfg.is_vae) {
        const float* eo = c->act_enc.back();
        if (mu) B200_CUDA_OK(cudaMemcpy2DAsync(mu, L * sizeof(float), eo, 2 * L * sizeof(float), L * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
        if (logvar) B200_CUDA_OK(cudaMemcpy2DAsync(logvar, L * sizeof(float), eo + L, 2 * L * sizeof(float), L * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
    } else if (mu) {   // DAE: `mu` receives the encoder output (AE_net.encode)
        B200_CUDA_OK(cudaMemcpyAsync(mu, c->act_enc.back(), (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    if (!scores) return 0;
    const Layer& DL = c->dec.back();
    if (c->tc_dec) {
        TcEpi e;
        B200_CHECK(launch_to_f16(c, st.h_last, c->h16, B, st.H, st.H, s));
        B200_CHECK(wait_wd16(c, s));
        if (c->predict_transposed) {
            // scores^T = W_d h^T: the item index sits on the TMEM lanes, so every epilogue store instruction writes 32
            // consecutive scores of one user (a full 128-byte line) instead of 16 bytes in each of 32 users' rows
            e.transpose_out = 1;
            e.row_bias = c->w + DL.b_off;
            B200_CHECK(launch_tc_gemm(c, TC_EPI_STORE, c->wd16, st.H, 0, c->h16, st.H, 0, scores, I, I, B, st.H, e, s));
        } else {
            e.bias = c->w + DL.b_off;
            B200_CHECK(launch_tc_gemm(c, TC_EPI_STORE, c->h16, st.H, 0, c->wd16, st.H, 0, scores, I, B, I, st.H, e, s));
        }
    } else {
        B200_CHECK(linear_fwd(c, st.h_last, B, DL, scores, s));
    }
    if (remove_train) B200_CHECK(launch_mask_seen(c, st.in, I, scores, s));
    return 0;
}

}  // namespace b200

// =============================================================================================
// C ABI
// =============================================================================================
using namespace b200;

extern "C" {

const char* b200vae_last_error(void) { return g_err; }
int b200vae_version(void) { return 100; }

int b200vae_ctx_create(b200vae_ctx** out, const b200vae_config* cfg) {
    B200_REQUIRE(out && cfg, B200VAE_EINVAL, "null argument");
    *out = nullptr;
    B200_REQUIRE(cfg->n_enc >= 1 && cfg->n_enc <= B200VAE_MAX_LAYERS && cfg->n_dec >= 1 && cfg->n_dec <= B200VAE_MAX_LAYERS,
                 B200VAE_EINVAL, "layer counts out of range");
    B200_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch_nnz >= 1, B200VAE_EINVAL, "capacities must be positive");
    int ndev = 0;
    B200_CUDA_OK(cudaGetDeviceCount(&ndev));
    B200_REQUIRE(cfg->device >= 0 && cfg->device < ndev, B200VAE_EINVAL, "no such CUDA device %d", cfg->device);
    DeviceGuard dev_guard(cfg->device);
    cudaDeviceProp prop;
    B200_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    B200_REQUIRE(prop.major == 10, B200VAE_ECUDA,
                 "b200vae is built for sm_100a (Blackwell) only; device %d is sm_%d%d", cfg->device, prop.major, prop.minor);
    Ctx* c = new (std::nothrow) Ctx();
    B200_REQUIRE(c, B200VAE_ECUDA, "out of host memory");
    c->cfg = *cfg;
    c->num_sms = prop.multiProcessorCount;
    c->n_items = cfg->dec_dims[cfg->n_dec];
    c->enc_in = cfg->enc_dims[0];
    c->latent = cfg->dec_dims[0];
    c->use_tc = cfg->use_tensor_cores != 0;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 2; ++j) cudaEventCreate(&c->ev[i][j]);
    cudaEventCreateWithFlags(&c->ev_wd, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_mark, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming);
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess) c->side = nullptr;
    if (const char* e = getenv("B200VAE_OVERLAP")) c->overlap = c->overlap_host = atoi(e) & 7;
    if (const char* e = getenv("B200VAE_HOST_OVERLAP")) c->overlap_host = atoi(e) & 7;
    if (const char* e = getenv("B200VAE_FUSE_SMALL")) c->fuse_small = atoi(e) != 0;
    if (const char* e = getenv("B200VAE_PREDICT_T")) c->predict_transposed = atoi(e) != 0;
    if (const char* e = getenv("B200VAE_WD_CHUNKS")) c->wd_chunks = std::max(0, std::min(64, atoi(e)));
    if (const char* e = getenv("B200VAE_WD_CHUNK_CTAS")) c->wd_chunk_ctas = std::max(1, std::min(8, atoi(e)));
    if (const char* e = getenv("B200VAE_WD_DISCARD")) c->wd_discard = atoi(e) != 0;
    // width of the decoder-output Adam on the side stream: narrow when there are hidden-layer kernels to share
    // the SMs with (cfg2: 2 CTAs/SM = 743 us/step vs 778 at 8), full width when the backward tail is only the
    // sparse scatter (cfg3, one hidden layer: 386 us at 8 vs 392 at 2)            [B200, profiles/r1_overlap_sweep.txt]
    c->side_ctas[0] = (cfg->n_enc + cfg->n_dec > 2) ? 2 : 8;
    if (const char* e = getenv("B200VAE_SIDE_CTAS")) {
        int a = 0, b = 0;
        int n = sscanf(e, "%d,%d", &a, &b);
        if (n >= 1 && a >= 1 && a <= 8) c->side_ctas[0] = c->side_ctas[1] = a;
        if (n >= 2 && b >= 1 && b <= 8) c->side_ctas[1] = b;
    }
    if (const char* e = getenv("B200VAE_SIDE_THREADS")) {
        int t = atoi(e);
        if (t >= 32 && t <= 256) c->side_threads = t & ~31;
    }
    if (cfg->cond_dim < 0 || cfg->cond_dim > 64 || c->enc_in != c->n_items + cfg->cond_dim || cfg->enc_dims[cfg->n_enc] != c->latent) {
        set_error("enc_dims/dec_dims are inconsistent (encoder input %d vs n_items %d + cond_dim %d, latent %d vs %d)", c->enc_in,
                  c->n_items, cfg->cond_dim, cfg->enc_dims[cfg->n_enc], c->latent);
        free_ctx(c);
        return B200VAE_EINVAL;
    }
    for (int i = 0; i < cfg->n_enc; ++i) {
        Layer L;
        L.in = cfg->enc_dims[i];
        L.out = cfg->enc_dims[i + 1];
        L.tanh_act = true;
        if (cfg->is_vae && i == cfg->n_enc - 1) { L.out *= 2; L.tanh_act = false; }
        L.w_off = L.b_off = -1;
        c->enc.push_back(L);
    }
    for (int i = 0; i < cfg->n_dec; ++i) {
        Layer L;
        L.in = cfg->dec_dims[i];
        L.out = cfg->dec_dims[i + 1];
        L.tanh_act = (i != cfg->n_dec - 1);
        L.w_off = L.b_off = -1;
        c->dec.push_back(L);
    }
    c->n_tensors = 2 * (cfg->n_enc + cfg->n_dec);
    c->max_width = 1;
    for (auto& L : c->enc) c->max_width = std::max(c->max_width, L.out);
    for (size_t i = 0; i + 1 < c->dec.size(); ++i) c->max_width = std::max(c->max_width, c->dec[i].out);
    c->max_width = std::max(c->max_width, c->latent);

    const int64_t Bm = cfg->max_batch, I = c->n_items, nnz = cfg->max_batch_nnz;
    const int H = c->dec.back().in;
    const int64_t Bp = round_up(Bm, 8);
    c->tc_dec = c->use_tc && tc_supported((int)Bm, (int)I, H, H, H) && (I >= 1024);
    // hidden layers on the tensor cores: every layer between the two item-sized ones needs 16-byte aligned fp16 rows
    {
        const char* ev = getenv("B200VAE_TC_HIDDEN");
        bool ok = c->tc_dec && !(ev && ev[0] == '0') && (c->enc.size() + c->dec.size() > 2);
        for (size_t i = 1; i < c->enc.size(); ++i) ok = ok && c->enc[i].in % 8 == 0 && c->enc[i].out % 8 == 0;
        for (size_t i = 0; i + 1 < c->dec.size(); ++i) ok = ok && c->dec[i].in % 8 == 0 && c->dec[i].out % 8 == 0;
        c->tc_hidden = ok;
    }
    int rc = 0;
#define A_(expr) do { if (!rc) rc = (expr); } while (0)
    for (int s = 0; s < 2; ++s) {
        A_(dmalloc(&c->slot[s].int_indptr, Bm + 1));
        A_(dmalloc(&c->slot[s].int_indices, nnz));
        A_(dmalloc(&c->slot[s].int_values, nnz));
        A_(dmalloc(&c->slot[s].bp, Bm + 1));
        A_(dmalloc(&c->slot[s].sp, Bm + 1));
    }
    A_(dmalloc(&c->xt, nnz));
    A_(dmalloc(&c->T, Bm)); A_(dmalloc(&c->loss_row, Bm)); A_(dmalloc(&c->kl_row, Bm)); A_(dmalloc(&c->lse, Bm)); A_(dmalloc(&c->rowscale, Bm));
    for (auto& L : c->enc) { float* p = nullptr; A_(dmalloc(&p, Bm * L.out)); c->act_enc.push_back(p); }
    for (size_t i = 0; i + 1 < c->dec.size(); ++i) { float* p = nullptr; A_(dmalloc(&p, Bm * c->dec[i].out)); c->act_dec.push_back(p); }
    A_(dmalloc(&c->z, Bm * c->latent)); A_(dmalloc(&c->eps, Bm * c->latent));
    A_(dmalloc(&c->gvec, Bm * H));
    A_(dmalloc(&c->P, c->tc_dec ? 1 : Bm * I));
    A_(dmalloc(&c->P16, c->tc_dec ? Bp * I : 1));
    A_(dmalloc(&c->hsT, c->tc_dec ? (int64_t)(H + 8) * Bp : 1));
    A_(dmalloc(&c->h16, c->tc_dec ? Bm * H : 1));
    A_(dmalloc(&c->wd16, c->tc_dec ? I * H : 1));
    A_(dmalloc(&c->dw_scale, 1));
    A_(dmalloc(&c->d_specs, 128));
    A_(dmalloc(&c->spmm_acc, Bm * std::max(c->max_width, H)));
    A_(dmalloc(&c->spmm_ticket, Bm + 1));
    A_(dmalloc(&c->mark, c->enc_in));
    A_(dmalloc(&c->dbuf[0], Bm * c->max_width)); A_(dmalloc(&c->dbuf[1], Bm * c->max_width));
    if (c->tc_hidden) {
        for (auto& L : c->enc) { __half* p = nullptr; int ld = 0; A_(alloc_image(&p, Bm, L.out, &ld)); c->act_enc16.push_back(p); c->ld_enc16.push_back(ld); }
        for (size_t i = 0; i + 1 < c->dec.size(); ++i) {
            __half* p = nullptr; int ld = 0;
            A_(alloc_image(&p, Bm, c->dec[i].out, &ld));
            c->act_dec16.push_back(p); c->ld_dec16.push_back(ld);
        }
        A_(alloc_image(&c->z16, Bm, c->latent, &c->ldz16));
        c->ld_d16 = (int)round_up(c->max_width, 8);
        A_(dmalloc(&c->dbuf16[0], Bm * c->ld_d16)); A_(dmalloc(&c->dbuf16[1], Bm * c->ld_d16));
    }
    c->n_lse_tiles = (int)std::max<int64_t>(std::max<int64_t>(cdiv(I, 64), tc_lse_parts_max((int)I, c->num_sms)), 1);
    A_(dmalloc(&c->part_max, (int64_t)c->n_lse_tiles * Bm)); A_(dmalloc(&c->part_sum, (int64_t)c->n_lse_tiles * Bm));
    c->splitk_elems = c->tc_dec ? (int64_t)64 * Bm * H : 1;
    A_(dmalloc(&c->splitk, c->splitk_elems));
    A_(dmalloc(&c->norms, c->n_tensors)); A_(dmalloc(&c->norm_partial, (int64_t)c->n_tensors * NORM_PARTS));
    A_(dmalloc(&c->d_toff, c->n_tensors)); A_(dmalloc(&c->d_tlen, c->n_tensors));
    A_(dmalloc(&c->loss_dev, 4)); A_(dmalloc(&c->d_err, 1)); A_(dmalloc(&c->lens_tmp, Bm + 1)); A_(dmalloc(&c->lens_tmp2, Bm + 1));
#undef A_
    if (!rc && cudaMemset(c->d_err, 0, sizeof(int)) != cudaSuccess) rc = B200VAE_ECUDA;
    if (!rc && cudaMemset(c->spmm_acc, 0, (size_t)Bm * std::max(c->max_width, H) * sizeof(float)) != cudaSuccess) rc = B200VAE_ECUDA;
    if (!rc && cudaMemset(c->spmm_ticket, 0, (size_t)(Bm + 1) * sizeof(int)) != cudaSuccess) rc = B200VAE_ECUDA;
    if (!rc && cudaMemset(c->mark, 0, (size_t)c->enc_in * sizeof(int32_t)) != cudaSuccess) rc = B200VAE_ECUDA;   // steps start at 1
    if (rc) { free_ctx(c); return rc; }
    *out = reinterpret_cast<b200vae_ctx*>(c);
    return 0;
}

int b200vae_ctx_destroy(b200vae_ctx* ctx) {
    if (!ctx) return 0;
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c->cfg.device);
    cudaDeviceSynchronize();
    free_ctx(c);
    return 0;
}

int b200vae_bind_params(b200vae_ctx* ctx, float* w, float* g, float* m, float* v, int64_t n_elems,
                        const int64_t* w_off, const int64_t* b_off) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && w && g && m && v && w_off && b_off, B200VAE_EINVAL, "null argument");
    c->w = w; c->g = g; c->m = m; c->v = v; c->n_elems = n_elems;
    c->toff.clear(); c->tlen.clear();
    size_t nl = c->enc.size() + c->dec.size();
    for (size_t l = 0; l < nl; ++l) {
        Layer& L = (l < c->enc.size()) ? c->enc[l] : c->dec[l - c->enc.size()];
        L.w_off = w_off[l];
        L.b_off = b_off[l];
        int64_t wl = (int64_t)L.in * L.out;
        B200_REQUIRE(L.w_off >= 0 && L.w_off + wl <= n_elems && L.b_off >= 0 && L.b_off + L.out <= n_elems,
                     B200VAE_EINVAL, "layer %zu offsets fall outside the arena", l);
        B200_REQUIRE(L.w_off % 4 == 0 && L.b_off % 4 == 0, B200VAE_EINVAL, "layer %zu offsets must be multiples of 4 floats", l);
        c->toff.push_back(L.w_off); c->tlen.push_back(wl);
        c->toff.push_back(L.b_off); c->tlen.push_back(L.out);
    }
    if (c->tc_hidden) {
        // the hidden-layer tensors lie between encoder layer 0 and the decoder output layer in the arena
        const Layer& first = c->enc.size() > 1 ? c->enc[1] : c->dec[0];
        c->ws_lo = first.w_off;
        c->ws_hi = c->dec.back().w_off;
        for (size_t l = 1; l + 1 < nl; ++l) {
            const Layer& L = (l < c->enc.size()) ? c->enc[l] : c->dec[l - c->enc.size()];
            B200_REQUIRE(L.w_off >= c->ws_lo && L.b_off + L.out <= c->ws_hi && L.w_off % 8 == 0, B200VAE_EINVAL,
                         "layer %zu lies outside the hidden-layer range of the arena", l);
        }
        if (c->ws16) cudaFree(c->ws16);
        c->ws16 = nullptr;
        B200_CHECK(dmalloc(&c->ws16, c->ws_hi - c->ws_lo));
    }
    B200_CUDA_OK(cudaMemcpy(c->d_toff, c->toff.data(), c->toff.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    B200_CUDA_OK(cudaMemcpy(c->d_tlen, c->tlen.data(), c->tlen.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    c->params_bound = true;
    return b200vae_sync_weights(ctx, nullptr);
}

int b200vae_bind_csr(b200vae_ctx* ctx, int slot, const int64_t* indptr, const int32_t* indices,
                     const float* values, int64_t n_rows) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && (slot == 0 || slot == 1), B200VAE_EINVAL, "bad slot");
    B200_REQUIRE(indptr && (indices || n_rows == 0), B200VAE_EINVAL, "null CSR arrays");
    c->slot[slot].indptr = indptr;
    c->slot[slot].indices = indices;
    c->slot[slot].values = values;
    c->slot[slot].n_rows = n_rows;
    return 0;
}

int b200vae_dense_to_csr(b200vae_ctx* ctx, int slot, const float* dense, int32_t B, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && dense && (slot == 0 || slot == 1), B200VAE_EINVAL, "bad argument");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    CsrSlot& S = c->slot[slot];
    const int width = slot == 0 ? c->enc_in : c->n_items;     // slot 0 = network input (items + condition flags)
    B200_CHECK(launch_dense_count(c, dense, B, width, c->lens_tmp, s));
    B200_CHECK(launch_scan_i64(c, c->lens_tmp, B, c->cfg.max_batch_nnz, S.int_indptr, s));
    B200_CHECK(launch_dense_fill(c, dense, B, width, S.int_indptr, S.int_indices, S.int_values, s));
    S.int_has_values = true;
    return 0;
}

int b200vae_build_cond_batch(b200vae_ctx* ctx, const int32_t* ex_rows, const int32_t* ex_conds, int32_t B,
                             const uint64_t* item_cond_mask, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && ex_rows && ex_conds && item_cond_mask, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    B200_REQUIRE(c->cfg.cond_dim >= 1, B200VAE_ESTATE, "the network has no condition inputs (cond_dim == 0)");
    B200_REQUIRE(c->slot[0].indptr && c->slot[1].indptr, B200VAE_ESTATE, "both CSR slots must be bound");
    return launch_build_cond_batch(c, ex_rows, ex_conds, B, item_cond_mask, (cudaStream_t)stream);
}

int b200vae_expand_batch(b200vae_ctx* ctx, int slot, const int32_t* row_ids, int32_t B, float* out, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && out && (slot == 0 || slot == 1), B200VAE_EINVAL, "bad argument");
    B200_REQUIRE(B >= 0 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    if (B == 0) return 0;
    BatchView v;
    B200_CHECK(make_view(c, slot, row_ids, B, &v, s));
    return launch_expand(c, v, slot == 0 ? c->enc_in : c->n_items, out, s);
}

int b200vae_forward_backward(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B, int32_t B_global,
                             int use_target, float beta, float lam, float dropout_p, uint64_t seed,
                             uint64_t step, int64_t row_offset, const uint8_t* keep_tape,
                             const float* eps_tape, float* loss_out, float* enc0_delta_out, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && loss_out, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200VAE_EINVAL, "dropout_p must be in [0,1)");
    return forward_backward(c, row_ids, B, B_global, use_target, beta, lam, dropout_p, seed, step, row_offset,
                            keep_tape, eps_tape, loss_out, (cudaStream_t)stream, nullptr, enc0_delta_out);
}

int b200vae_enc0_grad(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B_total, const float* delta, float dropout_p,
                      uint64_t seed, uint64_t step, int64_t row_offset, void* stream) {
    // Encoder-0 gradient of a GLOBAL batch from its factors (data parallelism without the dense all-reduce of
    // the [n_items x H1] gradient): the rows of every rank's batch are looked up in the CSR bound to slot 0
    // (which must therefore hold all of them), their normalised / dropped-out input values are recomputed
    // (same Philox keys as the forward pass of the rank that owns them) and scattered against the gathered
    // delta rows.  delta [B_total x H1] already carries the 1/B_global factor.
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && row_ids && delta && B_total >= 1, B200VAE_EINVAL, "bad argument");
    B200_REQUIRE(c->params_bound && c->slot[0].indptr, B200VAE_ESTATE, "parameters / CSR slot 0 are not bound");
    B200_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200VAE_EINVAL, "dropout_p must be in [0,1)");
    const int64_t parts = cdiv(B_total, c->cfg.max_batch);
    const int64_t cap = parts * c->cfg.max_batch_nnz;
    if (B_total > c->gl_rows || cap > c->gl_nnz) {
        B200_CUDA_OK(cudaStreamSynchronize(s));
        if (c->gl_bp) cudaFree(c->gl_bp);
        if (c->gl_sp) cudaFree(c->gl_sp);
        if (c->gl_xt) cudaFree(c->gl_xt);
        c->gl_bp = nullptr; c->gl_sp = nullptr; c->gl_xt = nullptr;
        c->gl_rows = 0; c->gl_nnz = 0;
        B200_CHECK(dmalloc(&c->gl_bp, (int64_t)B_total + 1));
        B200_CHECK(dmalloc(&c->gl_sp, (int64_t)B_total + 1));
        B200_CHECK(dmalloc(&c->gl_xt, cap));
        c->gl_rows = B_total;
        c->gl_nnz = cap;
    }
    const CsrSlot& S = c->slot[0];
    BatchView v;
    v.B = B_total; v.row_ids = row_ids; v.indptr = S.indptr; v.indices = S.indices; v.values = S.values;
    v.bp = c->gl_bp; v.sp = c->gl_sp;
    const int64_t saved_cap = c->cfg.max_batch_nnz;       // the launchers size their grids / bounds from the context
    c->cfg.max_batch_nnz = cap;
    int rc = launch_batch_scan(c, v.indptr, row_ids, B_total, cap, c->gl_bp, c->gl_sp, s);
    if (!rc) rc = launch_batch_prep(c, v, dropout_p, seed, step, row_offset, nullptr, true, c->gl_xt, nullptr, nullptr, 0, s);
    if (!rc) rc = enc0_grad(c, v, c->gl_xt, delta, B_total, s);
    c->cfg.max_batch_nnz = saved_cap;
    return rc;
}

int b200vae_adam_step_range(b200vae_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float lam, int64_t step, int64_t elem_lo, int64_t elem_hi, int narrow, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c, B200VAE_EINVAL, "null context");
    AdamHyper h = {lr, beta1, beta2, eps, weight_decay, lam, step};
    return adam_step(c, h, (cudaStream_t)stream, elem_lo, elem_hi, ADAM_ROWS_ALL, narrow ? c->side_ctas[0] : 8);
}

int b200vae_bind_shadow(b200vae_ctx* ctx, void* wd16, int64_t n_halfs) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && wd16, B200VAE_EINVAL, "null argument");
    if (!c->tc_dec) return 0;                       // the SIMT path keeps no image
    const Layer& DL = c->dec.back();
    B200_REQUIRE(n_halfs >= (int64_t)DL.in * DL.out && ((uintptr_t)wd16 & 15) == 0, B200VAE_EINVAL,
                 "shadow buffer too small (%lld halfs) or not 16-byte aligned", (long long)n_halfs);
    B200_CUDA_OK(cudaDeviceSynchronize());
    if (!c->wd16_external && c->wd16) cudaFree(c->wd16);
    c->wd16 = reinterpret_cast<__half*>(wd16);
    c->wd16_external = true;
    return c->params_bound ? b200vae_sync_weights(ctx, nullptr) : 0;
}

int b200vae_set_w1_sharding(b200vae_ctx* ctx, void* w1_gathered, int32_t mod_n, int32_t mod_r) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c, B200VAE_EINVAL, "null context");
    if (!w1_gathered || mod_n <= 1) {
        c->w1g = nullptr; c->w1_mod_n = 1; c->w1_mod_r = 0; c->w1_rows_per = 0;
        return 0;
    }
    const Layer& E0 = c->enc[0];
    B200_REQUIRE(mod_r >= 0 && mod_r < mod_n && E0.in % mod_n == 0 && E0.out % 4 == 0 && ((uintptr_t)w1_gathered & 15) == 0,
                 B200VAE_EINVAL, "encoder-0 sharding needs n_inputs %% ranks == 0, a hidden width %% 4 == 0 and an aligned buffer");
    c->w1g = reinterpret_cast<__half*>(w1_gathered); c->w1_mod_n = mod_n; c->w1_mod_r = mod_r; c->w1_rows_per = E0.in / mod_n;
    return c->params_bound ? b200vae_sync_weights(ctx, nullptr) : 0;
}

int b200vae_w1_rows(b200vae_ctx* ctx, float* arena_base, float* packed, int direction, void* stream) {
    // direction 0: packed[rows_per x H1] <- this rank's rows of the encoder-0 tensor inside `arena_base` (w, m or v arena)
    // direction 1: every row of that tensor <- packed_all [mod_n][rows_per x H1] (an all-gathered buffer)
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && arena_base && packed && c->w1_mod_n > 1 && c->params_bound, B200VAE_ESTATE, "encoder-0 sharding is not enabled");
    const Layer& E0 = c->enc[0];
    const int64_t n = (int64_t)E0.in * E0.out;
    if (direction == 0)
        k_w1_pack<<<(unsigned)cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(arena_base + E0.w_off, packed, E0.in, E0.out,
                                                                              c->w1_mod_n, c->w1_rows_per, c->w1_mod_r);
    else
        k_w1_unpack<<<(unsigned)cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(packed, arena_base + E0.w_off, E0.in, E0.out,
                                                                                c->w1_mod_n, c->w1_rows_per);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- data-parallel small exchange: one all-gather instead of four small collectives ------------------------------
// Every rank sends ONE packed record [delta (B x H1) | hidden-layer gradients | b_d gradient | loss[4]]; after the
// all-gather each rank sums the gradient / loss parts over the ranks itself (fixed rank order: bit-identical on all
// ranks) and lays the delta parts out as the [N*B x H1] matrix b200vae_enc0_grad reads.
__global__ void k_dp_pack(float* __restrict__ dst, const float* __restrict__ a, int64_t na, const float* __restrict__ b,
                          int64_t nb, const float* __restrict__ c, int64_t nc) {
    pdl_sync();
    const int64_t n = na + nb + nc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = i < na ? a[i] : (i < na + nb ? b[i - na] : c[i - na - nb]);
}
__global__ void k_dp_unpack(const float* __restrict__ recv, int n_ranks, int64_t stride, int64_t n_delta,
                            float* __restrict__ delta_all, int64_t na, float* __restrict__ a, int64_t nb,
                            float* __restrict__ b, int64_t nc, float* __restrict__ c) {
    pdl_sync();
    const int64_t n_copy = (int64_t)n_ranks * n_delta, n_sum = na + nb + nc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_copy + n_sum; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_copy) {
            const int64_t r = i / n_delta, k = i - r * n_delta;
            delta_all[i] = recv[r * stride + k];
        } else {
            const int64_t k = i - n_copy;
            float acc = 0.f;
            for (int r = 0; r < n_ranks; ++r) acc += recv[(int64_t)r * stride + n_delta + k];
            if (k < na) a[k] = acc;
            else if (k < na + nb) b[k - na] = acc;
            else c[k - na - nb] = acc;
        }
    }
}

int b200vae_dp_pack(float* dst, const float* a, int64_t na, const float* b, int64_t nb, const float* c, int64_t nc,
                    void* stream) {
    B200_REQUIRE(dst && na >= 0 && nb >= 0 && nc >= 0, B200VAE_EINVAL, "bad argument");
    const int64_t n = na + nb + nc;
    if (n == 0) return 0;
    B200_CUDA_OK(launch_pdl(k_dp_pack, dim3((unsigned)std::min<int64_t>(cdiv(n, 256), 1184)), dim3(256), 0, (cudaStream_t)stream,
                            dst, a, na, b, nb, c, nc));
    return 0;
}

int b200vae_dp_unpack(const float* recv, int32_t n_ranks, int64_t stride, int64_t n_delta, float* delta_all, float* a,
                      int64_t na, float* b, int64_t nb, float* c, int64_t nc, void* stream) {
    B200_REQUIRE(recv && n_ranks >= 1 && stride >= n_delta + na + nb + nc, B200VAE_EINVAL, "bad argument");
    const int64_t n = (int64_t)n_ranks * n_delta + na + nb + nc;
    if (n == 0) return 0;
    B200_CUDA_OK(launch_pdl(k_dp_unpack, dim3((unsigned)std::min<int64_t>(cdiv(n, 256), 2368)), dim3(256), 0, (cudaStream_t)stream,
                            recv, (int)n_ranks, stride, n_delta, delta_all, na, a, nb, b, nc, c));
    return 0;
}

int b200vae_defer_wait_event(b200vae_ctx* ctx, void* event) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c, B200VAE_EINVAL, "null context");
    c->wd16_pending = reinterpret_cast<cudaEvent_t>(event);
    return 0;
}

int b200vae_sync_weights(b200vae_ctx* ctx, void* stream) {
    // refresh every derived copy of the parameters (the fp16 image of W_d read by the tensor cores);
    // call after the weight arena was modified by anything other than b200vae_adam_step
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && c->params_bound, B200VAE_ESTATE, "bind_params has not been called");
    if (c->w1_mod_n > 1) {      // the gathered copy of the encoder-0 weight, from the (complete) arena
        const Layer& E0 = c->enc[0];
        const int64_t n = (int64_t)E0.in * E0.out;
        k_w1_image<<<(unsigned)cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(c->w + E0.w_off, c->w1g, E0.in, E0.out, c->w1_mod_n,
                                                                               c->w1_rows_per);
        B200_CUDA_OK(cudaGetLastError());
    }
    if (!c->tc_dec) return 0;
    const Layer& DL = c->dec.back();
    if (c->tc_hidden)
        B200_CHECK(launch_to_f16(c, c->w + c->ws_lo, c->ws16, 1, (int)(c->ws_hi - c->ws_lo), c->ws_hi - c->ws_lo,
                                 (cudaStream_t)stream));
    return launch_to_f16(c, c->w + DL.w_off, c->wd16, DL.out, DL.in, DL.in, (cudaStream_t)stream);
}

int b200vae_adam_step(b200vae_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay,
                      float lam, int64_t step, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c, B200VAE_EINVAL, "null context");
    AdamHyper h = {lr, beta1, beta2, eps, weight_decay, lam, step};
    return adam_step(c, h, (cudaStream_t)stream);
}

int b200vae_adam_step_split(b200vae_ctx* ctx, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float lam, int64_t step, const int32_t* touched_items, int32_t n_touched, int overlap_bits,
                            void* stream) {
    // The Adam schedule of the fused step on caller-provided gradients: rows of the encoder-0 weight listed in
    // `touched_items` are the only ones whose gradient may be non-zero.
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && (touched_items || n_touched == 0) && n_touched >= 0, B200VAE_EINVAL, "bad argument");
    B200_REQUIRE(c->params_bound, B200VAE_ESTATE, "bind_params has not been called");
    AdamHyper h = {lr, beta1, beta2, eps, weight_decay, lam, step};
    h.ov = effective_overlap(c, overlap_bits & 3);     // bit 2 only moves a launch inside the fused step
    if (!h.ov) return adam_step(c, h, s);
    if (lam != 0.f) {
        B200_CHECK(launch_tensor_norms(c, c->w, c->d_toff, c->d_tlen, c->n_tensors, c->norm_partial, c->norms, s));
        c->norms_valid = true;
    }
    if (h.ov & 2) {
        if (n_touched > 0) {
            k_stamp_rows<<<(int)cdiv(n_touched, 256), 256, 0, s>>>(touched_items, n_touched, c->enc_in, c->mark, (int32_t)step);
            note(c, "stamp_rows", s);
        }
        B200_CUDA_OK(cudaEventRecord(c->ev_mark, s));
        B200_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_mark, 0));
        const Layer& e0 = c->enc[0];
        B200_CHECK(adam_step(c, h, c->side, e0.w_off, e0.w_off + (int64_t)e0.in * e0.out, ADAM_ROWS_UNMARKED,
                             c->side_ctas[1]));
    }
    B200_CUDA_OK(cudaEventRecord(c->ev_wd, s));
    return fused_adam_finish(c, h, h.ov, s);
}

int b200vae_train_step(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B, int use_target, float beta,
                       float lam, float dropout_p, uint64_t seed, int64_t step, const uint8_t* keep_tape,
                       const float* eps_tape, float lr, float beta1, float beta2, float eps, float weight_decay,
                       float* loss_out, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && loss_out, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, B200VAE_EINVAL, "dropout_p must be in [0,1)");
    AdamHyper h = {lr, beta1, beta2, eps, weight_decay, lam, step};
    return train_step_fused(c, row_ids, B, use_target, beta, dropout_p, seed, keep_tape, eps_tape, h, loss_out,
                            (cudaStream_t)stream);
}

// ---- context-free helpers used by rectorch_b200.metrics / models.loss_function ---------------
}  // extern "C"
namespace b200 {
Ctx* null_ctx() {
    static Ctx nc;
    static bool init = false;
    if (!init) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) nc.num_sms = prop.multiProcessorCount;
        nc.cfg.max_batch_nnz = INT64_MAX;
        cudaMalloc(reinterpret_cast<void**>(&nc.d_err), sizeof(int));
        cudaMemset(nc.d_err, 0, sizeof(int));
        init = true;
    }
    return &nc;
}
}  // namespace b200
extern "C" {

int b200vae_topk_metrics_csr(const float* scores, int32_t B, int32_t n_items, const int64_t* gt_indptr,
                             const int32_t* gt_indices, const float* gt_values, const int32_t* kinds_dev,
                             const int32_t* ks_dev, int32_t n_metrics, int32_t kmax, float* out,
                             int32_t* topk_idx, void* stream) {
    B200_REQUIRE(scores && gt_indptr && kinds_dev && ks_dev && out && B >= 1, B200VAE_EINVAL, "bad argument");
    BatchView gt;
    gt.indptr = gt_indptr; gt.indices = gt_indices; gt.values = gt_values; gt.row_ids = nullptr; gt.bp = gt_indptr; gt.sp = nullptr; gt.B = B;
    return launch_topk_metrics(null_ctx(), scores, n_items, gt, kinds_dev, ks_dev, n_metrics, kmax, out, topk_idx,
                               (cudaStream_t)stream);
}

int b200vae_expand_rows_raw(const int64_t* indptr, const int32_t* indices, const float* values,
                            const int32_t* row_ids, int32_t B, int32_t n_items, float* out, void* stream) {
    B200_REQUIRE(indptr && out && B >= 0, B200VAE_EINVAL, "bad argument");
    if (B == 0) return 0;
    BatchView v;
    v.indptr = indptr; v.indices = indices; v.values = values; v.row_ids = row_ids; v.bp = nullptr; v.sp = nullptr; v.B = B;
    return launch_expand(null_ctx(), v, n_items, out, (cudaStream_t)stream);
}

int b200vae_dense_to_csr_raw(const float* dense, int32_t B, int32_t n_items, int64_t* lens_tmp, int64_t* indptr,
                             int32_t* indices, float* values, int64_t cap, void* stream) {
    // two-phase: indices == NULL -> only count + scan (indptr[B] = nnz); else fill
    cudaStream_t s = (cudaStream_t)stream;
    Ctx* c = null_ctx();
    B200_REQUIRE(dense && indptr && lens_tmp && B >= 1, B200VAE_EINVAL, "bad argument");
    if (!indices) {
        B200_CHECK(launch_dense_count(c, dense, B, n_items, lens_tmp, s));
        return launch_scan_i64(c, lens_tmp, B, INT64_MAX, indptr, s);
    }
    // dense_fill reads the capacity from the context
    c->cfg.max_batch_nnz = cap;
    int rc = launch_dense_fill(c, dense, B, n_items, indptr, indices, values, s);
    c->cfg.max_batch_nnz = INT64_MAX;
    return rc;
}



int b200vae_train_step_host(b200vae_ctx* ctx, const int64_t* indptr_host, const int32_t* indices_host,
                            const float* values_host, int32_t B, float beta, float lam, float dropout_p,
                            uint64_t seed, int64_t step, float lr, float weight_decay, float* loss_host,
                            void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && indptr_host && indices_host && loss_host, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    int64_t nnz = indptr_host[B];
    B200_REQUIRE(indptr_host[0] == 0 && nnz >= 0 && nnz <= c->cfg.max_batch_nnz, B200VAE_ECAPACITY,
                 "batch nnz %lld exceeds capacity %lld", (long long)nnz, (long long)c->cfg.max_batch_nnz);
    CsrSlot& S = c->slot[0];
    B200_CUDA_OK(cudaMemcpyAsync(S.int_indptr, indptr_host, (size_t)(B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    B200_CUDA_OK(cudaMemcpyAsync(S.int_indices, indices_host, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (values_host)
        B200_CUDA_OK(cudaMemcpyAsync(S.int_values, values_host, (size_t)nnz * sizeof(float), cudaMemcpyHostToDevice, s));
    S.int_has_values = values_host != nullptr;
    AdamHyper h = {lr, 0.9f, 0.999f, 1e-8f, weight_decay, lam, step};
    B200_CHECK(train_step_fused(c, nullptr, B, 0, beta, dropout_p, seed, nullptr, nullptr, h, c->loss_dev, s, true));
    B200_CUDA_OK(cudaMemcpyAsync(loss_host, c->loss_dev, 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
    B200_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int b200vae_predict(b200vae_ctx* ctx, const int32_t* row_ids, int32_t B, int remove_train, int train_mode,
                    float dropout_p, uint64_t seed, uint64_t step, float* scores, float* mu, float* logvar,
                    void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c, B200VAE_EINVAL, "null context");
    return predict(c, row_ids, B, remove_train, train_mode, dropout_p, seed, step, scores, mu, logvar, (cudaStream_t)stream);
}

int b200vae_decode(b200vae_ctx* ctx, const float* z, int32_t B, float* scores, void* stream) {
    // AE_net.decode(z) (nets.py:227-233, 413-417): decoder layers on a caller-provided latent batch
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && z && scores, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(c->params_bound, B200VAE_ESTATE, "bind_params has not been called");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    const float* h = z;
    for (size_t i = 0; i + 1 < c->dec.size(); ++i) {
        B200_CHECK(linear_fwd(c, h, B, c->dec[i], c->act_dec[i], s));
        h = c->act_dec[i];
    }
    const Layer& DL = c->dec.back();
    if (c->tc_dec) {
        TcEpi e;
        B200_CHECK(launch_to_f16(c, h, c->h16, B, DL.in, DL.in, s));
        B200_CHECK(wait_wd16(c, s));
        if (c->predict_transposed) {      // as in predict(): scores^T = W_d h^T, coalesced score stores
            e.transpose_out = 1;
            e.row_bias = c->w + DL.b_off;
            return launch_tc_gemm(c, TC_EPI_STORE, c->wd16, DL.in, 0, c->h16, DL.in, 0, scores, c->n_items, c->n_items, B, DL.in, e, s);
        }
        e.bias = c->w + DL.b_off;
        return launch_tc_gemm(c, TC_EPI_STORE, c->h16, DL.in, 0, c->wd16, DL.in, 0, scores, c->n_items, B, c->n_items, DL.in, e, s);
    }
    return linear_fwd(c, h, B, DL, scores, s);
}

int b200vae_topk_metrics(b200vae_ctx* ctx, const float* scores, const int32_t* gt_row_ids, int32_t B,
                         const int32_t* kinds, const int32_t* ks, int32_t n_metrics, float* out,
                         int32_t* topk_idx, void* stream) {
    // kinds / ks are HOST arrays (a handful of ints); they are staged through the context
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && scores && kinds && ks && out && n_metrics >= 1 && n_metrics <= 64, B200VAE_EINVAL, "bad argument");
    B200_REQUIRE(B >= 1 && B <= c->cfg.max_batch, B200VAE_ECAPACITY, "batch %d exceeds capacity %d", B, c->cfg.max_batch);
    int kmax = 1;
    for (int i = 0; i < n_metrics; ++i) {
        B200_REQUIRE(kinds[i] >= 0 && kinds[i] <= 3 && ks[i] >= 1, B200VAE_EINVAL, "bad metric spec %d", i);
        kmax = std::max(kmax, std::min(ks[i], c->n_items));
    }
    int32_t h_specs[128] = {0};
    for (int i = 0; i < n_metrics; ++i) { h_specs[i] = kinds[i]; h_specs[64 + i] = ks[i]; }
    // pageable source: the copy is staged before cudaMemcpyAsync returns, so the stack buffer is safe
    B200_CUDA_OK(cudaMemcpyAsync(c->d_specs, h_specs, sizeof(h_specs), cudaMemcpyHostToDevice, s));
    BatchView gt;
    B200_CHECK(make_view(c, 1, gt_row_ids, B, &gt, s));
    return launch_topk_metrics(c, scores, c->n_items, gt, c->d_specs, c->d_specs + 64, n_metrics, kmax, out, topk_idx, s);
}

int b200vae_gemm_f16(b200vae_ctx* ctx, const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb,
                     int b_mn_major, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c && A && B && C, B200VAE_EINVAL, "null argument");
    TcEpi e;
    return launch_tc_gemm(c, TC_EPI_STORE, A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, e, (cudaStream_t)stream);
}

int b200vae_dec_fwd_lse(b200vae_ctx* ctx, const void* h16, const void* W16, const float* bias, int32_t B,
                        int32_t n_items, int32_t H, float* lse, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(c && h16 && W16, B200VAE_EINVAL, "null argument");
    B200_REQUIRE(B <= c->cfg.max_batch && n_items <= c->n_items, B200VAE_ECAPACITY, "exceeds context capacity");
    B200_REQUIRE(tc_supported(B, n_items, H, H, H), B200VAE_EINVAL, "shape not supported by the tcgen05 path (H %% 8 != 0?)");
    TcEpi e;
    e.bias = bias;
    e.part_max = c->part_max;
    e.part_sum = c->part_sum;
    e.part_rows = c->n_lse_tiles;
    tick(c, 0, 0, s);
    B200_CHECK(launch_tc_gemm(c, TC_EPI_LSE, h16, H, 0, W16, H, 0, nullptr, 0, B, n_items, H, e, s));
    tick(c, 0, 1, s);
    if (lse)   // lse == NULL: only the fused GEMM + log-sum-exp kernel (the partials stay in the workspace)
        B200_CHECK(launch_lse_merge(c, c->part_max, c->part_sum, tc_lse_parts(B, n_items, H, c->num_sms), B, lse, s));
    return 0;
}

// ---- launch-cost probe (scripts/launch_probe.py): an empty kernel with a given grid / block / dynamic shared memory /
// cluster size, so that the fixed cost of the tcgen05 kernel's launch configuration can be measured in isolation
__global__ void k_probe_empty(int* sink) {
    extern __shared__ uint8_t probe_smem[];
    if (sink && threadIdx.x == 0 && blockIdx.x == 0xFFFFFFu) sink[0] = probe_smem[0];
}
int b200vae_probe_launch(int grid, int threads, int smem_bytes, int cluster, void* stream) {
    static int max_set = 0;
    if (smem_bytes > max_set) {
        B200_CUDA_OK(cudaFuncSetAttribute(k_probe_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        max_set = smem_bytes;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)std::max(1, cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, k_probe_empty, (int*)nullptr));
    return 0;
}

int64_t b200vae_launch_count(b200vae_ctx* ctx, int reset) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    if (!c) return -1;
    int64_t n = c->launches;
    if (reset) c->launches = 0;
    return n;
}

int b200vae_set_deterministic(b200vae_ctx* ctx, int enable) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    if (!c) return B200VAE_EINVAL;
    if (enable && !c->det_count) {
        c->det_rows = std::max(c->enc_in, c->dec.back().out);
        const size_t n = (size_t)c->det_rows + 1, cap = (size_t)std::max<int64_t>(c->cfg.max_batch_nnz, 1);
        if (cudaMalloc(&c->det_count, n * sizeof(int)) != cudaSuccess || cudaMalloc(&c->det_off, n * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&c->det_cursor, n * sizeof(int)) != cudaSuccess || cudaMalloc(&c->det_ent, cap * sizeof(int64_t)) != cudaSuccess ||
            cudaMalloc(&c->det_sorted, cap * sizeof(int64_t)) != cudaSuccess)
        {
            set_error("deterministic mode: workspace allocation failed");
            return B200VAE_ECUDA;
        }
    }
    c->deterministic = enable != 0;
    return 0;
}

int b200vae_set_timing(b200vae_ctx* ctx, int enable) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    if (!c) return B200VAE_EINVAL;
    c->timing = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    return 0;
}

float b200vae_kernel_ms(b200vae_ctx* ctx, int which) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    if (!c || which < 0 || which >= 5 || !c->ev_valid[which]) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, c->ev[which][0], c->ev[which][1]) != cudaSuccess) return -1.f;
    return ms;
}

int b200vae_timing_report(b200vae_ctx* ctx, char* buf, int cap) {
    // "name ms\n" per launch of the most recent instrumented step (valid after a stream sync)
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    if (!c || !buf || cap <= 0) return B200VAE_EINVAL;
    int pos = 0;
    buf[0] = 0;
    for (int i = 1; i < c->tcount; ++i) {
        float ms = 0.f;
        // mode 1: interval since the previous launch (one stream); mode 2: completion time since the step started,
        // with a '+' suffix on the launches that ran on the side stream
        if (cudaEventElapsedTime(&ms, c->tev[c->timing == 2 ? 0 : i - 1], c->tev[i]) != cudaSuccess) ms = -1.f;
        int n = snprintf(buf + pos, cap - pos, "%s%s %.6f\n", c->tnames[i],
                         (c->timing == 2 && c->side && c->tstream[i] == c->side) ? "+" : "", ms);
        if (n <= 0 || n >= cap - pos) break;
        pos += n;
    }
    return c->tcount;
}

int b200vae_wait_wd_ready(b200vae_ctx* ctx, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    B200_REQUIRE(c, B200VAE_EINVAL, "null context");
    B200_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, c->ev_wd, 0));
    return 0;
}

int b200vae_check_error_flag(b200vae_ctx* ctx) {
    // device-side capacity overflow flag (set by the scan kernels); synchronises the device
    Ctx* c = reinterpret_cast<Ctx*>(ctx);
    DeviceGuard dev_guard(c ? c->cfg.device : -1);
    if (!c) return B200VAE_EINVAL;
    int flag = 0;
    B200_CUDA_OK(cudaMemcpy(&flag, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(c->d_err, 0, sizeof(int));
        if (flag & 2) {
            set_error("a target row had |t| > 4 |sum t| (mixed-sign ratings are not multinomial targets): "
                      "the gradients of that step were clamped");
            return B200VAE_EINVAL;
        }
        set_error("a batch exceeded max_batch_nnz (%lld): results of that step are invalid", (long long)c->cfg.max_batch_nnz);
        return B200VAE_ECAPACITY;
    }
    return 0;
}

}  // extern "C"
