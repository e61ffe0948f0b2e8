// common.cuh -- shared declarations of the b200vae engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/b200vae.h"

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CUDA_OK(expr)                                                            \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            b200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                            cudaGetErrorString(_e));                                  \
            return B200VAE_ECUDA;                                                     \
        }                                                                             \
    } while (0)

#define B200_CHECK(rc_expr)                                                           \
    do {                                                                              \
        int _rc = (rc_expr);                                                          \
        if (_rc != 0) return _rc;                                                     \
    } while (0)

#define B200_REQUIRE(cond, code, ...)                                                 \
    do {                                                                              \
        if (!(cond)) {                                                                \
            b200::set_error(__VA_ARGS__);                                             \
            return (code);                                                            \
        }                                                                             \
    } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// A batch of rows of a CSR matrix.  Row r of the batch is CSR row
// (row_ids ? row_ids[r] : r); its non-zeros live at [indptr[gr], indptr[gr+1]).
// `bp` [B+1] is the compact (batch-local) exclusive scan of the row lengths, so
// per-non-zero batch arrays (scaled values, dropout keeps) are indexed bp[r]+k.
struct BatchView {
    const int64_t* indptr;
    const int32_t* indices;
    const float*   values;    // nullable -> 1.0f
    const int32_t* row_ids;   // nullable -> identity
    const int64_t* bp;        // [B+1]
    const int32_t* sp;        // [B+1] exclusive scan of the per-row segment counts (SPMM_SEG non-zeros each)
    int32_t        B;
};

constexpr int SPMM_SEG = 32;
constexpr int NORM_PARTS = 512;   // partial sums per tensor in the parameter-norm reduction   // non-zeros per sparse gather/scatter work unit

struct Layer {
    int in, out;              // nn.Linear(in, out)
    int64_t w_off, b_off;     // offsets into the arenas
    bool tanh_act;            // activation applied after this layer in the forward pass
};

// Epilogue description for the SIMT GEMM (simt_gemm.cu).
enum { EPI_STORE = 0, EPI_LSE = 1, EPI_PROB = 2 };
struct GemmEpi {
    const float* bias = nullptr;      // per column n
    int   act = 0;                    // 1 = tanh
    const float* mulY = nullptr;      // out *= (1 - Y[m,n]^2)
    int64_t ldy = 0;
    float alpha = 1.f;
    const float* addend = nullptr;    // out += addend_scale * addend[m,n]
    int64_t ld_addend = 0;
    float addend_scale = 0.f;
    float* part_max = nullptr;        // EPI_LSE: [n_tiles x M]
    float* part_sum = nullptr;
    const float* lse = nullptr;       // EPI_PROB
    const float* rowscale = nullptr;  // EPI_PROB: P = exp(logit - lse[m]) * rowscale[m]
};

struct Ctx;   // engine.cu

// ---- kernel launchers (each returns a B200VAE_* code) -----------------------------------
// sparse.cu
int launch_batch_scan(Ctx* c, const int64_t* indptr, const int32_t* row_ids, int B, int64_t cap,
                      int64_t* bp, int32_t* sp, cudaStream_t s);
int launch_batch_prep(Ctx* c, const BatchView& in, float p, uint64_t seed, uint64_t step,
                      int64_t row_offset, const uint8_t* keep_tape, bool train, float* xt, float* row_sum_out,
                      int32_t* mark, int32_t mark_step, cudaStream_t s, bool with_scan = false);
int launch_build_cond_batch(Ctx* c, const int32_t* ex_rows, const int32_t* ex_conds, int B,
                            const uint64_t* item_cond_mask, cudaStream_t s);
// what k_loss_final computes, handed to the fix-up kernel so that its last CTA closes the loss (no extra launch)
struct LossTail {
    float* loss_out;        // nullptr: the fix-up kernel stops after the per-row terms
    const float* kl_row;
    const float* norms;
    int* ticket;            // one zero-initialised int; the closing CTA resets it
    int B, n_tensors;
    float inv_Bg, beta, lam;
};
int launch_target_fixup(Ctx* c, const BatchView& tgt, __half* PT, int64_t ldp, const float* T, const float* lse,
                        const __half* h16, int64_t ldh, const __half* W16, int64_t ldw, const float* bias, int H,
                        float* loss_row, cudaStream_t s, const LossTail* tail = nullptr);
int launch_row_sums(Ctx* c, const BatchView& v, float* out, cudaStream_t s);
// Wt16 != NULL (then mod_n > 1): the weight rows are read from the rank-interleaved gathered fp16 image
// [mod_n][rows_per x H] (row j lives at block j % mod_n, index j / mod_n) instead of Wt
int launch_spmm_gather(Ctx* c, const BatchView& v, const float* vals, const float* Wt, int H,
                       const float* bias, int act, float* out, cudaStream_t s, __half* out16 = nullptr, int64_t ld16 = 0,
                       const __half* Wt16 = nullptr, int mod_n = 1, int64_t rows_per = 0);
// mod_n > 1: only the rows (items) j with j % mod_n == mod_r are written
int launch_spmm_scatter(Ctx* c, const BatchView& v, const float* vals, float scale, const float* dY,
                        int H, float* dWt, cudaStream_t s, int mod_n = 1, int mod_r = 0);
int launch_spmm_scatter_bias(Ctx* c, const BatchView& v, const float* vals, float scale, const float* dY,
                             int H, float* dWt, float* db, cudaStream_t s, int mod_n = 1, int mod_r = 0);
int launch_dense_count(Ctx* c, const float* dense, int B, int I, int64_t* lens, cudaStream_t s);
int launch_dense_fill(Ctx* c, const float* dense, int B, int I, const int64_t* indptr,
                      int32_t* indices, float* values, cudaStream_t s);
int launch_scan_i64(Ctx* c, const int64_t* lens, int n, int64_t cap, int64_t* out, cudaStream_t s);
int launch_expand(Ctx* c, const BatchView& v, int I, float* out, cudaStream_t s);
int launch_mask_seen(Ctx* c, const BatchView& v, int I, float* scores, cudaStream_t s);
int launch_row_loss(Ctx* c, const BatchView& tgt, const float* h, const float* gvec, int H,
                    const float* bias, const float* pmax, const float* psum, int n_tiles, float* lse,
                    const float* T, float inv_Bg, float* loss_row, float* rowscale, cudaStream_t s);

// simt_gemm.cu
int launch_simt_gemm(Ctx* c, int mode, const float* A, int64_t a_rs, int64_t a_cs, const float* B,
                     int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, int M, int N, int K,
                     const GemmEpi& e, cudaStream_t s);
int launch_colsum(Ctx* c, const float* X, int64_t ldx, int M, int N, float* out, cudaStream_t s);
int launch_lse_merge(Ctx* c, const float* pmax, const float* psum, int n_tiles, int M, float* lse,
                     cudaStream_t s);

// elementwise.cu
int launch_reparam_kl(Ctx* c, const float* enc_out, int B, int L, bool train, const float* eps_tape,
                      uint64_t seed, uint64_t step, int64_t row_offset, const int32_t* row_ids,
                      float* z, float* eps_out, float* kl_row, __half* z16, int64_t ldz16, cudaStream_t s);
int launch_dz_to_denc(Ctx* c, const float* dz, const float* enc_out, const float* eps, int B, int L,
                      float beta_over_B, bool train, float* denc, __half* denc16, int64_t ld16, cudaStream_t s);
int launch_loss_final(Ctx* c, const float* loss_row, const float* kl_row, int B, float inv_Bg,
                      float beta, float lam, const float* norms, int n_tensors, float* loss_out,
                      cudaStream_t s);
int launch_tensor_norms(Ctx* c, const float* w, const int64_t* offs, const int64_t* lens,
                        int n_tensors, float* partial, float* norms, cudaStream_t s);
// Row filter + launch footprint of one Adam launch.  Elements inside the re-zero window [z_lo, z_hi) (the
// encoder-0 weight, rows of `row_len` floats) are filtered by the per-row step marks batch_prep wrote:
//   ADAM_ROWS_ALL       every row (gradient read from g)
//   ADAM_ROWS_MARKED    only rows with mark == mark_step (the rows this step's sparse scatter wrote)
//   ADAM_ROWS_UNMARKED  only the other rows; their gradient is exactly zero, so g is neither read nor re-zeroed
// ctas_per_sm x threads is the grid-stride footprint: the full-width default for a launch that owns the GPU,
// a narrow one for launches that share it with the other stream's kernels.
//   ADAM_ROWS_MOD       only rows r with r % mod_n == mod_r (the rank's shard of the encoder-0 weight under data
//                       parallelism); the updated rows are also written, packed, into block mod_r of w1g
enum { ADAM_ROWS_ALL = 0, ADAM_ROWS_MARKED = 1, ADAM_ROWS_UNMARKED = 2, ADAM_ROWS_MOD = 3 };
struct AdamW1 {                     // ADAM_ROWS_MOD (by value to the kernel)
    __half* w1g;
    int mod_n, mod_r;
    int64_t rows_per;
    int stream;                     // 1: w / m / v go through L2 with evict-first hints (AdamOpt::stream)
};
struct AdamOpt {
    const int32_t* mark = nullptr;
    int32_t mark_step = 0;
    int filter = ADAM_ROWS_ALL;
    int row_len = 0;
    int ctas_per_sm = 8;
    int threads = 256;
    __half* w1g = nullptr;          // ADAM_ROWS_MOD: gathered fp16 image of the encoder-0 weight [mod_n][rows_per x row_len]
    int mod_n = 1, mod_r = 0;
    int64_t rows_per = 0;
    bool stream = false;            // the gradient of this range is expected in L2 (written just before by its producer):
                                    // load / store w, m, v with evict-first hints so they do not push it out
    __half* shadow2 = nullptr;      // fp16 image of the hidden-layer tensors: elements [s2_lo, s2_hi) of THIS launch
    int64_t s2_lo = 0, s2_hi = 0;   // (indices relative to the launch's base pointers) go to shadow2[e - s2_lo]
};
int launch_adam(Ctx* c, float* w, float* g, float* m, float* v, int64_t n, float lr_over_bc1,
                float beta1, float beta2, float bc2_sqrt, float eps, float wd, float lam,
                const float* norm_ptr, __half* shadow, int64_t sh_lo, int64_t sh_hi, int64_t z_lo, int64_t z_hi,
                const AdamOpt& opt, cudaStream_t s);
int launch_discard_l2(Ctx* c, const float* p, int64_t n, cudaStream_t s);
int launch_to_f16(Ctx* c, const float* x, __half* y, int64_t rows, int cols, int64_t ldy, cudaStream_t s);

// topk.cu
int launch_topk_metrics(Ctx* c, const float* scores, int I, const BatchView& gt, const int32_t* kinds,
                        const int32_t* ks, int n_metrics, int kmax, float* out, int32_t* topk_idx,
                        cudaStream_t s);

// tc_gemm.cu  (tcgen05 / TMEM / TMA; fp16 operands, fp32 accumulate)
enum { TC_EPI_STORE = 0, TC_EPI_LSE = 1, TC_EPI_PROB = 2 };
constexpr float PROB_LOG2_SCALE = 14.f;   // P~ = softmax * 2^14 (fp16: <= 16384, normal down to 3.7e-9)
constexpr float HS_LOG2_SCALE = 8.f;      // hs = h * (T_u / max T) * 2^8
struct TcEpi {
    const float* bias = nullptr;       // per column
    float* part_max = nullptr;         // TC_EPI_LSE partials [part_rows x M]
    float* part_sum = nullptr;
    int part_rows = 0;                 // capacity of the partial buffers (rows of M floats)
    const float* lse = nullptr;        // TC_EPI_PROB
    float prob_log2_scale = PROB_LOG2_SCALE;
    const float* row_bias = nullptr;   // TC_EPI_STORE with transpose_out: added per ROW m of the product (C[n*ldc + m] += row_bias[m])
    float* bias_grad = nullptr;        // TC_EPI_STORE: column `bias_col` of the product goes here
    int bias_col = -1;
    int split_k = 1;                   // TC_EPI_STORE: partial products C[s] at C + s*split_stride
    int64_t split_stride = 0;
    int transpose_out = 0;             // TC_EPI_STORE: write C[n*ldc + m] (a warp stores 32 consecutive m = 128 B);
                                       // bias_col then names the ROW m whose values go to bias_grad[n]
    int n_fastest = 0;                 // walk N tiles first (A tile shared through L2 by consecutive CTAs)
    float out_scale = 1.f;             // TC_EPI_STORE: product * out_scale * (*out_scale_ptr)
    const float* out_scale_ptr = nullptr;
    // TC_EPI_STORE, plain (non-transposed, no split-K) output only -- the hidden layers (nets.py:398-404, 413-416):
    int act = 0;                       // 1 = tanh after the bias
    const float* mulY = nullptr;       // out *= 1 - Y[m,n]^2   (tanh' of the layer below, backward pass)
    int64_t ldy = 0;
    __half* C16 = nullptr;             // fp16 image of the output for the next tensor-core GEMM (pitch ldc16)
    int64_t ldc16 = 0;
    int accumulate = 0;                // TC_EPI_STORE plain output: C += product (EASE Gram matrix over user chunks)
};
// pitches (lda, ldb) in halfs, multiples of 8; A / B are __half arrays; C is float (STORE) or __half (PROB)
bool tc_supported(int M, int N, int K, int64_t lda, int64_t ldb);
int launch_tc_gemm(Ctx* c, int mode, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb,
                   int b_mn, void* C, int64_t ldc, int M, int N, int K, const TcEpi& e,
                   cudaStream_t s);
int tc_lse_parts(int M, int N, int K, int num_sms);   // (max, sum) partial rows per user the LSE epilogue writes
int tc_lse_parts_max(int N, int num_sms);             // upper bound over all batch sizes / hidden widths
int tc_output_tiles(int M, int N);             // output tiles of the current tiling (256-row pair tiles)
int tc_parallel_tiles(int num_sms);            // tiles that run concurrently (SM pairs)
int launch_splitk_reduce(Ctx* c, const float* parts, int n_split, int64_t split_stride, float* out,
                         int64_t ld_out, int M, int N, int64_t ld_part, const float* addend,
                         int64_t ld_add, float addend_scale, const float* mulY, int64_t ldy,
                         const float* rowscale, float scale, cudaStream_t s, __half* out16 = nullptr, int64_t ld16 = 0);

// ---- programmatic dependent launch ---------------------------------------------------------
// A training step is a chain of ~20 short kernels on one stream; between two of them the GPU normally idles for the
// launch latency of the next.  Every hot-path kernel therefore (a) is launched with the programmatic-stream-
// serialization attribute (launch_pdl) and (b) starts with pdl_sync(): griddepcontrol.wait blocks until the previous
// kernel of the stream has completed and its writes are visible -- nothing before it touches global memory -- and
// griddepcontrol.launch_dependents lets the NEXT kernel be scheduled (its CTAs then park at their own wait).
// The ordering guarantees are those of a plain stream; only the launch latency is overlapped.  B200VAE_PDL=0 turns
// the attribute off (the device-side instructions are then no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- device helpers ----------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0,k1).
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
// tensor-core operands are fp16 images (round to nearest) of fp32 values: keep them finite
__device__ __forceinline__ float f16_clamp(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ float u32_to_unit(uint32_t x) {   // (0,1]
    return (float)(x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
}
#endif

}  // namespace b200
