// tc_gemm.cu -- the item-sized contractions of the decoder output layer on the 5th-gen
// tensor cores: tcgen05.mma.cta_group::2 (kind::f16: fp16 operands, fp32 accumulate in TMEM),
// operands staged in shared memory by TMA (128 B swizzle), warp-specialised persistent CTA pairs.
//
//   D[M x N] = A[M x K] * B[N x K]^T
//
//   TC_EPI_LSE   K4  logits = h W_d^T + b never leave the SM: each epilogue thread owns one user
//                    row (one TMEM lane) and folds its share of every item tile into an online
//                    (max, sum exp); output = a few (max, sum) partials per user, merged by
//                    row_loss.                  (F.log_softmax over nets.py:417's output, models.py:813)
//   TC_EPI_PROB  K5  same mainloop, epilogue recomputes softmax from the saved lse and stores
//                    P~^T[item, user] = softmax * 2^S as fp16 (coalesced: a warp's 32 lanes are 32
//                    consecutive users)                              (dlogits of loss.backward())
//   TC_EPI_STORE     plain product * out_scale (+ bias), optional split-K partials and a "bias row":
//                    dW_d|db_d = P~^T [h*T/B | T/B]  and  dh = P~ W_d,  predict scores.
//
// Why fp16 and not tf32 operands.  Both carry a 10-bit mantissa; what differs is the exponent range
// (5 bits vs 8).  Every operand of these GEMMs has a bounded range -- tanh outputs, weights, softmax
// probabilities times an exact power of two -- so fp16 with round-to-nearest conversion in the producers
// has the precision of tf32 at half the bytes through HBM, L2 and shared memory and at twice the
// tensor-core rate.  Round 1's tf32 version of this kernel moved 4.0x its DRAM bytes over the L2->SM
// crossbar (489 MB for 121 MB) and was bound by it.
//
// Operand "majorness": K-major = the contraction index is contiguous in memory (TMA box 64 halfs of K
// x rows, SWIZZLE_128B), MN-major = the M/N index is contiguous (box 64 halfs of M/N x 64 K-rows,
// one box per 64-wide chunk, SWIZZLE_128B).  The instruction descriptor's a_major/b_major bits select
// the interpretation; the shared-memory descriptors follow cute::UMMA's canonical layouts.
//
// Resident A (TcArgs::resident).  With 2-byte operands the pair's activation slice (128 rows x K per CTA:
// 154 KB at K = 600) fits in shared memory next to a 4-deep ring of W_d stages, so it is loaded ONCE per
// pass and only W_d streams: the L2->SM traffic of K4 drops from 4.0x to ~1.0x of its DRAM bytes (x number
// of 256-user row groups).  The pairs are split into one group per 256-row block of A; the pairs of a
// group share the N tiles round-robin and keep their A slice for all of them.
//
// Pipelines (all mbarrier based, no __syncthreads in the steady state):
//   warp 0    TMA producer     : empty[s] -> issue loads -> full[s] (complete_tx on the LEADER's barrier)
//   warp 1    MMA issuer       : (leader CTA only) full[s] -> 4 x tcgen05.mma (K=16 each) -> commit ->
//                                empty[s] (multicast to both CTAs); after the last K block -> tmem_full[a]
//   warps 2.. epilogue         : tmem_full[a] -> tcgen05.ld 32 columns at a time, software pipelined
//                                -> math -> global stores -> tmem_empty[a] (on the leader).
//                                Warps sharing a TMEM lane quarter interleave the column chunks.
// TMEM: 512 columns = 2 accumulator stages x 256 columns, so the epilogue of tile i overlaps
// the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include "ctx.cuh"

namespace b200 {

constexpr int TC_BM = 128;                      // A rows per CTA; the pair's UMMA is M = 256
constexpr int TC_BK = 64;                       // halfs per K block = 128 B = one swizzle row
constexpr int TC_KBLK_BYTES = TC_BM * 128;      // 16 KB: one K block of a 128-row operand tile
constexpr int TC_PIPE_BYTES = 220 * 1024;       // operand storage: ring (+ resident A slice)
constexpr int TC_MAX_STAGES = 12;
constexpr int TC_BAR_BYTES = 512;
constexpr int TC_BIAS_BYTES = 2 * 256 * 4;
constexpr int TC_SMEM = TC_PIPE_BYTES + 1024 /*align*/ + TC_BAR_BYTES + TC_BIAS_BYTES;
static_assert(TC_SMEM <= 227 * 1024, "exceeds the per-CTA shared memory limit of sm_100");
// Epilogue warps per CTA: 2 per TMEM lane quarter for the register-hungry log-sum-exp epilogue, 4 per quarter for
// the store-heavy epilogues (more stores in flight).  Warps sharing a quarter interleave the 32-column chunks.
__host__ __device__ constexpr int tc_epi_warps(int mode) { return mode == TC_EPI_LSE ? 8 : 16; }
__host__ __device__ constexpr int tc_threads(int mode) { return 64 + 32 * tc_epi_warps(mode); }
constexpr unsigned long long SPIN_LIMIT = 1ull << 28;

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    unsigned long long spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) {   // a protocol bug must fail loudly, not hang the GPU
            printf("b200vae tc_gemm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default (cta-scope release) semantics: a cluster-scope release would drain this thread's
    // outstanding TMA traffic first and serialise the peer's load stream (measured: 2x slower GEMM)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on a barrier given as a shared::cluster address
// (the leader CTA's full barrier)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint32_t bar) {   // arrives on the same barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// TMEM -> registers, 32 lanes x 32 columns per warp; asynchronous until tmem_wait_ld
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// wait for all outstanding tcgen05.ld of this thread; the registers are passed through the asm so
// that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// SM100 shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout:
// start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1 layout_type[61,64)), layout_type 2 = SWIZZLE_128B
// (8 rows x 128 B atoms, 16 B swizzle granules) for both majornesses of a 16-bit operand:
//   K-major : row = one M/N index, 64 halfs of K.        SBO = 1024 (stride between 8-row groups), LBO unused;
//             the 4 MMAs of a K block start 32 B apart inside the swizzle row.
//   MN-major: row = one K index, 64 halfs of M/N (cute: ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in halfs).
//             SBO = 1024 (stride between 8-K-row groups), LBO = 8192 (stride between 64-wide M/N chunks, one TMA
//             box of 64 K rows each); the 4 MMAs of a K block start 16 K rows = 2048 B apart.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
struct TcArgs {
    void* C;                // STORE: float [M x ldc] (or its transpose / split-K partials); PROB: __half [N x ldc]
    int64_t ldc;
    int M, N, K;
    int BN;                 // N tile (multiple of 16, <= 256); each CTA of the pair stages BN/2 rows of B
    int tiles_m, tiles_n;   // tiles_m counts 256-row pair tiles
    int n_fastest;          // streaming tile order: consecutive pairs walk N first (share the A tile through L2)
    int kb_total;           // K blocks overall
    int kb_per_split;
    int split_k;
    int64_t split_stride;
    int n_store;            // STORE: columns < n_store go to C
    int bias_col;           // STORE: this column goes to bias_grad[m] (or -1)
    int transpose_out;      // STORE: C[n*ldc + m]; rows m < n_store are stored, row m == bias_col -> bias_grad[n]
    const float* bias;
    float* part_max;
    float* part_sum;
    const float* lse;
    float* bias_grad;
    const float* row_bias;  // STORE + transpose_out: per-row (m) addend, or nullptr
    float out_scale;        // STORE: D * out_scale (* *out_scale_ptr): exact power-of-two un-scaling of fp16 operands
    const float* out_scale_ptr;
    int act;                // STORE: tanh after the bias
    const float* mulY;      // STORE: out *= 1 - Y^2
    int64_t ldy;
    __half* C16;            // STORE: fp16 image of the output
    int64_t ldc16;
    int accumulate;         // STORE: C += product
    float prob_log2_scale;  // PROB: S of P~ = softmax * 2^S
    int vec_ok;             // C rows are 16 B aligned
    int dbg;                // B200VAE_TC_DBG bit mask (probes, garbage results): 1 = no operand loads,
                            // 2 = epilogue only waits and releases the accumulator, 4 = no MMAs are issued,
                            // 8 = no tiles at all (launch + prologue + teardown only), 16 = with 8: also no TMEM
                            // allocation and no cluster barriers
    // schedule
    int resident;           // the A slice stays in shared memory for all N tiles of a pass
    int n_stages;           // ring depth
    int stage_bytes;        // ring slot size (multiple of 1024)
    int ring_off;           // byte offset of the ring behind the resident A slice
    int b_bytes;            // bytes of B one CTA stages per K block
    int pairs_per_m;        // resident: pairs sharing one 256-row block of A
    int n_groups;           // resident: 256-row blocks processed concurrently
};

// Tile walk of one CTA pair; the producer, the MMA issuer and the epilogue warps all run the same walk.
struct TileWalk {
    int m_idx, n_idx, sp;
    bool pass_first, pass_last;   // resident mode: first / last tile that uses the current A slice
    int pass;
    int t, t_step, r;
};
__device__ __forceinline__ void walk_decode(TileWalk& w, const TcArgs& a) {
    const int mn = a.tiles_m * a.tiles_n;
    w.sp = w.t / mn;
    const int q = w.t - w.sp * mn;
    if (a.n_fastest) { w.n_idx = q % a.tiles_n; w.m_idx = q / a.tiles_n; }
    else             { w.m_idx = q % a.tiles_m; w.n_idx = q / a.tiles_m; }
}
__device__ __forceinline__ bool walk_begin(TileWalk& w, const TcArgs& a, int pair, int npairs) {
    w.pass = 0; w.t = 0; w.t_step = 0;
    if (a.resident) {
        const int g = pair / a.pairs_per_m;
        w.r = pair - g * a.pairs_per_m;
        if (g >= a.n_groups) return false;
        w.m_idx = g; w.n_idx = w.r; w.sp = 0;
        w.pass_first = true;
        w.pass_last = (w.n_idx + a.pairs_per_m >= a.tiles_n);
        return true;
    }
    w.t = pair; w.t_step = npairs; w.r = 0;
    w.pass_first = w.pass_last = false;
    if (w.t >= a.tiles_m * a.tiles_n * a.split_k) return false;
    walk_decode(w, a);
    return true;
}
__device__ __forceinline__ bool walk_next(TileWalk& w, const TcArgs& a) {
    if (a.resident) {
        w.n_idx += a.pairs_per_m;
        w.pass_first = false;
        if (w.n_idx >= a.tiles_n) {
            w.m_idx += a.n_groups;
            if (w.m_idx >= a.tiles_m) return false;
            w.n_idx = w.r;
            w.pass++;
            w.pass_first = true;
        }
        w.pass_last = (w.n_idx + a.pairs_per_m >= a.tiles_n);
        return true;
    }
    w.t += w.t_step;
    if (w.t >= a.tiles_m * a.tiles_n * a.split_k) return false;
    walk_decode(w, a);
    return true;
}

constexpr float LOG2E_F = 1.4426950408889634f;

// one 32-column chunk of the accumulator for the thread's row
template <int MODE>
__device__ __forceinline__ void epi_chunk(const TcArgs& a, const float* v, const float* __restrict__ bias_s, int c0, int nc,
                                          int n0, int m, int sp, float& run_max, float& run_sum, float off_l2, float oscale) {
    if (MODE == TC_EPI_LSE) {
        float x[32];
        if (nc == 32) {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + i);
                x[i] = v[i] + b4.x; x[i + 1] = v[i + 1] + b4.y; x[i + 2] = v[i + 2] + b4.z; x[i + 3] = v[i + 3] + b4.w;
                m4[0] = fmaxf(m4[0], x[i]); m4[1] = fmaxf(m4[1], x[i + 1]);
                m4[2] = fmaxf(m4[2], x[i + 2]); m4[3] = fmaxf(m4[3], x[i + 3]);
            }
            const float new_max = fmaxf(run_max, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
            const float off = new_max * LOG2E_F;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                s4[0] += ex2_approx(fmaf(x[i], LOG2E_F, -off));
                s4[1] += ex2_approx(fmaf(x[i + 1], LOG2E_F, -off));
                s4[2] += ex2_approx(fmaf(x[i + 2], LOG2E_F, -off));
                s4[3] += ex2_approx(fmaf(x[i + 3], LOG2E_F, -off));
            }
            run_sum = run_sum * ex2_approx((run_max - new_max) * LOG2E_F) + ((s4[0] + s4[1]) + (s4[2] + s4[3]));
            run_max = new_max;
        } else {
            float cmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                x[i] = (i < nc) ? v[i] + bias_s[c0 + i] : -INFINITY;
                cmax = fmaxf(cmax, x[i]);
            }
            const float new_max = fmaxf(run_max, cmax);
            const float off = new_max * LOG2E_F;
            float csum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < nc) csum += ex2_approx(fmaf(x[i], LOG2E_F, -off));
            run_sum = run_sum * ex2_approx((run_max - new_max) * LOG2E_F) + csum;
            run_max = new_max;
        }
    } else if (MODE == TC_EPI_PROB) {
        // P~^T[(n0+c0+i) * ldc + m] = softmax * 2^S as fp16: for a fixed i the warp's 32 lanes write 32
        // consecutive halfs (64 contiguous bytes)
        if (m < a.M) {
            __half* dst = reinterpret_cast<__half*>(a.C) + (int64_t)(n0 + c0) * a.ldc + m;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (i < nc) {
                    const float x = v[i] + bias_s[c0 + i];
                    dst[(int64_t)i * a.ldc] = __float2half_rn(ex2_approx(fmaf(x, LOG2E_F, off_l2)));
                }
            }
        }
    } else if (a.transpose_out) {
        // D^T: for a fixed column the warp's 32 lanes (consecutive m) write 32 consecutive floats
        if (m < a.M) {
            if (m < a.n_store) {
                float* dst = reinterpret_cast<float*>(a.C) + (int64_t)(n0 + c0) * a.ldc + m;
                const float rb = a.row_bias ? a.row_bias[m] : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nc) dst[(int64_t)i * a.ldc] = fmaf(v[i], oscale, rb);
            } else if (m == a.bias_col) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nc) a.bias_grad[n0 + c0 + i] = v[i] * oscale;
            }
        }
    } else {
        if (m < a.M) {
            float* crow = reinterpret_cast<float*>(a.C) + (int64_t)sp * a.split_stride + (int64_t)m * a.ldc;
            const int col0 = n0 + c0;
            if (a.act || a.mulY || a.C16) {
                // hidden-layer epilogue: bias, tanh / tanh', fp32 result + fp16 image for the next GEMM
                const float* yrow = a.mulY ? a.mulY + (int64_t)m * a.ldy : nullptr;
                __half* hrow = a.C16 ? a.C16 + (int64_t)m * a.ldc16 : nullptr;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = col0 + i;
                    if (i < nc) {
                        if (col < a.n_store) {
                            float y = fmaf(v[i], oscale, bias_s[c0 + i]);
                            if (a.act) y = tanhf(y);
                            if (yrow) { const float t = yrow[col]; y *= (1.f - t * t); }
                            crow[col] = y;
                            if (hrow) hrow[col] = __float2half_rn(f16_clamp(y));
                        } else if (col == a.bias_col) {
                            a.bias_grad[m] = v[i] * oscale;
                        }
                    }
                }
            } else if (a.vec_ok && nc == 32 && col0 + 32 <= a.n_store) {   // nc < 32: the N tile ends inside this chunk
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + i);
                    if (a.accumulate) {
                        const float4 o4 = *reinterpret_cast<const float4*>(crow + col0 + i);
                        b4.x += o4.x; b4.y += o4.y; b4.z += o4.z; b4.w += o4.w;
                    }
                    *reinterpret_cast<float4*>(crow + col0 + i) =
                        make_float4(fmaf(v[i], oscale, b4.x), fmaf(v[i + 1], oscale, b4.y), fmaf(v[i + 2], oscale, b4.z),
                                    fmaf(v[i + 3], oscale, b4.w));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = col0 + i;
                    if (i < nc) {
                        if (col < a.n_store) crow[col] = fmaf(v[i], oscale, bias_s[c0 + i]) + (a.accumulate ? crow[col] : 0.f);
                        else if (col == a.bias_col) a.bias_grad[m] = v[i] * oscale;
                    }
                }
            }
        }
    }
}

template <int MODE, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(tc_threads(MODE), 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    constexpr int EPI_WARPS = tc_epi_warps(MODE);
    constexpr int EPI_SUB = EPI_WARPS / 4;
    constexpr int EPI_THREADS = 32 * EPI_WARPS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + TC_PIPE_BYTES;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = (cta_rank == 0);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TC_MAX_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + s); };
    const uint32_t ares_empty = bar_base + 8u * (2 * TC_MAX_STAGES + 4);     // resident A slice no longer read
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 5);
    uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gen_base + TC_PIPE_BYTES + 8 * (2 * TC_MAX_STAGES + 5));
    float* bias_smem = reinterpret_cast<float*>(gen_base + TC_PIPE_BYTES + TC_BAR_BYTES);   // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_stages = a.n_stages;

    if (a.dbg & 16) return;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < n_stages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS * 2); }
        mbar_init(ares_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast
    if (warp == 1) {      // TMEM allocation: all 512 columns (1 CTA / SM)
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // everything above (barriers, tensor-map prefetch, TMEM allocation) overlapped the tail of the previous kernel;
    // from here on global memory written by it is read
    pdl_sync();

    const int bn_cta = a.BN >> 1;                                    // B rows staged by this CTA
    const int pair = (a.dbg & 8) ? (1 << 29) : (int)(blockIdx.x >> 1);     // probe: a pair index that owns no tile
    const int npairs = (int)(gridDim.x >> 1);
    const uint32_t ring_base = smem_base + (uint32_t)a.ring_off;
    const bool resident = a.resident != 0;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0 && !(a.dbg & 1)) {
            int stage = 0;
            uint32_t phase = 0;
            TileWalk w;
            for (bool ok = walk_begin(w, a, pair, npairs); ok; ok = walk_next(w, a)) {
                const int kb0 = w.sp * a.kb_per_split;
                const int kb1 = min(a.kb_total, kb0 + a.kb_per_split);
                const int m0 = (w.m_idx * 2 + (int)cta_rank) * TC_BM;
                const int n0 = w.n_idx * a.BN + (int)cta_rank * bn_cta;
                const bool load_a = !resident || w.pass_first;
                if (resident && w.pass_first && w.pass > 0)
                    mbar_wait(ares_empty, (uint32_t)(w.pass - 1) & 1u);   // the previous slice's MMAs have retired
                const uint32_t tx = (uint32_t)a.b_bytes + (load_a ? (uint32_t)TC_KBLK_BYTES : 0u);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t slot = ring_base + (uint32_t)stage * (uint32_t)a.stage_bytes;
                    const uint32_t sa = resident ? smem_base + (uint32_t)kb * TC_KBLK_BYTES : slot;
                    const uint32_t sb = resident ? slot : slot + TC_KBLK_BYTES;
                    const int k0 = kb * TC_BK;
                    // both CTAs' bytes complete on the leader's full barrier
                    const uint32_t lead_full = mapa_cluster(full_bar(stage), 0u);
                    if (leader) mbar_expect_tx(full_bar(stage), 2u * tx);
                    else        mbar_arrive_cluster(lead_full);
                    if (load_a) {
                        if (!A_MN) {
                            tma_load_2d(sa, &tmA, lead_full, k0, m0);
                        } else {
#pragma unroll
                            for (int c = 0; c < TC_BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, lead_full, m0 + 64 * c, k0);
                        }
                    }
                    if (!B_MN) {
                        tma_load_2d(sb, &tmB, lead_full, k0, n0);
                    } else {
                        for (int c = 0; c * 64 < bn_cta; ++c) tma_load_2d(sb + c * 8192, &tmB, lead_full, n0 + 64 * c, k0);
                    }
                    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && leader) {
        // =============================== MMA issuer (leader CTA only) ===============
        // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b_format F16 [7,10)/[10,13)=0,
        // a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)((TC_BM * 2) >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        TileWalk w;
        for (bool ok = walk_begin(w, a, pair, npairs); ok; ok = walk_next(w, a)) {
            const int kb0 = w.sp * a.kb_per_split;
            const int kb1 = min(a.kb_total, kb0 + a.kb_per_split);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
            for (int kb = kb0; kb < kb1; ++kb) {
                if (!(a.dbg & 1)) mbar_wait(full_bar(stage), phase);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t slot = ring_base + (uint32_t)stage * (uint32_t)a.stage_bytes;
                    const uint32_t sa = resident ? smem_base + (uint32_t)kb * TC_KBLK_BYTES : slot;
                    const uint32_t sb = resident ? slot : slot + TC_KBLK_BYTES;
#pragma unroll
                    for (int j = 0; j < TC_BK / 16; ++j) {
                        const uint64_t ad = A_MN ? make_sdesc(sa + j * 2048, 8192, 1024) : make_sdesc(sa + j * 32, 0, 1024);
                        const uint64_t bd = B_MN ? make_sdesc(sb + j * 2048, 8192, 1024) : make_sdesc(sb + j * 32, 0, 1024);
                        if (a.dbg & 4) continue;
                        tcgen05_mma_f16(d_tmem, ad, bd, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
                    }
                    tcgen05_commit_mc(empty_bar(stage));                  // frees the slot in BOTH CTAs
                    if (kb == kb1 - 1) {
                        tcgen05_commit_mc(tfull_bar(acc));                 // wakes both CTAs' epilogues
                        if (resident && w.pass_last) tcgen05_commit_mc(ares_empty);
                    }
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp >= 2) {
        // =============================== epilogue ===================================
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which of the EPI_SUB warps of that quarter
        const int et = threadIdx.x - 64;
        const int row_in_tile = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t lead_tempty[2] = {mapa_cluster(tempty_bar(0), 0u), mapa_cluster(tempty_bar(1), 0u)};
        float oscale = a.out_scale;
        if (MODE == TC_EPI_STORE && a.out_scale_ptr) oscale *= __ldg(a.out_scale_ptr);
        float run_max = -INFINITY, run_sum = 0.f;
        TileWalk w;
        for (bool ok = walk_begin(w, a, pair, npairs); ok; ok = walk_next(w, a)) {
            const int m = (w.m_idx * 2 + (int)cta_rank) * TC_BM + row_in_tile;
            const int n0 = w.n_idx * a.BN;
            const int n_valid = min(a.BN, a.N - n0);
            const int nch = (n_valid + 31) >> 5;
            // the tile's bias slice travels global -> register while the MMAs are still running
            float bias_reg = 0.f;
            if (a.bias != nullptr && et < n_valid && et < 256) bias_reg = __ldg(a.bias + n0 + et);
            float off_l2 = 0.f;
            if (MODE == TC_EPI_PROB && m < a.M) off_l2 = fmaf(-a.lse[m], LOG2E_F, a.prob_log2_scale);
            if (MODE == TC_EPI_LSE && (!resident || w.pass_first)) { run_max = -INFINITY; run_sum = 0.f; }
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            float* bias_s = bias_smem + acc * 256;
            if (et < 256) bias_s[et] = bias_reg;
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            float va[32], vb[32];
            int c = (a.dbg & 2) ? nch : half;
            if (c < nch) tmem_ld32_issue(taddr + (uint32_t)(c * 32), va);
            for (; c < nch; c += 2 * EPI_SUB) {
                tmem_wait_ld(va);
                if (c + EPI_SUB < nch) tmem_ld32_issue(taddr + (uint32_t)((c + EPI_SUB) * 32), vb);
                epi_chunk<MODE>(a, va, bias_s, c * 32, min(32, n_valid - c * 32), n0, m, w.sp, run_max, run_sum, off_l2, oscale);
                if (c + EPI_SUB < nch) {
                    tmem_wait_ld(vb);
                    if (c + 2 * EPI_SUB < nch) tmem_ld32_issue(taddr + (uint32_t)((c + 2 * EPI_SUB) * 32), va);
                    epi_chunk<MODE>(a, vb, bias_s, (c + EPI_SUB) * 32, min(32, n_valid - (c + EPI_SUB) * 32), n0, m, w.sp, run_max,
                                    run_sum, off_l2, oscale);
                }
            }
            if (MODE == TC_EPI_LSE && m < a.M && !(a.dbg & 2) && (!resident || w.pass_last)) {
                // streaming: one partial per (item tile, sub-warp); resident: one per (pair of the group, sub-warp),
                // accumulated over all the item tiles this pair owns
                const int slot = resident ? w.r : w.n_idx;
                const int64_t pi = (int64_t)(slot * EPI_SUB + half) * a.M + m;
                a.part_max[pi] = run_max;
                a.part_sum[pi] = run_sum;
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lead_tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    // teardown
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp16 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row pitch ld halfs; out-of-bounds
// elements of a box read as zero.  A training step issues the same descriptors every time (same buffers, same
// shapes), and encoding one costs a few microseconds of host time: descriptors are cached by their full key.
struct TmapKey {
    const void* base; int64_t inner, outer, ld; int box_inner, box_outer;
    bool operator==(const TmapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
               box_outer == o.box_outer;
    }
};
static int make_tmap_uncached(CUtensorMap* tm, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                              int box_outer) {
    EncodeTiledFn enc = get_encode();
    B200_REQUIRE(enc, B200VAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2ull};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200VAE_ECUDA,
                 "cuTensorMapEncodeTiled failed (%d): base %p inner %lld outer %lld ld %lld box %dx%d", (int)r, base,
                 (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer);
    return 0;
}
static int make_tmap(CUtensorMap* tm, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                     int box_outer) {
    constexpr int CAP = 32;
    static thread_local TmapKey keys[CAP];
    static thread_local CUtensorMap maps[CAP];
    static thread_local int used = 0, next = 0;
    const TmapKey k = {base, inner, outer, ld, box_inner, box_outer};
    for (int i = 0; i < used; ++i)
        if (keys[i] == k) { *tm = maps[i]; return 0; }
    B200_CHECK(make_tmap_uncached(tm, base, inner, outer, ld, box_inner, box_outer));
    const int slot = used < CAP ? used++ : (next++ % CAP);
    keys[slot] = k;
    maps[slot] = *tm;
    return 0;
}

// operand pitches are in halfs and must keep every row 16-byte aligned (TMA global strides)
bool tc_supported(int M, int N, int K, int64_t lda, int64_t ldb) {
    return M >= 1 && N >= 16 && K >= 1 && (lda % 8 == 0) && (ldb % 8 == 0);
}

// The launch geometry of one GEMM: N tile, tile counts, resident-A decision, ring shape, grid.
struct TcPlan {
    int BN, tiles_m, tiles_n, kb_total, split_k, kb_per_split;
    int resident, n_stages, stage_bytes, ring_off, b_bytes, pairs_per_m, n_groups;
    int lse_parts, grid;
};

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// N tile <= 256, a multiple of 16 (the pair MMA's N granule; each CTA stages BN/2 rows, a multiple of the 8-row
// swizzle atom): fewest tiles, then the smallest tile that covers N
static int pick_bn(int N) {
    const int64_t tiles = cdiv(N, 256);
    return (int)std::min<int64_t>(256, round_up(cdiv(N, tiles), 16));
}
// N tile of a GEMM whose CTAs walk `tiles per pair` item tiles one after the other: all tiles of a launch have the
// same shape, so the launch takes ceil(tiles / parallel) waves of one tile each: pick the tile that minimises
// waves x (BN + fixed cost) instead of always the widest one -- 50 000 items on 74 pairs: 196 x 256 = 3 waves of
// 256, 209 x 240 = 3 waves of 240.
static int pick_bn_items(int N, int64_t tiles_m_serial, int parallel) {
    if (N <= 256) return (int)std::min<int64_t>(256, round_up(N, 16));
    int best = 256;
    int64_t best_cost = INT64_MAX;
    for (int bn = 256; bn >= 128; bn -= 16) {
        const int64_t waves = cdiv(cdiv(N, bn) * tiles_m_serial, parallel);
        const int64_t cost = waves * (bn + 24);        // ~24 columns' worth of per-tile fixed cost (ring refill, epilogue tail)
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

static TcPlan tc_plan(int mode, int M, int N, int K, int a_mn, int b_mn, int split_k_req, int num_sms) {
    TcPlan p = {};
    const int npairs = std::max(1, num_sms / 2);
    p.tiles_m = (int)cdiv(M, 2 * TC_BM);
    p.kb_total = (int)cdiv(K, TC_BK);
    p.split_k = std::max(1, std::min(split_k_req, p.kb_total));
    p.kb_per_split = (int)cdiv(p.kb_total, p.split_k);
    p.split_k = (int)cdiv(p.kb_total, p.kb_per_split);     // no empty splits
    // resident A: K-major A, no split-K, a wide N (several tiles per pair), and the slice + >= 3 ring stages fit
    const int64_t a_res = (int64_t)p.kb_total * TC_KBLK_BYTES;
    const bool want_res = !a_mn && p.split_k == 1 && N >= 1024 && env_int("B200VAE_TC_RESIDENT", 1) != 0;
    int groups = std::min(p.tiles_m, npairs);
    int ppm = std::max(1, npairs / groups);
    const bool items = (mode == TC_EPI_LSE || mode == TC_EPI_PROB || N >= 1024);
    if (want_res) {
        p.BN = pick_bn_items(N, cdiv(p.tiles_m, groups), ppm);
    } else if (N < 1024 && p.split_k == 1) {
        // hidden-layer sized outputs: these GEMMs are latency bound, so spread them over the SM pairs with narrow
        // N tiles instead of filling 256 columns on a handful of pairs
        const int64_t want_tiles_n = std::max<int64_t>(1, npairs / std::max(1, p.tiles_m));
        p.BN = (int)std::min<int64_t>(256, std::max<int64_t>(32, round_up(cdiv(N, want_tiles_n), 16)));
    } else {
        p.BN = items && !b_mn ? pick_bn_items(N, p.tiles_m, npairs) : pick_bn(N);
    }
    {
        const int forced = env_int("B200VAE_TC_BN", 0);
        if (forced >= 32 && forced <= 256 && forced % 16 == 0 && N > 256) p.BN = forced;
    }
    p.tiles_n = (int)cdiv(N, p.BN);
    const int bn_cta = p.BN / 2;
    p.b_bytes = b_mn ? (int)cdiv(bn_cta, 64) * 8192 : bn_cta * 128;
    const int b_slot = (int)round_up(p.b_bytes, 1024);
    p.resident = 0;
    if (want_res && a_res + 3 * (int64_t)b_slot <= TC_PIPE_BYTES) {
        p.resident = 1;
        p.ring_off = (int)a_res;
        p.stage_bytes = b_slot;
        p.n_stages = (int)std::min<int64_t>(TC_MAX_STAGES, (TC_PIPE_BYTES - a_res) / b_slot);
        ppm = std::min(ppm, p.tiles_n);
        p.pairs_per_m = ppm;
        p.n_groups = groups;
        p.grid = 2 * std::min(npairs, groups * ppm);
        p.lse_parts = ppm * (tc_epi_warps(TC_EPI_LSE) / 4);
    } else {
        p.ring_off = 0;
        p.stage_bytes = TC_KBLK_BYTES + b_slot;
        p.n_stages = std::min(TC_MAX_STAGES, TC_PIPE_BYTES / p.stage_bytes);
        p.pairs_per_m = 1;
        p.n_groups = 1;
        const int total = p.tiles_m * p.tiles_n * p.split_k;
        p.grid = 2 * std::min(total, npairs);
        p.lse_parts = p.tiles_n * (tc_epi_warps(TC_EPI_LSE) / 4);
    }
    return p;
}

// number of (max, sum) partial rows per user the LSE epilogue writes for this shape
int tc_lse_parts(int M, int N, int K, int num_sms) { return tc_plan(TC_EPI_LSE, M, N, K, 0, 0, 1, num_sms).lse_parts; }
// upper bound over every shape a context can see (used to size the partial buffers)
int tc_lse_parts_max(int N, int num_sms) { return std::max((int)cdiv(N, 128), num_sms / 2) * (tc_epi_warps(TC_EPI_LSE) / 4); }
int tc_output_tiles(int M, int N) { return (int)(cdiv(M, 2 * TC_BM) * cdiv(N, pick_bn(N))); }
int tc_parallel_tiles(int num_sms) { return std::max(1, num_sms / 2); }

template <int MODE, bool A_MN, bool B_MN>
static int launch_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& args, int grid, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm<MODE, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc_threads(MODE));
    cfg.dynamicSmemBytes = TC_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_gemm<MODE, A_MN, B_MN>, tmA, tmB, args));
    return 0;
}

int launch_tc_gemm(Ctx* c, int mode, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn,
                   void* C, int64_t ldc, int M, int N, int K, const TcEpi& e, cudaStream_t s) {
    B200_REQUIRE(tc_supported(M, N, K, lda, ldb), B200VAE_EINVAL, "tc_gemm: unsupported shape M=%d N=%d K=%d lda=%lld ldb=%lld",
                 M, N, K, (long long)lda, (long long)ldb);
    B200_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, B200VAE_EINVAL, "tc_gemm: operands must be 16-byte aligned");
    const TcPlan p = tc_plan(mode, M, N, K, a_mn, b_mn, e.split_k, c->num_sms);
    TcArgs a = {};
    a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
    a.BN = p.BN; a.tiles_m = p.tiles_m; a.tiles_n = p.tiles_n;
    a.n_fastest = e.n_fastest;
    a.kb_total = p.kb_total; a.split_k = p.split_k; a.kb_per_split = p.kb_per_split;
    a.split_stride = e.split_stride;
    a.bias = e.bias; a.part_max = e.part_max; a.part_sum = e.part_sum; a.lse = e.lse;
    a.bias_grad = e.bias_grad;
    a.row_bias = e.row_bias;
    a.bias_col = e.bias_col;
    a.transpose_out = e.transpose_out;
    a.n_store = (e.bias_col >= 0) ? e.bias_col : (e.transpose_out ? M : N);
    a.out_scale = e.out_scale; a.out_scale_ptr = e.out_scale_ptr; a.prob_log2_scale = e.prob_log2_scale;
    a.act = e.act; a.mulY = e.mulY; a.ldy = e.ldy; a.C16 = e.C16; a.ldc16 = e.ldc16;
    a.accumulate = e.accumulate;
    B200_REQUIRE(!a.accumulate || (mode == TC_EPI_STORE && !e.transpose_out && e.split_k <= 1 && !(a.act || a.mulY || a.C16)),
                 B200VAE_EINVAL, "tc_gemm: accumulate needs a plain STORE output");
    B200_REQUIRE(!(a.act || a.mulY || a.C16) || (mode == TC_EPI_STORE && !e.transpose_out && e.split_k <= 1), B200VAE_EINVAL,
                 "tc_gemm: the activation epilogue needs a plain STORE output");
    a.vec_ok = (mode == TC_EPI_STORE && C && ((uintptr_t)C & 15) == 0 && (ldc % 4 == 0) && (e.split_stride % 4 == 0)) ? 1 : 0;
    a.resident = p.resident; a.n_stages = p.n_stages; a.stage_bytes = p.stage_bytes; a.ring_off = p.ring_off;
    a.b_bytes = p.b_bytes; a.pairs_per_m = p.pairs_per_m; a.n_groups = p.n_groups;
    {
        static int dbg = -1;
        if (dbg < 0) dbg = env_int("B200VAE_TC_DBG", 0);
        a.dbg = dbg;
    }
    if (mode == TC_EPI_LSE)
        B200_REQUIRE(p.lse_parts <= e.part_rows, B200VAE_ECAPACITY, "tc_gemm: %d log-sum-exp partial rows, buffer holds %d",
                     p.lse_parts, e.part_rows);
    if (mode == TC_EPI_STORE && e.split_k > 1 && a.split_k != e.split_k) {
        // the caller sized its reduction for e.split_k partials: zero the ones we will not write
        B200_CUDA_OK(cudaMemsetAsync(reinterpret_cast<float*>(C) + (int64_t)a.split_k * e.split_stride, 0,
                                     (size_t)(e.split_k - a.split_k) * e.split_stride * sizeof(float), s));
    }
    CUtensorMap tmA, tmB;
    if (!a_mn) B200_CHECK(make_tmap(&tmA, A, K, M, lda, TC_BK, TC_BM));
    else       B200_CHECK(make_tmap(&tmA, A, M, K, lda, 64, TC_BK));
    if (!b_mn) B200_CHECK(make_tmap(&tmB, B, K, N, ldb, TC_BK, a.BN / 2));
    else       B200_CHECK(make_tmap(&tmB, B, N, K, ldb, 64, TC_BK));
    int rc;
#define INST(MODE_, AM, BM_) rc = launch_inst<MODE_, AM, BM_>(tmA, tmB, a, p.grid, s)
    if (mode == TC_EPI_LSE) { B200_REQUIRE(!a_mn && !b_mn, B200VAE_EINVAL, "LSE epilogue expects K-major operands"); INST(TC_EPI_LSE, false, false); }
    else if (mode == TC_EPI_PROB) { B200_REQUIRE(!a_mn && !b_mn, B200VAE_EINVAL, "PROB epilogue expects K-major operands"); INST(TC_EPI_PROB, false, false); }
    else if (!a_mn && !b_mn) INST(TC_EPI_STORE, false, false);
    else if (a_mn && !b_mn) INST(TC_EPI_STORE, true, false);
    else if (!a_mn && b_mn) INST(TC_EPI_STORE, false, true);
    else INST(TC_EPI_STORE, true, true);
#undef INST
    note(c, __func__, s);
    return rc;
}

// out[m,n] = (sum_s parts[s][m,n] * rowscale[m] * scale + addend_scale*addend[m,n]) * (1 - Y[m,n]^2)
__global__ void k_splitk_reduce(const float* __restrict__ parts, int n_split, int64_t split_stride, float* __restrict__ out,
                                int64_t ld_out, int M, int N, int64_t ld_part, const float* __restrict__ addend, int64_t ld_add,
                                float addend_scale, const float* __restrict__ mulY, int64_t ldy,
                                const float* __restrict__ rowscale, float scale, __half* __restrict__ out16, int64_t ld16) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    int m = (int)(i / N), n = (int)(i % N);
    float acc = 0.f;
    for (int s = 0; s < n_split; ++s) acc += parts[(int64_t)s * split_stride + (int64_t)m * ld_part + n];
    acc *= scale;
    if (rowscale) acc *= rowscale[m];
    if (addend) acc += addend_scale * addend[(int64_t)m * ld_add + n];
    if (mulY) {
        float t = mulY[(int64_t)m * ldy + n];
        acc *= (1.f - t * t);
    }
    out[(int64_t)m * ld_out + n] = acc;
    if (out16) out16[(int64_t)m * ld16 + n] = __float2half_rn(f16_clamp(acc));
}

int launch_splitk_reduce(Ctx* c, const float* parts, int n_split, int64_t split_stride, float* out, int64_t ld_out, int M,
                         int N, int64_t ld_part, const float* addend, int64_t ld_add, float addend_scale, const float* mulY,
                         int64_t ldy, const float* rowscale, float scale, cudaStream_t s, __half* out16, int64_t ld16) {
    int64_t n = (int64_t)M * N;
    if (n == 0) return 0;
    B200_CUDA_OK(launch_pdl(k_splitk_reduce, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, s, parts, n_split, split_stride, out,
                            ld_out, M, N, ld_part, addend, ld_add, addend_scale, mulY, ldy, rowscale, scale, out16, ld16));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
