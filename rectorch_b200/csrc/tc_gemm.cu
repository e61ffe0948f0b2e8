// tc_gemm.cu -- the item-sized contractions of the decoder output layer on the 5th-gen
// tensor cores: tcgen05.mma (kind::tf32, fp32 accumulate in TMEM), operands staged in shared
// memory by TMA (128 B swizzle), warp-specialised persistent CTAs, one CTA per SM.
//
//   D[M x N] = A[M x K] * B[N x K]^T
//
//   TC_EPI_LSE   K4  logits = h W_d^T + b never leave the SM: each epilogue thread owns one user
//                    row (one TMEM lane) and folds its share of the 256-item tile into an online
//                    (max, sum exp); output = per-(item half-tile, user) partials, merged by
//                    k_lse_merge.            (F.log_softmax over nets.py:417's output, models.py:813)
//   TC_EPI_PROB  K5  same mainloop, epilogue recomputes softmax from the saved lse and stores
//                    P^T[item, user] = exp(logit - lse_u) * T_u/B  (coalesced: a warp's 32 lanes
//                    are 32 consecutive users)                      (dlogits of loss.backward())
//   TC_EPI_STORE     plain product (+ bias), optional split-K partials and a "bias column":
//                    dW_d|db_d = P^T [h | 1]  and  dh = P W_d.
//
// Operand "majorness": K-major = the contraction index is contiguous in memory (TMA box
// 32 floats of K x rows, SWIZZLE_128B), MN-major = the M/N index is contiguous (box 32 floats of
// M/N x 32 K-rows, one box per 32-wide chunk, SWIZZLE_128B with 32 B atoms -- the only MN-major
// layout the tensor core takes for tf32).  The instruction descriptor's a_major/b_major bits
// select the interpretation.
//
// Pipelines (all mbarrier based, no __syncthreads in the steady state):
//   warp 0    TMA producer     : empty[s] -> issue loads -> full[s] (complete_tx)
//   warp 1    MMA issuer       : full[s] -> 4 x tcgen05.mma (K=8 each) -> commit -> empty[s];
//                                after the last K block: commit -> tmem_full[a]
//   warps 2-9 epilogue (8)     : tmem_full[a] -> tcgen05.ld 32 columns at a time, software
//                                pipelined (the load of chunk c+1 is in flight while chunk c is
//                                reduced) -> math -> global stores -> tmem_empty[a].
//                                Two warps share each TMEM lane quarter and interleave the column
//                                chunks, so every SM sub-partition has two epilogue warps to issue from.
// TMEM: 512 columns = 2 accumulator stages x 256 columns, so the epilogue of tile i overlaps
// the MMAs of tile i+1.
//
// CTA pairs (template CG2, the default): the kernel is launched in clusters of two CTAs that sit
// on the two SMs of a TPC and execute ONE tcgen05.mma.cta_group::2 of shape M=256 x N=256: each
// CTA stages its own 128 rows of A and only HALF of the B tile (its 128 of the 256 N rows); the
// tensor cores of both SMs read both halves.  Per unit of math this cuts the bytes every SM pulls
// through the L2 crossbar by a third -- these tf32 GEMMs run at the chip-wide L2->SM throughput
// cap (~6.5 KB/clk, profiles/), not at the tensor or HBM roofline, so that is what buys time.
//   - the peer's TMA completes its bytes on the LEADER's full barrier (leader expects 2x bytes),
//   - only the leader's warp 1 issues MMAs; tcgen05.commit multicasts the "stage free" /
//     "accumulator ready" arrivals to both CTAs,
//   - both CTAs' epilogue warps arrive on the leader's tmem_empty barrier.
#include <cuda.h>
#include <algorithm>
#include "ctx.cuh"

namespace b200 {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // floats per K block = 128 B = one swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_PIPE_BYTES = 192 * 1024;       // operand ring: 4 x 48 KB (1 CTA) or 6 x 32 KB (CTA pair)
template <bool CG2> struct TcCfg {
    static constexpr int B_BYTES = (CG2 ? 128 : 256) * TC_BK * 4;   // this CTA's share of the max N tile
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int STAGES = TC_PIPE_BYTES / STAGE_BYTES;
};
// Epilogue warps per CTA: 2 per TMEM lane quarter for the register-hungry log-sum-exp epilogue (measured:
// 4 per quarter caps it at 96 registers and costs 15 %), 4 per quarter for the store-heavy epilogues
// (softmax / plain store: more stores in flight, -9 %).  Warps sharing a quarter interleave the 32-column chunks.
__host__ __device__ constexpr int tc_epi_warps(int mode) { return mode == TC_EPI_LSE ? 8 : 16; }
__host__ __device__ constexpr int tc_threads(int mode) { return 64 + 32 * tc_epi_warps(mode); }
constexpr int TC_BAR_BYTES = 256;
constexpr int TC_BIAS_BYTES = 2 * 256 * 4;
constexpr int TC_SMEM = TC_PIPE_BYTES + 1024 /*align*/ + TC_BAR_BYTES + TC_BIAS_BYTES;
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// ---- cluster / CTA-pair variants ----
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
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_cg2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                     uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
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
// start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1 layout_type[61,64)).
//   K-major  operands: layout_type 2 = SWIZZLE_128B (8 rows x 128 B atoms, 16 B swizzle granules),
//                      SBO = 1024 (stride between 8-row groups), LBO unused.
//   MN-major tf32 operands: the ONLY layout the tensor core accepts is layout_type 1 =
//                      SWIZZLE_128B_BASE32B (4 K-rows x 128 B atoms, 32 B swizzle granules;
//                      TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): SBO = 512 (stride between
//                      4-K-row groups), LBO = stride between 32-float MN chunks.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
struct TcArgs {
    float* C;
    int64_t ldc;
    int M, N, K;
    int BN;                 // N tile (multiple of 16, <= 256)
    int tiles_m, tiles_n;
    int n_fastest;          // tile order: consecutive CTAs walk N first (share the A tile through L2)
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
    const float* rowscale;
    float* bias_grad;
    int vec_ok;             // C rows are 16 B aligned
    int dbg;                // B200VAE_TC_DBG bit mask (probes, garbage results): 1 = no operand loads,
                            // 2 = epilogue only waits and releases the accumulator, 4 = no MMAs are issued
};

struct TileCoord { int m_idx, n_idx, sp; };
__device__ __forceinline__ TileCoord decode_tile(int t, const TcArgs& a) {
    TileCoord c;
    const int mn = a.tiles_m * a.tiles_n;
    c.sp = t / mn;
    const int r = t - c.sp * mn;
    if (a.n_fastest) { c.n_idx = r % a.tiles_n; c.m_idx = r / a.tiles_n; }
    else             { c.m_idx = r % a.tiles_m; c.n_idx = r / a.tiles_m; }
    return c;
}

constexpr float LOG2E_F = 1.4426950408889634f;

// one 32-column chunk of the accumulator for the thread's row
template <int MODE>
__device__ __forceinline__ void epi_chunk(const TcArgs& a, const float* v, const float* __restrict__ bias_s, int c0, int nc,
                                          int n0, int m, int sp, float& run_max, float& run_sum, float lse_l2, float rs_m) {
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
        // P^T[(n0+c0+i) * ldc + m]: for a fixed i the warp's 32 lanes write 32 consecutive floats
        if (m < a.M) {
            float* dst = a.C + (int64_t)(n0 + c0) * a.ldc + m;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (i < nc) {
                    const float x = v[i] + bias_s[c0 + i];
                    dst[(int64_t)i * a.ldc] = tf32_rn(ex2_approx(fmaf(x, LOG2E_F, -lse_l2)) * rs_m);
                }
            }
        }
    } else if (a.transpose_out) {
        // D^T: for a fixed column the warp's 32 lanes (consecutive m) write 32 consecutive floats
        if (m < a.M) {
            if (m < a.n_store) {
                float* dst = a.C + (int64_t)(n0 + c0) * a.ldc + m;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nc) dst[(int64_t)i * a.ldc] = v[i];
            } else if (m == a.bias_col) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < nc) a.bias_grad[n0 + c0 + i] = v[i];
            }
        }
    } else {
        if (m < a.M) {
            float* crow = a.C + (int64_t)sp * a.split_stride + (int64_t)m * a.ldc;
            const int col0 = n0 + c0;
            if (a.vec_ok && col0 + 32 <= a.n_store) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + i);
                    *reinterpret_cast<float4*>(crow + col0 + i) =
                        make_float4(v[i] + b4.x, v[i + 1] + b4.y, v[i + 2] + b4.z, v[i + 3] + b4.w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int col = col0 + i;
                    if (i < nc) {
                        if (col < a.n_store) crow[col] = v[i] + bias_s[c0 + i];
                        else if (col == a.bias_col) a.bias_grad[m] = v[i];
                    }
                }
            }
        }
    }
}

template <int MODE, bool A_MN, bool B_MN, bool CG2>
__global__ void __launch_bounds__(tc_threads(MODE), 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    constexpr int TC_EPI_WARPS = tc_epi_warps(MODE);
    constexpr int TC_EPI_SUB = TC_EPI_WARPS / 4;
    constexpr int TC_EPI_THREADS = 32 * TC_EPI_WARPS;
    constexpr int TC_STAGES = TcCfg<CG2>::STAGES;
    constexpr int TC_STAGE_BYTES = TcCfg<CG2>::STAGE_BYTES;
    constexpr int NCTA = CG2 ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + TC_PIPE_BYTES;
    const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;
    const bool leader = (cta_rank == 0);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 4);
    uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gen_base + TC_PIPE_BYTES + 8 * (2 * TC_STAGES + 4));
    float* bias_smem = reinterpret_cast<float*>(gen_base + TC_PIPE_BYTES + TC_BAR_BYTES);   // [2][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), NCTA); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), TC_EPI_WARPS * NCTA); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (CG2) cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast
    if (warp == 1) {   // TMEM allocation: all 512 columns (1 CTA / SM)
        if (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if (CG2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int total_tiles = a.tiles_m * a.tiles_n * a.split_k;      // CG2: tiles_m counts 256-row pair tiles
    const int bn_cta = CG2 ? (a.BN >> 1) : a.BN;                     // B rows staged by this CTA
    const uint32_t b_bytes = (uint32_t)bn_cta * TC_BK * 4u;
    const int tile0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0 && !(a.dbg & 1)) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = tile0; t < total_tiles; t += tile_step) {
                const TileCoord tc = decode_tile(t, a);
                const int kb0 = tc.sp * a.kb_per_split;
                const int kb1 = min(a.kb_total, kb0 + a.kb_per_split);
                const int m0 = (tc.m_idx * NCTA + (int)cta_rank) * TC_BM;
                const int n0 = tc.n_idx * a.BN + (int)cta_rank * bn_cta;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
                    const uint32_t sb = sa + TC_A_BYTES;
                    const int k0 = kb * TC_BK;
                    if (!CG2) {
                        mbar_expect_tx(full_bar(stage), TC_A_BYTES + b_bytes);
                        if (!A_MN) {
                            tma_load_2d(sa, &tmA, full_bar(stage), k0, m0);
                        } else {
#pragma unroll
                            for (int c = 0; c < TC_BM / 32; ++c) tma_load_2d(sa + c * 4096, &tmA, full_bar(stage), m0 + 32 * c, k0);
                        }
                        if (!B_MN) {
                            tma_load_2d(sb, &tmB, full_bar(stage), k0, n0);
                        } else {
                            for (int c = 0; c < bn_cta / 32; ++c) tma_load_2d(sb + c * 4096, &tmB, full_bar(stage), n0 + 32 * c, k0);
                        }
                    } else {
                        // both CTAs' bytes complete on the leader's full barrier
                        const uint32_t lead_full = mapa_cluster(full_bar(stage), 0u);
                        if (leader) mbar_expect_tx(full_bar(stage), 2u * (TC_A_BYTES + b_bytes));
                        else        mbar_arrive_cluster(lead_full);
                        if (!A_MN) {
                            tma_load_2d_cg2(sa, &tmA, lead_full, k0, m0);
                        } else {
#pragma unroll
                            for (int c = 0; c < TC_BM / 32; ++c) tma_load_2d_cg2(sa + c * 4096, &tmA, lead_full, m0 + 32 * c, k0);
                        }
                        if (!B_MN) {
                            tma_load_2d_cg2(sb, &tmB, lead_full, k0, n0);
                        } else {
                            for (int c = 0; c < bn_cta / 32; ++c) tma_load_2d_cg2(sb + c * 4096, &tmB, lead_full, n0 + 32 * c, k0);
                        }
                    }
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && leader) {
        // =============================== MMA issuer (leader CTA only) ===============
        // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1,
        // a/b_format TF32 [7,10)/[10,13)=2, a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(a.BN >> 3) << 17) | ((uint32_t)((TC_BM * NCTA) >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = tile0; t < total_tiles; t += tile_step) {
            const TileCoord tc = decode_tile(t, a);
            const int kb0 = tc.sp * a.kb_per_split;
            const int kb1 = min(a.kb_total, kb0 + a.kb_per_split);
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
            for (int kb = kb0; kb < kb1; ++kb) {
                if (!(a.dbg & 1)) mbar_wait(full_bar(stage), phase);
                tcgen05_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
                    const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
                    for (int j = 0; j < TC_BK / 8; ++j) {
                        const uint64_t ad = A_MN ? make_sdesc(sa + j * 1024, 4096, 512, 1) : make_sdesc(sa + j * 32, 0, 1024, 2);
                        const uint64_t bd = B_MN ? make_sdesc(sb + j * 1024, 4096, 512, 1) : make_sdesc(sb + j * 32, 0, 1024, 2);
                        if (a.dbg & 4) continue;
                        if (CG2) tcgen05_mma_tf32_cg2(d_tmem, ad, bd, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
                        else     tcgen05_mma_tf32(d_tmem, ad, bd, idesc, (kb > kb0 || j > 0) ? 1u : 0u);
                    }
                    if (CG2) {
                        tcgen05_commit_cg2(empty_bar(stage));                  // frees the slot in BOTH CTAs
                        if (kb == kb1 - 1) tcgen05_commit_cg2(tfull_bar(acc)); // wakes both CTAs' epilogues
                    } else {
                        tcgen05_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
                        if (kb == kb1 - 1) tcgen05_commit(tfull_bar(acc)); // accumulator complete
                    }
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp >= 2) {
        // =============================== epilogue ===================================
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which of the TC_EPI_SUB warps of that quarter
        const int et = threadIdx.x - 64;                // 0..255
        const int row_in_tile = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t lead_tempty[2] = {CG2 ? mapa_cluster(tempty_bar(0), 0u) : 0u, CG2 ? mapa_cluster(tempty_bar(1), 0u) : 0u};
        for (int t = tile0; t < total_tiles; t += tile_step) {
            const TileCoord tc = decode_tile(t, a);
            const int m = (tc.m_idx * NCTA + (int)cta_rank) * TC_BM + row_in_tile;
            const int n0 = tc.n_idx * a.BN;
            const int n_valid = min(a.BN, a.N - n0);
            const int nch = (n_valid + 31) >> 5;
            // the tile's bias slice travels global -> register while the MMAs are still running
            float bias_reg = 0.f;
            if (a.bias != nullptr && et < n_valid && et < 256) bias_reg = __ldg(a.bias + n0 + et);
            float lse_l2 = 0.f, rs_m = 0.f;
            if (MODE == TC_EPI_PROB && m < a.M) { lse_l2 = a.lse[m] * LOG2E_F; rs_m = a.rowscale[m]; }
            mbar_wait(tfull_bar(acc), acc_phase);
            tcgen05_fence_after();
            float* bias_s = bias_smem + acc * 256;
            if (et < 256) bias_s[et] = bias_reg;
            asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            float run_max = -INFINITY, run_sum = 0.f;
            float va[32], vb[32];
            int c = (a.dbg & 2) ? nch : half;
            if (c < nch) tmem_ld32_issue(taddr + (uint32_t)(c * 32), va);
            for (; c < nch; c += 2 * TC_EPI_SUB) {
                tmem_wait_ld(va);
                if (c + TC_EPI_SUB < nch) tmem_ld32_issue(taddr + (uint32_t)((c + TC_EPI_SUB) * 32), vb);
                epi_chunk<MODE>(a, va, bias_s, c * 32, min(32, n_valid - c * 32), n0, m, tc.sp, run_max, run_sum, lse_l2, rs_m);
                if (c + TC_EPI_SUB < nch) {
                    tmem_wait_ld(vb);
                    if (c + 2 * TC_EPI_SUB < nch) tmem_ld32_issue(taddr + (uint32_t)((c + 2 * TC_EPI_SUB) * 32), va);
                    epi_chunk<MODE>(a, vb, bias_s, (c + TC_EPI_SUB) * 32, min(32, n_valid - (c + TC_EPI_SUB) * 32), n0, m, tc.sp, run_max,
                                    run_sum, lse_l2, rs_m);
                }
            }
            if (MODE == TC_EPI_LSE && m < a.M && !(a.dbg & 2)) {
                const int64_t pi = (int64_t)(tc.n_idx * TC_EPI_SUB + half) * a.M + m;
                a.part_max[pi] = run_max;
                a.part_sum[pi] = run_sum;
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG2) mbar_arrive_cluster(lead_tempty[acc]);
                else     mbar_arrive(tempty_bar(acc));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    // teardown
    tcgen05_fence_before();
    if (CG2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else     asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

// 2-D fp32 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row pitch ld floats.
// A training step issues the same eight descriptors every time (same buffers, same shapes), and encoding one
// costs a few microseconds of host time on a path that is host-bound when the caller synchronises per step:
// descriptors are cached by their full key.
struct TmapKey {
    const float* base; int64_t inner, outer, ld; int box_inner, box_outer; bool mn_major;
    bool operator==(const TmapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
               box_outer == o.box_outer && mn_major == o.mn_major;
    }
};
static int make_tmap_uncached(CUtensorMap* tm, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                              int box_outer, bool mn_major);
static int make_tmap(CUtensorMap* tm, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                     int box_outer, bool mn_major) {
    constexpr int CAP = 32;
    static thread_local TmapKey keys[CAP];
    static thread_local CUtensorMap maps[CAP];
    static thread_local int used = 0, next = 0;
    const TmapKey k = {base, inner, outer, ld, box_inner, box_outer, mn_major};
    for (int i = 0; i < used; ++i)
        if (keys[i] == k) { *tm = maps[i]; return 0; }
    B200_CHECK(make_tmap_uncached(tm, base, inner, outer, ld, box_inner, box_outer, mn_major));
    const int slot = used < CAP ? used++ : (next++ % CAP);
    keys[slot] = k;
    maps[slot] = *tm;
    return 0;
}
static int make_tmap_uncached(CUtensorMap* tm, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner,
                              int box_outer, bool mn_major) {
    EncodeTiledFn enc = get_encode();
    B200_REQUIRE(enc, B200VAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4ull};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200VAE_ECUDA,
                 "cuTensorMapEncodeTiled failed (%d): base %p inner %lld outer %lld ld %lld box %dx%d", (int)r, base,
                 (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer);
    return 0;
}

bool tc_supported(int M, int N, int K, int64_t lda, int64_t ldb) {
    return M >= 1 && N >= 16 && K >= 8 && (lda % 4 == 0) && (ldb % 4 == 0);
}

static bool use_cg2() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200VAE_TC_1CTA");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}
// N tile: <= 256; each CTA of a pair stages BN/2 rows, so the granule doubles in pair mode
static int pick_bn(int N, bool b_mn) {
    const int g = (b_mn ? 32 : 16) * (use_cg2() ? 2 : 1);
    const int64_t tiles = cdiv(N, 256);                 // fewest tiles, then the smallest tile that covers N
    return (int)std::min<int64_t>(256, round_up(cdiv(N, tiles), g));
}
// N tile of the item-sized K-major GEMMs with a per-row epilogue (LSE / PROB: N = n_items, no split-K).  All tiles
// of a launch have the same shape and the persistent CTAs (pairs) walk them round-robin, so the launch takes
// ceil(tiles / parallel) waves of one tile each: pick the tile that minimises waves x (BN + fixed cost) instead of
// always the widest one -- 50 000 items on 74 CTA pairs: 196 x 256 = 3 waves of 256, 209 x 240 = 3 waves of 240.
// A pair tile needs N % 16 == 0 (each CTA stages N/2 rows, a multiple of the 8-row swizzle atom).
static int pick_bn_items(int M, int N, int num_sms) {
    const bool cg2 = use_cg2();
    const int g = 16;
    if (N <= 256) return (int)std::min<int64_t>(256, round_up(N, cg2 ? 32 : 16));
    const int64_t tiles_m = cdiv(M, cg2 ? 2 * TC_BM : TC_BM);
    const int64_t parallel = cg2 ? std::max(1, num_sms / 2) : num_sms;
    int best = 256;
    int64_t best_cost = INT64_MAX;
    for (int bn = 256; bn >= 128; bn -= g) {
        const int64_t waves = cdiv(cdiv(N, bn) * tiles_m, parallel);
        const int64_t cost = waves * (bn + 24);        // ~24 columns' worth of per-tile fixed cost (ring refill, epilogue tail)
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}
// number of (max, sum) partial rows the LSE epilogue writes per user: two warps per item tile
int tc_lse_tiles(int M, int N, int num_sms) { return (tc_epi_warps(TC_EPI_LSE) / 4) * (int)cdiv(N, pick_bn_items(M, N, num_sms)); }

int tc_output_tiles(int M, int N, int b_mn) {
    return (int)(cdiv(M, use_cg2() ? 2 * TC_BM : TC_BM) * cdiv(N, pick_bn(N, b_mn != 0)));
}
int tc_parallel_tiles(int num_sms) { return use_cg2() ? std::max(1, num_sms / 2) : num_sms; }

template <int MODE, bool A_MN, bool B_MN, bool CG2>
static int launch_inst2(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& args, int grid, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        B200_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm<MODE, A_MN, B_MN, CG2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc_threads(MODE));
    cfg.dynamicSmemBytes = TC_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG2 ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_gemm<MODE, A_MN, B_MN, CG2>, tmA, tmB, args));
    return 0;
}
template <int MODE, bool A_MN, bool B_MN>
static int launch_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& args, int grid, cudaStream_t s) {
    return use_cg2() ? launch_inst2<MODE, A_MN, B_MN, true>(tmA, tmB, args, grid, s)
                     : launch_inst2<MODE, A_MN, B_MN, false>(tmA, tmB, args, grid, s);
}

int launch_tc_gemm(Ctx* c, int mode, const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn,
                   float* C, int64_t ldc, int M, int N, int K, const TcEpi& e, cudaStream_t s) {
    B200_REQUIRE(tc_supported(M, N, K, lda, ldb), B200VAE_EINVAL, "tc_gemm: unsupported shape M=%d N=%d K=%d lda=%lld ldb=%lld",
                 M, N, K, (long long)lda, (long long)ldb);
    B200_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, B200VAE_EINVAL, "tc_gemm: operands must be 16-byte aligned");
    TcArgs a;
    a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K;
    a.BN = (mode == TC_EPI_LSE || mode == TC_EPI_PROB) ? pick_bn_items(M, N, c->num_sms) : pick_bn(N, b_mn != 0);
    const bool cg2 = use_cg2();
    a.tiles_m = (int)cdiv(M, cg2 ? 2 * TC_BM : TC_BM);
    a.tiles_n = (int)cdiv(N, a.BN);
    a.n_fastest = e.n_fastest;
    a.kb_total = (int)cdiv(K, TC_BK);
    a.split_k = std::max(1, std::min(e.split_k, a.kb_total));
    a.kb_per_split = (int)cdiv(a.kb_total, a.split_k);
    a.split_k = (int)cdiv(a.kb_total, a.kb_per_split);     // no empty splits
    a.split_stride = e.split_stride;
    a.bias = e.bias; a.part_max = e.part_max; a.part_sum = e.part_sum; a.lse = e.lse; a.rowscale = e.rowscale;
    a.bias_grad = e.bias_grad;
    a.bias_col = e.bias_col;
    a.transpose_out = e.transpose_out;
    a.n_store = (e.bias_col >= 0) ? e.bias_col : (e.transpose_out ? M : N);
    a.vec_ok = (C && ((uintptr_t)C & 15) == 0 && (ldc % 4 == 0) && (e.split_stride % 4 == 0)) ? 1 : 0;
    {
        static int dbg = -1;
        if (dbg < 0) { const char* ev = getenv("B200VAE_TC_DBG"); dbg = ev ? atoi(ev) : 0; }
        a.dbg = dbg;
    }
    if (mode == TC_EPI_STORE && e.split_k > 1 && a.split_k != e.split_k) {
        // the caller sized its reduction for e.split_k partials: zero the ones we will not write
        B200_CUDA_OK(cudaMemsetAsync(C + (int64_t)a.split_k * e.split_stride, 0,
                                     (size_t)(e.split_k - a.split_k) * e.split_stride * sizeof(float), s));
    }
    CUtensorMap tmA, tmB;
    if (!a_mn) B200_CHECK(make_tmap(&tmA, A, K, M, lda, TC_BK, TC_BM, false));
    else       B200_CHECK(make_tmap(&tmA, A, M, K, lda, 32, TC_BK, true));
    if (!b_mn) B200_CHECK(make_tmap(&tmB, B, K, N, ldb, TC_BK, cg2 ? a.BN / 2 : a.BN, false));
    else       B200_CHECK(make_tmap(&tmB, B, N, K, ldb, 32, TC_BK, true));
    const int total = a.tiles_m * a.tiles_n * a.split_k;
    const int grid = cg2 ? 2 * std::min(total, std::max(1, c->num_sms / 2)) : std::min(total, c->num_sms);
    int rc;
#define INST(MODE_, AM, BM_) rc = launch_inst<MODE_, AM, BM_>(tmA, tmB, a, grid, s)
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

// out[m,n] = (sum_s parts[s][m,n] + addend_scale*addend[m,n]) * (1 - Y[m,n]^2)
__global__ void k_splitk_reduce(const float* __restrict__ parts, int n_split, int64_t split_stride, float* __restrict__ out,
                                int64_t ld_out, int M, int N, int64_t ld_part, const float* __restrict__ addend, int64_t ld_add,
                                float addend_scale, const float* __restrict__ mulY, int64_t ldy) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    int m = (int)(i / N), n = (int)(i % N);
    float acc = 0.f;
    for (int s = 0; s < n_split; ++s) acc += parts[(int64_t)s * split_stride + (int64_t)m * ld_part + n];
    if (addend) acc += addend_scale * addend[(int64_t)m * ld_add + n];
    if (mulY) {
        float t = mulY[(int64_t)m * ldy + n];
        acc *= (1.f - t * t);
    }
    out[(int64_t)m * ld_out + n] = acc;
}

int launch_splitk_reduce(Ctx* c, const float* parts, int n_split, int64_t split_stride, float* out, int64_t ld_out, int M,
                         int N, int64_t ld_part, const float* addend, int64_t ld_add, float addend_scale, const float* mulY,
                         int64_t ldy, cudaStream_t s) {
    int64_t n = (int64_t)M * N;
    if (n == 0) return 0;
    k_splitk_reduce<<<(int)cdiv(n, 256), 256, 0, s>>>(parts, n_split, split_stride, out, ld_out, M, N, ld_part, addend, ld_add,
                                                       addend_scale, mulY, ldy);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
