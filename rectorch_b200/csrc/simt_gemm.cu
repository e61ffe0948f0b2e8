// simt_gemm.cu -- fp32 CUDA-core GEMM with fused epilogues for the SMALL dense layers
// (hidden x hidden / hidden x latent Linear layers, nets.py:398-404, 413-416, and their
// backward), which are <2% of the step's FLOPs and need exact fp32 for loss parity.
// The item-sized contractions go through tc_gemm.cu (tcgen05) when shapes allow; this
// kernel also serves them for shapes TMA cannot address (rows not 16 B aligned, e.g.
// the 2-item nets of the reference's unit tests).
//
//   C[m,n] = epi( alpha * sum_k A(m,k) * B(k,n) )
//   A(m,k) = A[m*a_rs + k*a_cs],  B(k,n) = B[k*b_rs + n*b_cs]   (any strides -> any transposes)
//
// Tiles are 64 x 64 x 16 (256 threads, 4x4 per thread) or 32 x 32 x 16 (128 threads, 2x4); the
// smaller tile is picked when the 64-row grid would leave most of the 148 SMs idle (the hidden layers are only
// ~500 x 600).  K blocks stream global -> shared through a 3-stage cp.async pipeline, so two L2
// round trips are always in flight behind the FMAs of the current block.
#include "ctx.cuh"

namespace b200 {

constexpr int BK = 16;

template <int MODE, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_simt_gemm(const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ B,
            int64_t b_rs, int64_t b_cs, float* __restrict__ C, int64_t ldc, int M, int N, int K,
            GemmEpi e) {
    constexpr int TXN = BN / TN;                    // threads across N
    constexpr int THREADS = (BM / TM) * TXN;
    constexpr int A_PER = BM * BK / THREADS;
    constexpr int B_PER = BN * BK / THREADS;
    static_assert(BM * BK % THREADS == 0 && BN * BK % THREADS == 0, "tile / thread mismatch");
    constexpr int NST = 3;                          // cp.async pipeline depth
    __shared__ float As[NST][BK][BM + 4];
    __shared__ float Bs[NST][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const bool a_kfast = (a_cs == 1);
    const bool b_nfast = (b_cs == 1);

    // one K block: every thread copies its 4-byte elements global -> shared asynchronously
    // (src-size 0 zero-fills out-of-range elements)
    auto issue_tiles = [&](int k0, int st) {
#pragma unroll
        for (int u = 0; u < A_PER; ++u) {
            const int i = tid + u * THREADS;
            int mm, kk;
            if (a_kfast) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
            const int gm = m0 + mm, gk = k0 + kk;
            const bool ok = (gm < M && gk < K);
            const float* src = ok ? A + (int64_t)gm * a_rs + (int64_t)gk * a_cs : A;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&As[st][kk][mm]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
#pragma unroll
        for (int u = 0; u < B_PER; ++u) {
            const int i = tid + u * THREADS;
            int nn, kk;
            if (b_nfast) { nn = i % BN; kk = i / BN; } else { kk = i % BK; nn = i / BK; }
            const int gn = n0 + nn, gk = k0 + kk;
            const bool ok = (gn < N && gk < K);
            const float* src = ok ? B + (int64_t)gk * b_rs + (int64_t)gn * b_cs : B;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&Bs[st][kk][nn]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int nkb = (K + BK - 1) / BK;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
        if (p < nkb) issue_tiles(p * BK, p);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < nkb; ++kb) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");   // K block kb has landed
        __syncthreads();
        // refill the stage that was consumed in the previous iteration
        if (kb + NST - 1 < nkb) issue_tiles((kb + NST - 1) * BK, (kb + NST - 1) % NST);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const int st = kb % NST;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[st][kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[st][kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

    if (MODE == EPI_STORE) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int gm = m0 + ty * TM + i;
            if (gm >= M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int gn = n0 + tx * TN + j;
                if (gn >= N) continue;
                float y = e.alpha * acc[i][j];
                if (e.bias) y += e.bias[gn];
                if (e.addend) y += e.addend_scale * e.addend[(int64_t)gm * e.ld_addend + gn];
                if (e.act) y = tanhf(y);
                if (e.mulY) {
                    float t = e.mulY[(int64_t)gm * e.ldy + gn];
                    y *= (1.f - t * t);
                }
                C[(int64_t)gm * ldc + gn] = y;
            }
        }
    } else if (MODE == EPI_LSE) {
        static_assert(MODE != EPI_LSE || (TXN == 16 && BN == 64), "LSE epilogue is written for 64-column tiles");
        // per (row, 64-column tile): running max and sum of exp(logit - max)
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int gm = m0 + ty * TM + i;
            float mx = -INFINITY;
            float v[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int gn = n0 + tx * TN + j;
                v[j] = (gn < N) ? acc[i][j] + (e.bias ? e.bias[gn] : 0.f) : -INFINITY;
                mx = fmaxf(mx, v[j]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) sacc += (v[j] == -INFINITY) ? 0.f : expf(v[j] - mx);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
            if (tx == 0 && gm < M) {
                e.part_max[(int64_t)blockIdx.x * M + gm] = mx;
                e.part_sum[(int64_t)blockIdx.x * M + gm] = sacc;
            }
        }
    } else {   // EPI_PROB: P = exp(logit - lse[m]) * rowscale[m]
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int gm = m0 + ty * TM + i;
            if (gm >= M) continue;
            float l = e.lse[gm], rs = e.rowscale[gm];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int gn = n0 + tx * TN + j;
                if (gn >= N) continue;
                float y = acc[i][j] + (e.bias ? e.bias[gn] : 0.f);
                C[(int64_t)gm * ldc + gn] = expf(y - l) * rs;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Vectorised variant for the aligned case (all hidden layers of real configurations):
// 32 x 64 x 16 tiles, 128 threads, 4x4 per thread.  Each operand is staged in the shared-memory
// orientation that matches its contiguous global dimension, so every global->shared copy is a
// 16-byte cp.async (8x fewer LDGSTS than the scalar kernel) and every fragment read is one
// LDS.128 per 4 K-steps.  Measured on B200: the scalar kernel was bound by LDGSTS/LDS issue
// (~28 us for 500x400x600), not by FMAs.
// ------------------------------------------------------------------------------------------
//
// Intra-CTA split-K: the hidden layers give only ~400 warps of 4x4 micro-tiles for 148 SMs (one
// warp per scheduler -> every LDS->FFMA dependency is exposed; ncu: 23 % issue-active, stalled on
// the short scoreboard).  KSPLIT groups of 4 warps each run the pipeline on a quarter of K for the
// SAME output tile with their own operand ring and named barrier, then reduce through shared
// memory in a fixed order (deterministic), giving every scheduler 4 warps to switch between.
constexpr int V_KSPLIT = 4;
template <bool A_KFAST, bool B_KFAST>
struct VSmem {
    float As[3][A_KFAST ? 32 : BK][(A_KFAST ? BK : 32) + 4];
    float Bs[3][B_KFAST ? 64 : BK][(B_KFAST ? BK : 64) + 4];
};

template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(128 * V_KSPLIT)
k_simt_gemm_v(const float* __restrict__ A, int64_t a_ld, const float* __restrict__ B, int64_t b_ld,
              float* __restrict__ C, int64_t ldc, int M, int N, int K, GemmEpi e) {
    constexpr int VBM = 32, VBN = 64, NST = 3;
    pdl_sync();
    // A: [m][k] if K is contiguous (a_ld = row pitch of m) else [k][m] (a_ld = row pitch of k); same for B with n
    extern __shared__ __align__(16) uint8_t vsmem_raw[];
    const int grp = threadIdx.x >> 7;                 // K-split group
    VSmem<A_KFAST, B_KFAST>& sm = reinterpret_cast<VSmem<A_KFAST, B_KFAST>*>(vsmem_raw)[grp];
    auto& As = sm.As;
    auto& Bs = sm.Bs;
    const int tid = threadIdx.x & 127;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * VBM, n0 = blockIdx.x * VBN;
    // this group's K range, in whole K blocks
    const int nkb_all = (K + BK - 1) / BK;
    const int kb_per = (nkb_all + V_KSPLIT - 1) / V_KSPLIT;
    const int kb_lo = min(nkb_all, grp * kb_per), kb_hi = min(nkb_all, kb_lo + kb_per);
    auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    auto cp16 = [](void* dst, const float* src, int bytes) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
    };
    auto issue = [&](int k0, int st) {
        {   // A: 128 chunks of 4 floats, one per thread
            if (A_KFAST) {
                const int row = tid >> 2, kq = (tid & 3) * 4;
                const int gm = m0 + row, gk = k0 + kq;
                const int bytes = (gm < M) ? max(0, min(16, (K - gk) * 4)) : 0;
                cp16(&As[st][row][kq], bytes > 0 ? A + (int64_t)gm * a_ld + gk : A, bytes);
            } else {
                const int kk = tid >> 3, mq = (tid & 7) * 4;
                const int gk = k0 + kk, gm = m0 + mq;
                const int bytes = (gk < K) ? max(0, min(16, (M - gm) * 4)) : 0;
                cp16(&As[st][kk][mq], bytes > 0 ? A + (int64_t)gk * a_ld + gm : A, bytes);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {   // B: 256 chunks
            const int c = tid + u * 128;
            if (B_KFAST) {
                const int row = c >> 2, kq = (c & 3) * 4;
                const int gn = n0 + row, gk = k0 + kq;
                const int bytes = (gn < N) ? max(0, min(16, (K - gk) * 4)) : 0;
                cp16(&Bs[st][row][kq], bytes > 0 ? B + (int64_t)gn * b_ld + gk : B, bytes);
            } else {
                const int kk = c >> 4, nq = (c & 15) * 4;
                const int gk = k0 + kk, gn = n0 + nq;
                const int bytes = (gk < K) ? max(0, min(16, (N - gn) * 4)) : 0;
                cp16(&Bs[st][kk][nq], bytes > 0 ? B + (int64_t)gk * b_ld + gn : B, bytes);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int nkb = kb_hi - kb_lo;
#pragma unroll
    for (int p = 0; p < NST - 1; ++p) {
        if (p < nkb) issue((kb_lo + p) * BK, p);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < nkb; ++kb) {
        asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
        group_sync();
        if (kb + NST - 1 < nkb) issue((kb_lo + kb + NST - 1) * BK, (kb + NST - 1) % NST);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        const int st = kb % NST;
#pragma unroll
        for (int kk0 = 0; kk0 < BK; kk0 += 4) {
            float a[4][4], b[4][4];   // [row or col][k step]
            if (A_KFAST) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[st][ty * 4 + i][kk0]);
                    a[i][0] = t.x; a[i][1] = t.y; a[i][2] = t.z; a[i][3] = t.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[st][kk0 + q][ty * 4]);
                    a[0][q] = t.x; a[1][q] = t.y; a[2][q] = t.z; a[3][q] = t.w;
                }
            }
            if (B_KFAST) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // columns tx, tx+16, tx+32, tx+48: consecutive lanes read consecutive rows (80 B apart),
                    // which is bank-conflict free; 4 consecutive columns per lane would be a 4-way conflict
                    const float4 t = *reinterpret_cast<const float4*>(&Bs[st][tx + 16 * j][kk0]);
                    b[j][0] = t.x; b[j][1] = t.y; b[j][2] = t.z; b[j][3] = t.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(&Bs[st][kk0 + q][tx * 4]);
                    b[0][q] = t.x; b[1][q] = t.y; b[2][q] = t.z; b[3][q] = t.w;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][q], b[j][q], acc[i][j]);
        }
    }
    // fixed-order reduction of the K-split partial tiles: groups 1..3 park their accumulators in their
    // own (now idle) operand ring, group 0 adds them in group order and runs the epilogue
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float* red = reinterpret_cast<float*>(&sm);        // >= 32*64 floats per group
    if (grp > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(red + ((ty * 4 + i) * 16 + tx) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    if (grp > 0) return;
#pragma unroll
    for (int g = 1; g < V_KSPLIT; ++g) {
        const float* rg = reinterpret_cast<const float*>(&reinterpret_cast<VSmem<A_KFAST, B_KFAST>*>(vsmem_raw)[g]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(rg + ((ty * 4 + i) * 16 + tx) * 4);
            acc[i][0] += t.x; acc[i][1] += t.y; acc[i][2] += t.z; acc[i][3] += t.w;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + (B_KFAST ? tx + 16 * j : tx * 4 + j);
            if (gn >= N) continue;
            float y = e.alpha * acc[i][j];
            if (e.bias) y += e.bias[gn];
            if (e.addend) y += e.addend_scale * e.addend[(int64_t)gm * e.ld_addend + gn];
            if (e.act) y = tanhf(y);
            if (e.mulY) {
                const float t = e.mulY[(int64_t)gm * e.ldy + gn];
                y *= (1.f - t * t);
            }
            C[(int64_t)gm * ldc + gn] = y;
        }
    }
}

template <bool A_KFAST, bool B_KFAST>
static int launch_v(dim3 grid, const float* A, int64_t a_ld, const float* B, int64_t b_ld, float* C, int64_t ldc, int M,
                    int N, int K, const GemmEpi& e, cudaStream_t s) {
    constexpr int smem = (int)sizeof(VSmem<A_KFAST, B_KFAST>) * V_KSPLIT;
    static bool attr = false;
    if (!attr) {
        B200_CUDA_OK(cudaFuncSetAttribute(k_simt_gemm_v<A_KFAST, B_KFAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    B200_CUDA_OK(launch_pdl(k_simt_gemm_v<A_KFAST, B_KFAST>, grid, dim3(128 * V_KSPLIT), smem, s, A, a_ld, B, b_ld, C, ldc, M, N,
                            K, e));
    return 0;
}

int launch_simt_gemm(Ctx* c, int mode, const float* A, int64_t a_rs, int64_t a_cs, const float* B,
                     int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, int M, int N, int K,
                     const GemmEpi& e, cudaStream_t s) {
    if (M == 0 || N == 0) return 0;
    dim3 g64((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 64));
    dim3 g32((unsigned)cdiv(N, 32), (unsigned)cdiv(M, 32));
    B200_REQUIRE(g32.y <= 65535, B200VAE_EINVAL, "simt_gemm: M too large (%d)", M);
    const int sms = c->num_sms > 0 ? c->num_sms : 148;
    // the hidden layers are ~500 x 600: 64x64 tiles would occupy half the SMs with one 4-warp CTA each
    // (one warp per scheduler = every dependency stall is exposed).  32x32 tiles with a 2x4 micro-tile
    // give ~450 CTAs of 4 warps -> 3 CTAs / SM to switch between.
    const bool small = (int64_t)g64.x * g64.y < 4 * (int64_t)sms;
    if (mode == EPI_STORE) {
        // aligned operands -> vectorised kernel
        const bool a_k = (a_cs == 1), a_m = (a_rs == 1);
        const bool b_k = (b_rs == 1), b_n = (b_cs == 1);
        const int64_t a_ld = a_k ? a_rs : a_cs, b_ld = b_k ? b_cs : b_rs;
        const bool ok = (a_k || a_m) && (b_k || b_n) && (a_ld % 4 == 0) && (b_ld % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
        if (ok) {
            dim3 gv((unsigned)cdiv(N, 64), (unsigned)cdiv(M, 32));
            if (a_k && b_k)       B200_CHECK((launch_v<true, true>(gv, A, a_ld, B, b_ld, C, ldc, M, N, K, e, s)));
            else if (a_k && !b_k) B200_CHECK((launch_v<true, false>(gv, A, a_ld, B, b_ld, C, ldc, M, N, K, e, s)));
            else if (!a_k && b_k) B200_CHECK((launch_v<false, true>(gv, A, a_ld, B, b_ld, C, ldc, M, N, K, e, s)));
            else                  B200_CHECK((launch_v<false, false>(gv, A, a_ld, B, b_ld, C, ldc, M, N, K, e, s)));
            note(c, __func__, s);
            B200_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    switch (mode) {
        case EPI_STORE:
            if (small) k_simt_gemm<EPI_STORE, 32, 32, 2, 4><<<g32, 128, 0, s>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, e);
            else       k_simt_gemm<EPI_STORE, 64, 64, 4, 4><<<g64, 256, 0, s>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, e);
            break;
        case EPI_LSE:
            k_simt_gemm<EPI_LSE, 64, 64, 4, 4><<<g64, 256, 0, s>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, e);
            break;
        default:
            k_simt_gemm<EPI_PROB, 64, 64, 4, 4><<<g64, 256, 0, s>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, e);
            break;
    }
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// out[n] = sum_m X[m*ldx + n]: 32 columns x 32 row-lanes per CTA, fixed summation order
// (row-lane partial sums, then a fixed-order tree over the 32 lanes) -> deterministic.
__global__ void __launch_bounds__(1024)
k_colsum(const float* __restrict__ X, int64_t ldx, int M, int N, float* __restrict__ out) {
    __shared__ float sh[32][33];
    pdl_sync();
    const int n = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (n < N)
        for (int m = threadIdx.y; m < M; m += 32) acc += X[(int64_t)m * ldx + n];
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
        if (threadIdx.y < o) sh[threadIdx.y][threadIdx.x] += sh[threadIdx.y + o][threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.y == 0 && n < N) out[n] = sh[0][threadIdx.x];
}

int launch_colsum(Ctx* c, const float* X, int64_t ldx, int M, int N, float* out, cudaStream_t s) {
    if (N == 0) return 0;
    B200_CUDA_OK(launch_pdl(k_colsum, dim3((unsigned)cdiv(N, 32)), dim3(32, 32), 0, s, X, ldx, M, N, out));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// lse[r] = M + log(sum_t psum[t,r] * exp(pmax[t,r] - M)),  M = max_t pmax[t,r]; one warp per row
__global__ void k_lse_merge(const float* __restrict__ pmax, const float* __restrict__ psum, int n_tiles,
                            int M, float* __restrict__ lse) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= M) return;
    float mx = -INFINITY;
    for (int t = lane; t < n_tiles; t += 32) mx = fmaxf(mx, pmax[(int64_t)t * M + warp]);
    mx = warp_max(mx);
    float sacc = 0.f;
    for (int t = lane; t < n_tiles; t += 32) {
        float pm = pmax[(int64_t)t * M + warp];
        if (pm != -INFINITY) sacc += psum[(int64_t)t * M + warp] * expf(pm - mx);
    }
    sacc = warp_sum(sacc);
    if (lane == 0) lse[warp] = mx + logf(sacc);
}

int launch_lse_merge(Ctx* c, const float* pmax, const float* psum, int n_tiles, int M, float* lse,
                     cudaStream_t s) {
    if (M == 0) return 0;
    int threads = 256;
    k_lse_merge<<<(int)cdiv((int64_t)M * 32, threads), threads, 0, s>>>(pmax, psum, n_tiles, M, lse);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
