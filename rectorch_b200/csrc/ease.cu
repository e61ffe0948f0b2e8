// ease.cu -- EASE closed form (rectorch/models.py:1006-1026) on the device-resident CSR matrix:
//
//   G = X^T X            tcgen05 GEMM on fp16 images of user chunks expanded from the CSR (0/1 and small integer
//                        ratings are exact in fp16, the fp32 accumulation of counts is exact below 2^24)
//   P = (G + lam I)^-1   blocked in-place Gauss-Jordan in fp64 (the reference inverts in float64; the Gram matrix of
//                        an implicit-feedback matrix has a condition number that fp32 cannot carry).  SPD: no pivoting.
//   B = P / (-diag P), diag(B) = 0
//   S_u = X_u B          prediction = gather-sum of B rows over the user's history (the reference materialises the
//                        whole [n_users x n_items] score matrix; here B stays in HBM and rows are scored on demand)
#include <cuda_fp16.h>
#include <algorithm>
#include "ctx.cuh"

namespace b200 {

Ctx* null_ctx();    // engine.cu

// dense fp16 image of CSR rows [row0, row0 + rows): out[r * ld + col] = value; `out` is zeroed by the caller
__global__ void k_expand_f16(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                             const float* __restrict__ values, int64_t row0, int rows, int64_t ld, __half* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int64_t a = indptr[row0 + warp], b = indptr[row0 + warp + 1];
    for (int64_t k = a + lane; k < b; k += 32)
        out[(int64_t)warp * ld + indices[k]] = __float2half_rn(values ? values[k] : 1.f);
}

__global__ void k_gram_to_f64(const float* __restrict__ G32, double* __restrict__ G, int64_t n, double lam) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const int64_t r = i / n, c = i - r * n;
    G[i] = (double)G32[i] + (r == c ? lam : 0.0);
}

// ---- blocked in-place Gauss-Jordan inversion (fp64, no pivoting: the matrix is symmetric positive definite) ----
constexpr int GJ_NB = 64;

// T = inverse of the diagonal block A[k0:k0+nb, k0:k0+nb]; one CTA, block in shared memory
__global__ void __launch_bounds__(1024)
k_gj_block_inverse(const double* __restrict__ A, int64_t n, int64_t k0, int nb, double* __restrict__ T) {
    __shared__ double S[GJ_NB][GJ_NB + 1];
    __shared__ double col[GJ_NB];
    const int tid = threadIdx.x;
    for (int e = tid; e < nb * nb; e += blockDim.x) S[e / nb][e % nb] = A[(k0 + e / nb) * n + k0 + e % nb];
    __syncthreads();
    for (int p = 0; p < nb; ++p) {
        const double d = 1.0 / S[p][p];
        __syncthreads();
        if (tid < nb) {
            col[tid] = S[tid][p];
            if (tid != p) S[p][tid] *= d;
        }
        __syncthreads();
        for (int e = tid; e < nb * nb; e += blockDim.x) {
            const int i = e / nb, j = e % nb;
            if (i == p) continue;
            if (j == p) S[i][p] = -col[i] * d;
            else        S[i][j] -= col[i] * S[p][j];
        }
        if (tid == 0) S[p][p] = d;
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += blockDim.x) T[e] = S[e / nb][e % nb];
}

// C[M x N] (ldc) = beta * C + alpha * A[M x K] (lda) * B[K x N] (ldb), all row-major fp64.  64x64 tiles, 256 threads,
// 4x4 outputs per thread, K in steps of 16 through shared memory.
__global__ void __launch_bounds__(256)
k_dgemm(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t lda, const double* __restrict__ B,
        int64_t ldb, double beta, double* __restrict__ C, int64_t ldc) {
    __shared__ double As[16][64 + 1];
    __shared__ double Bs[16][64 + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int e = threadIdx.x; e < 64 * 16; e += 256) {
            const int m = e >> 4, k = e & 15;                     // A tile: 64 rows x 16 k (k contiguous in memory)
            As[k][m] = (m0 + m < M && k0 + k < K) ? A[(m0 + m) * lda + k0 + k] : 0.0;
            const int kk = e >> 6, nn = e & 63;                   // B tile: 16 k x 64 cols (n contiguous in memory)
            Bs[kk][nn] = (k0 + kk < K && n0 + nn < N) ? B[(int64_t)(k0 + kk) * ldb + n0 + nn] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t nn = n0 + tx + 16 * j;
            if (nn >= N) continue;
            double* c = C + m * ldc + nn;
            *c = (beta == 0.0 ? 0.0 : beta * *c) + alpha * acc[i][j];
        }
    }
}

// Panels of one elimination step over the pivot block K = [k0, k0 + nb):
//   R[:, K]   = T                                    (R = T * A[K, :] was computed by the GEMM before)
//   Cp[i, :]  = A[i, K] for i outside K, 0 inside    (the multipliers)
//   A[i, K]   = 0 for i outside K                    (so that A -= Cp R leaves -A[i,K] T there)
__global__ void k_gj_panels(double* __restrict__ A, int64_t n, int64_t k0, int nb, const double* __restrict__ T,
                            double* __restrict__ R, double* __restrict__ Cp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    const int j = threadIdx.x;
    if (i >= n || j >= nb) return;
    const bool in_k = (i >= k0 && i < k0 + nb);
    if (in_k) {
        R[(i - k0) * n + k0 + j] = T[(i - k0) * nb + j];
        Cp[i * nb + j] = 0.0;
    } else {
        Cp[i * nb + j] = A[i * n + k0 + j];
        A[i * n + k0 + j] = 0.0;
    }
}

// B = P / (-diag P) column-wise, diag(B) = 0, written as fp32
__global__ void k_ease_finish(const double* __restrict__ P, int64_t n, float* __restrict__ Bm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const int64_t r = i / n, c = i - r * n;
    Bm[i] = (r == c) ? 0.f : (float)(P[i] / (-P[c * n + c]));
}

// out[r, :] = sum_k x_k * Bm[item_k, :] over the non-zeros of CSR row row_ids[r]; CTA = (row, 1024-column slab)
__global__ void __launch_bounds__(256)
k_ease_scores(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
              const int32_t* __restrict__ row_ids, int64_t n_items, const float* __restrict__ Bm, float* __restrict__ out) {
    const int r = blockIdx.y;
    const int64_t gr = row_ids ? (int64_t)row_ids[r] : (int64_t)r;
    const int64_t a = indptr[gr], b = indptr[gr + 1];
    const int64_t c0 = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    if (c0 >= n_items) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (n_items % 4 == 0);
    for (int64_t k = a; k < b; ++k) {
        const float x = values ? values[k] : 1.f;
        const float* row = Bm + (int64_t)indices[k] * n_items + c0;
        if (vec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(row));
            acc[0] = fmaf(x, t.x, acc[0]); acc[1] = fmaf(x, t.y, acc[1]); acc[2] = fmaf(x, t.z, acc[2]); acc[3] = fmaf(x, t.w, acc[3]);
        } else {
            for (int i = 0; i < 4 && c0 + i < n_items; ++i) acc[i] = fmaf(x, __ldg(row + i), acc[i]);
        }
    }
    for (int i = 0; i < 4 && c0 + i < n_items; ++i) out[(int64_t)r * n_items + c0 + i] = acc[i];
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vae_ease_gram(const int64_t* indptr, const int32_t* indices, const float* values, int64_t n_users, int32_t n_items,
                      float* G32, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(indptr && G32 && n_users >= 1 && n_items >= 16, B200VAE_EINVAL, "bad argument (n_items must be >= 16)");
    Ctx* c = null_ctx();
    const int64_t ld = round_up(n_items, 8);
    // chunk of users whose dense fp16 image stays under ~1 GB
    const int64_t chunk = std::max<int64_t>(64, std::min<int64_t>(n_users, (int64_t)(1ll << 30) / (2 * ld) / 64 * 64));
    __half* X16 = nullptr;
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&X16), (size_t)(chunk * ld) * sizeof(__half)));
    int rc = 0;
    for (int64_t u0 = 0; u0 < n_users && !rc; u0 += chunk) {
        const int rows = (int)std::min<int64_t>(chunk, n_users - u0);
        if (cudaMemsetAsync(X16, 0, (size_t)(chunk * ld) * sizeof(__half), s) != cudaSuccess) { rc = B200VAE_ECUDA; break; }
        k_expand_f16<<<(unsigned)cdiv((int64_t)rows * 32, 256), 256, 0, s>>>(indptr, indices, values, u0, rows, ld, X16);
        TcEpi e;
        e.accumulate = (u0 > 0) ? 1 : 0;
        e.n_fastest = 1;
        // G += Xc^T Xc : both operands are the chunk, given as [K = users x M/N = items] (MN-major)
        rc = launch_tc_gemm(c, TC_EPI_STORE, X16, ld, 1, X16, ld, 1, G32, n_items, n_items, n_items, rows, e, s);
    }
    cudaStreamSynchronize(s);
    cudaFree(X16);
    if (!rc && cudaGetLastError() != cudaSuccess) rc = B200VAE_ECUDA;
    return rc;
}

int b200vae_ease_solve(const float* G32, int32_t n_items, double lam, float* Bm, void* stream) {
    // Bm = P / (-diag P) with P = (G + lam I)^-1, diag(Bm) = 0        (models.py:1012-1017)
    cudaStream_t s = (cudaStream_t)stream;
    B200_REQUIRE(G32 && Bm && n_items >= 1, B200VAE_EINVAL, "bad argument");
    const int64_t n = n_items;
    double *A = nullptr, *T = nullptr, *R = nullptr, *Cp = nullptr;
    B200_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&A), (size_t)(n * n) * sizeof(double)));
    int rc = 0;
    if (cudaMalloc(reinterpret_cast<void**>(&T), GJ_NB * GJ_NB * sizeof(double)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&R), (size_t)(GJ_NB * n) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&Cp), (size_t)(n * GJ_NB) * sizeof(double)) != cudaSuccess)
        rc = B200VAE_ECUDA;
    if (!rc) {
        k_gram_to_f64<<<(unsigned)cdiv(n * n, 256), 256, 0, s>>>(G32, A, n, lam);
        for (int64_t k0 = 0; k0 < n; k0 += GJ_NB) {
            const int nb = (int)std::min<int64_t>(GJ_NB, n - k0);
            k_gj_block_inverse<<<1, 1024, 0, s>>>(A, n, k0, nb, T);
            // R [nb x n] = T [nb x nb] * A[K, :] [nb x n]
            k_dgemm<<<dim3((unsigned)cdiv(n, 64), 1), 256, 0, s>>>(nb, (int)n, nb, 1.0, T, nb, A + k0 * n, n, 0.0, R, n);
            k_gj_panels<<<(unsigned)cdiv(n, 4), dim3(GJ_NB, 4), 0, s>>>(A, n, k0, nb, T, R, Cp);
            // A -= Cp [n x nb] * R [nb x n]     (rows of the pivot block have zero multipliers and are replaced below)
            k_dgemm<<<dim3((unsigned)cdiv(n, 64), (unsigned)cdiv(n, 64)), 256, 0, s>>>((int)n, (int)n, nb, -1.0, Cp, nb, R, n, 1.0, A, n);
            if (cudaMemcpyAsync(A + k0 * n, R, (size_t)nb * n * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
                rc = B200VAE_ECUDA;
                break;
            }
        }
        k_ease_finish<<<(unsigned)cdiv(n * n, 256), 256, 0, s>>>(A, n, Bm);
    }
    cudaStreamSynchronize(s);
    if (!rc && cudaGetLastError() != cudaSuccess) { set_error("EASE solve: %s", cudaGetErrorString(cudaPeekAtLastError())); rc = B200VAE_ECUDA; }
    cudaFree(A); cudaFree(T); cudaFree(R); cudaFree(Cp);
    return rc;
}

int b200vae_ease_scores(const int64_t* indptr, const int32_t* indices, const float* values, const int32_t* row_ids,
                        int32_t n_rows, int32_t n_items, const float* Bm, float* out, void* stream) {
    B200_REQUIRE(indptr && Bm && out && n_rows >= 0, B200VAE_EINVAL, "bad argument");
    if (n_rows == 0) return 0;
    B200_REQUIRE(n_rows <= 65535, B200VAE_EINVAL, "at most 65535 rows per call");
    k_ease_scores<<<dim3((unsigned)cdiv(n_items, 1024), (unsigned)n_rows), 256, 0, (cudaStream_t)stream>>>(
        indptr, indices, values, row_ids, n_items, Bm, out);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
