// topk.cu -- K10: per-user top-K by MSB-first radix select + ranking metrics on device.
//
// Replaces bottleneck.argpartition + numpy fancy indexing on a [B x n_items] host copy
// (metrics.py:140-147, 190-196, 233-238, 276-285; evaluation.py:102-104): only
// n_metrics x B floats leave the device.
//
// One CTA (256 threads) per user row.
//   Long rows (n_items >= 16384) -- ONE read of the scores instead of seven:
//   0. 4096 keys are sampled from the row (512 evenly spaced 32-byte sectors); a radix select IN SHARED MEMORY finds
//      the sample's r-th largest key tau, r chosen so that ~k + 6 sigma + margin row elements lie above it;
//      one coalesced pass over the row appends every key > tau to a shared-memory candidate list (<= 4096);
//      if the list holds at least k entries it contains the exact top-k (all ties of the k-th key included) and goes
//      straight to step 3.  Otherwise (heavy ties, adversarial rows; P ~ 3e-4 on continuous scores) fall through:
//   1. 4 passes of an 8-bit histogram over the order-preserving uint32 image of the scores
//      narrow down the k-th largest key T and the number of ties at T to keep;
//   2. all keys > T plus the lowest-index ties are compacted into shared memory
//      (deterministic tie rule: smaller item id first -- argpartition's is unspecified);
//   3. a bitonic sort of the candidates (key descending, item id ascending) gives the ranked list (needed by ndcg/mrr);
//   4. membership of each ranked item in the held-out CSR row is a binary search;
//   5. each requested (metric, k) is reduced from the ranked hit list.
#include "ctx.cuh"

namespace b200 {

constexpr int TOPK_MAX = 1024;
constexpr int TOPK_THREADS = 256;
constexpr int TOPK_CAND = 4096;        // candidate capacity of the single-read path
constexpr int TOPK_SAMPLE = 4096;      // sampled keys (alias the candidate storage)
constexpr int TOPK_SAMPLE_MIN_I = 16384;

__device__ __forceinline__ uint32_t f2key(float x) {
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// k-th largest of n keys served by key_at(j): MSB-first radix select, 4 passes of an 8-bit shared-memory histogram.
// Leaves the key in *s_prefix and the number of ties at that key still to take in *s_remaining.  Block-uniform.
template <typename F>
__device__ __forceinline__ void radix_select(F key_at, int n, uint32_t k, unsigned int* hist, uint32_t* s_prefix,
                                             uint32_t* s_remaining) {
    const int tid = threadIdx.x;
    if (tid == 0) { *s_prefix = 0; *s_remaining = k; }
    uint32_t mask = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = *s_prefix;
        for (int j = tid; j < n; j += TOPK_THREADS) {
            uint32_t u = key_at(j);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t rem = *s_remaining, above = 0;
            int b = 255;
            for (; b > 0; --b) {
                if (above + hist[b] >= rem) break;
                above += hist[b];
            }
            *s_remaining = rem - above;
            *s_prefix = prefix | ((uint32_t)b << shift);
        }
        mask |= 0xFFu << shift;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TOPK_THREADS)
k_topk_metrics(const float* __restrict__ scores, int I, BatchView gt, const int32_t* __restrict__ kinds,
               const int32_t* __restrict__ ks, int n_metrics, int kmax, float* __restrict__ out,
               int32_t* __restrict__ topk_idx, int sampling) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long cand[TOPK_CAND];
    __shared__ float hitval[TOPK_MAX];
    __shared__ uint32_t s_prefix, s_remaining;
    __shared__ unsigned int s_count;
    __shared__ unsigned int warp_ties[TOPK_THREADS / 32 + 1];

    const int r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* row = scores + (int64_t)r * I;

    // ---- 0. single-read path ----------------------------------------------------------------
    int n_sort = 0;                   // > 0: cand[0 .. n_sort) is ready for the sort
    if (sampling && I >= TOPK_SAMPLE_MIN_I) {
        uint32_t* samp = reinterpret_cast<uint32_t*>(cand);
        for (int i = tid; i < TOPK_SAMPLE; i += TOPK_THREADS) {
            const int cl = i >> 3;                                            // 512 clusters of 8 consecutive scores
            const int j = (int)(((int64_t)cl * I) / (TOPK_SAMPLE / 8)) + (i & 7);
            samp[i] = (j < I) ? f2key(row[j]) : 0u;
        }
        __syncthreads();
        const float f = (float)TOPK_SAMPLE / (float)I;
        const float mean = f * (float)kmax;
        // 6 sigma + 8: the sample is 512 clusters of 8 neighbouring items, whose scores are correlated (similar
        // popularity), so its effective size is below 4096; a larger margin only lengthens the candidate list a little
        const int rs = min(TOPK_SAMPLE, (int)ceilf(mean + 6.f * sqrtf(mean) + 8.f));
        radix_select([&](int j) { return samp[j]; }, TOPK_SAMPLE, (uint32_t)rs, hist, &s_prefix, &s_remaining);
        const uint32_t tau = s_prefix;
        if (tid == 0) s_count = 0;
        __syncthreads();              // the sample is dead from here: cand is reused for the candidates
        auto take = [&](float x, int j) {
            const uint32_t u = f2key(x);
            if (u > tau) {
                const unsigned int p = atomicAdd(&s_count, 1u);
                if (p < (unsigned)TOPK_CAND)
                    cand[p] = ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
            }
        };
        int j = tid;
        for (; j + 7 * TOPK_THREADS < I; j += 8 * TOPK_THREADS) {     // 8 independent loads in flight per thread
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = __ldcs(row + j + q * TOPK_THREADS);
#pragma unroll
            for (int q = 0; q < 8; ++q) take(x[q], j + q * TOPK_THREADS);
        }
        for (; j < I; j += TOPK_THREADS) take(row[j], j);
        __syncthreads();
        const unsigned int c = s_count;
        if (c >= (unsigned)kmax && c <= (unsigned)TOPK_CAND) {
            n_sort = 1;
            while (n_sort < (int)c) n_sort <<= 1;
            for (int i = (int)c + tid; i < n_sort; i += TOPK_THREADS) cand[i] = 0ull;   // padding sorts last
        }
        __syncthreads();
    }

    if (n_sort == 0) {
    // ---- 1. radix select ------------------------------------------------------------------
    radix_select([&](int j) { return f2key(row[j]); }, I, (uint32_t)kmax, hist, &s_prefix, &s_remaining);
    const uint32_t T = s_prefix;
    const uint32_t n_ties = s_remaining;         // ties at T to keep (>= 1)
    const uint32_t n_gt = (uint32_t)kmax - n_ties;

    // ---- 2. compaction ---------------------------------------------------------------------
    if (tid == 0) s_count = 0;
    for (int i = tid; i < TOPK_MAX; i += TOPK_THREADS) cand[i] = 0ull;   // padding sorts last
    __syncthreads();
    // keys strictly above T: any order, the sort fixes it
    for (int j = tid; j < I; j += TOPK_THREADS) {
        uint32_t u = f2key(row[j]);
        if (u > T) {
            unsigned int p = atomicAdd(&s_count, 1u);
            cand[p] = ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
        }
    }
    // ties at T in increasing item order: each warp owns a contiguous range of the row
    const int n_warps = TOPK_THREADS / 32;
    const int per_warp = (int)(((int64_t)I + n_warps - 1) / n_warps);
    const int lo = wid * per_warp, hi = min(I, lo + per_warp);
    unsigned int my_ties = 0;
    for (int j0 = lo; j0 < hi; j0 += 32) {
        int j = j0 + lane;
        bool is = (j < hi) && (f2key(row[j]) == T);
        my_ties += __popc(__ballot_sync(0xffffffffu, is));
    }
    if (lane == 0) warp_ties[wid] = my_ties;
    __syncthreads();
    if (tid == 0) {
        unsigned int acc = 0;
        for (int w = 0; w < n_warps; ++w) { unsigned int t = warp_ties[w]; warp_ties[w] = acc; acc += t; }
    }
    __syncthreads();
    {
        unsigned int pos = warp_ties[wid];
        for (int j0 = lo; j0 < hi && pos < n_ties; j0 += 32) {
            int j = j0 + lane;
            bool is = (j < hi) && (f2key(row[j]) == T);
            unsigned int m = __ballot_sync(0xffffffffu, is);
            if (is) {
                unsigned int p = pos + __popc(m & ((1u << lane) - 1u));
                if (p < n_ties)
                    cand[n_gt + p] = ((unsigned long long)T << 32) |
                                     (unsigned long long)(0xFFFFFFFFu - (uint32_t)j);
            }
            pos += __popc(m);
        }
    }
    __syncthreads();
    n_sort = 1;
    while (n_sort < kmax) n_sort <<= 1;
    }   // multi-pass path

    // ---- 3. bitonic sort, descending, over a power of two >= the number of candidates ----------
    for (int size = 2; size <= n_sort; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < n_sort; i += TOPK_THREADS) {
                int p = i ^ stride;
                if (p > i) {
                    bool desc = ((i & size) == 0);
                    unsigned long long a = cand[i], b = cand[p];
                    if ((a < b) == desc) { cand[i] = b; cand[p] = a; }
                }
            }
            __syncthreads();
        }
    }

    // ---- 4. membership in the held-out row ---------------------------------------------------
    const int64_t gr = gt.row_ids ? (int64_t)gt.row_ids[r] : (int64_t)r;
    const int64_t ga = gt.indptr[gr], gb = gt.indptr[gr + 1];
    for (int i = tid; i < kmax; i += TOPK_THREADS) {
        int item = (int)(0xFFFFFFFFu - (uint32_t)(cand[i] & 0xFFFFFFFFull));
        if (topk_idx) topk_idx[(int64_t)r * kmax + i] = item;
        int64_t a = ga, b = gb;
        float val = 0.f;
        while (a < b) {
            int64_t mid = (a + b) >> 1;
            int c = gt.indices[mid];
            if (c == item) { val = gt.values ? gt.values[mid] : 1.f; break; }
            if (c < item) a = mid + 1; else b = mid;
        }
        hitval[i] = val;
    }
    // row statistics of the ground truth: #positives and sum of values
    __shared__ float s_npos, s_gsum;
    if (wid == 0) {
        float np_ = 0.f, gs = 0.f;
        for (int64_t k = ga + lane; k < gb; k += 32) {
            float v = gt.values ? gt.values[k] : 1.f;
            np_ += (v > 0.f) ? 1.f : 0.f;
            gs += v;
        }
        np_ = warp_sum(np_);
        gs = warp_sum(gs);
        if (lane == 0) { s_npos = np_; s_gsum = gs; }
    }
    __syncthreads();

    // ---- 5. metrics ---------------------------------------------------------------------------
    if (tid < n_metrics) {
        int kind = kinds[tid];
        int k = min(ks[tid], I);
        float res = 0.f;
        if (kind == 0) {                           // recall  (metrics.py:190-196)
            float num = 0.f;
            for (int i = 0; i < k; ++i) num += (hitval[i] > 0.f) ? 1.f : 0.f;
            res = num / fminf((float)k, s_npos);
        } else if (kind == 1) {                    // ndcg    (metrics.py:140-147)
            float dcg = 0.f, idcg = 0.f;
            int n_ideal = min((int)s_gsum, k);
            for (int i = 0; i < k; ++i) {
                float tp = 1.f / log2f((float)(i + 2));
                dcg += hitval[i] * tp;
                if (i < n_ideal) idcg += tp;
            }
            res = dcg / idcg;
        } else if (kind == 2) {                    // hit     (metrics.py:233-238)
            for (int i = 0; i < k; ++i) if (hitval[i] > 0.f) { res = 1.f; break; }
        } else {                                   // mrr     (metrics.py:276-285)
            for (int i = 0; i < k; ++i) if (hitval[i] != 0.f) { res = 1.f / (float)(i + 1); break; }
        }
        out[(int64_t)tid * gt.B + r] = res;
    }
}

int launch_topk_metrics(Ctx* c, const float* scores, int I, const BatchView& gt, const int32_t* kinds,
                        const int32_t* ks, int n_metrics, int kmax, float* out, int32_t* topk_idx,
                        cudaStream_t s) {
    if (gt.B == 0) return 0;
    B200_REQUIRE(kmax >= 1 && kmax <= TOPK_MAX && kmax <= I, B200VAE_EINVAL,
                 "topk: k must be in [1, min(%d, n_items)] (got %d)", TOPK_MAX, kmax);
    B200_REQUIRE(n_metrics <= TOPK_THREADS, B200VAE_EINVAL, "topk: too many metrics");
    const char* e = getenv("B200VAE_TOPK_SAMPLE");          // 0: always the multi-pass radix select (tests, comparisons)
    const int sampling = (e && atoi(e) == 0) ? 0 : 1;
    k_topk_metrics<<<gt.B, TOPK_THREADS, 0, s>>>(scores, I, gt, kinds, ks, n_metrics, kmax, out, topk_idx, sampling);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
