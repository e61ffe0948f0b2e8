// sparse.cu -- CSR batch handling and the sparse ends of the network.
//
//  K1  expand / dense_to_csr   : DataSampler.__iter__ (samplers.py:99-105) and its inverse
//  K2  spmm_gather             : F.normalize + nn.Dropout + first encoder nn.Linear
//                                (nets.py:395-401) as a gather-sum of item-major weight rows;
//                                also  g_u = sum_j t_uj W_d[j,:]  for the multinomial NLL
//  K7  spmm_scatter            : the transposed operation for the weight gradients
//
// All kernels are HBM/L2-bound row gathers: one CTA per user row, threads across the
// hidden dimension so every warp reads 128 B-contiguous pieces of a weight row.
#include "ctx.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// scans
// ------------------------------------------------------------------------------------------
__global__ void k_batch_scan(const int64_t* __restrict__ indptr, const int32_t* __restrict__ row_ids,
                             int B, int64_t cap, int64_t* __restrict__ bp, int32_t* __restrict__ sp, int* err) {
    // single CTA, 1024 threads; chunked Hillis-Steele scans of the row lengths (bp) and of the
    // number of SPMM_SEG-sized work segments per row (sp)
    __shared__ int64_t sh[1024];
    __shared__ int32_t sg[1024];
    __shared__ int64_t carry;
    __shared__ int32_t carry_s;
    pdl_sync();
    if (threadIdx.x == 0) { carry = 0; carry_s = 0; }
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        int r = base + threadIdx.x;
        int64_t len = 0;
        if (r < B) {
            int64_t gr = row_ids ? (int64_t)row_ids[r] : (int64_t)r;
            len = indptr[gr + 1] - indptr[gr];
        }
        int32_t nseg = (r < B) ? (int32_t)max((int64_t)1, (len + SPMM_SEG - 1) / SPMM_SEG) : 0;
        sh[threadIdx.x] = len;
        sg[threadIdx.x] = nseg;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int64_t t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            int32_t u = (threadIdx.x >= o) ? sg[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            sg[threadIdx.x] += u;
            __syncthreads();
        }
        if (r < B) {
            bp[r] = carry + sh[threadIdx.x] - len;
            sp[r] = carry_s + sg[threadIdx.x] - nseg;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry += sh[1023]; carry_s += sg[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bp[B] = carry;
        sp[B] = carry_s;
        if (carry > cap) *err = 1;
    }
}

int launch_batch_scan(Ctx* c, const int64_t* indptr, const int32_t* row_ids, int B, int64_t cap,
                      int64_t* bp, int32_t* sp, cudaStream_t s) {
    B200_CUDA_OK(launch_pdl(k_batch_scan, dim3(1), dim3(1024), 0, s, indptr, row_ids, B, cap, bp, sp, c->d_err));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void k_scan_i64(const int64_t* __restrict__ lens, int n, int64_t cap,
                           int64_t* __restrict__ out, int* err) {
    __shared__ int64_t sh[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int r = base + threadIdx.x;
        int64_t len = (r < n) ? lens[r] : 0;
        sh[threadIdx.x] = len;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int64_t t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (r < n) out[r] = carry + sh[threadIdx.x] - len;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[n] = carry;
        if (carry > cap) *err = 1;
    }
}

int launch_scan_i64(Ctx* c, const int64_t* lens, int n, int64_t cap, int64_t* out, cudaStream_t s) {
    k_scan_i64<<<1, 1024, 0, s>>>(lens, n, cap, out, c->d_err);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// batch_prep: per row  s = 1/max(||x||,1e-12);  xt = x*s*keep/(1-p)   (nets.py:395-397)
// one warp per row
// ------------------------------------------------------------------------------------------
template <bool SCAN>
__global__ void __launch_bounds__(256)
k_batch_prep(BatchView v, float p, uint64_t seed, uint64_t step, int64_t row_offset,
             const uint8_t* __restrict__ keep_tape, int train, int64_t cap,
             float* __restrict__ xt, float* __restrict__ row_sum_out,
             int32_t* __restrict__ mark, int32_t mark_step, int n_items,
             int64_t* __restrict__ bp_out, int32_t* __restrict__ sp_out, int* __restrict__ err) {
    __shared__ int64_t s_len[8];
    __shared__ int64_t s_pl[8];
    __shared__ int32_t s_ps[8];
    pdl_sync();
    const int wid = threadIdx.x >> 5;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t o_scan = 0;
    if (SCAN) {
        // The exclusive scans of the row lengths (bp) and of the per-row segment counts (sp) that k_batch_scan would
        // produce, without the extra kernel on the critical path: every CTA sums the rows before its own (<= 2048 rows,
        // L2-resident row pointers) and scans its 8 rows in shared memory.
        const int row0 = blockIdx.x * 8;
        int64_t pl = 0;
        int32_t ps = 0;
        for (int r = threadIdx.x; r < row0; r += 256) {
            const int64_t g = v.row_ids ? (int64_t)v.row_ids[r] : (int64_t)r;
            const int64_t len = v.indptr[g + 1] - v.indptr[g];
            pl += len;
            ps += (int32_t)max((int64_t)1, (len + SPMM_SEG - 1) / SPMM_SEG);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            pl += __shfl_xor_sync(0xffffffffu, pl, d);
            ps += __shfl_xor_sync(0xffffffffu, ps, d);
        }
        int64_t my_len = 0;
        if (warp < v.B) {
            const int64_t g = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
            my_len = v.indptr[g + 1] - v.indptr[g];
        }
        if (lane == 0) { s_pl[wid] = pl; s_ps[wid] = ps; s_len[wid] = my_len; }
        __syncthreads();
        int64_t off = 0;
        int32_t soff = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { off += s_pl[i]; soff += s_ps[i]; }
        for (int i = 0; i < wid; ++i) {
            off += s_len[i];
            soff += (int32_t)max((int64_t)1, (s_len[i] + SPMM_SEG - 1) / SPMM_SEG);
        }
        if (warp < v.B && lane == 0) {
            bp_out[warp] = off;
            sp_out[warp] = soff;
            if (warp == v.B - 1) {
                bp_out[v.B] = off + my_len;
                sp_out[v.B] = soff + (int32_t)max((int64_t)1, (my_len + SPMM_SEG - 1) / SPMM_SEG);
                if (off + my_len > cap) *err = 1;
            }
        }
        o_scan = off;
    }
    if (warp >= v.B) return;
    int64_t gr = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
    int64_t a = v.indptr[gr], b = v.indptr[gr + 1];
    int64_t o = SCAN ? o_scan : v.bp[warp];
    if (o + (b - a) > cap) return;   // capacity overflow flagged by the scan
    float ss = 0.f, sx = 0.f;
    // columns >= n_items are condition flags (CMultiVAE_net.encode, nets.py:466-470): concatenated AFTER
    // F.normalize / nn.Dropout, so they take no part in the norm and are passed through unchanged
    for (int64_t k = a + lane; k < b; k += 32) {
        if (v.indices[k] >= n_items) continue;
        float x = v.values ? v.values[k] : 1.f;
        ss += x * x;
        sx += x;
    }
    ss = warp_sum(ss);
    if (row_sum_out) {               // T_u = sum_j x_uj when the input is also the target
        sx = warp_sum(sx);
        if (lane == 0) row_sum_out[warp] = sx;
    }
    float denom = fmaxf(sqrtf(ss), 1e-12f);
    bool drop = train && p > 0.f;
    float inv_keep = 1.0f / (1.0f - p);
    for (int64_t k = a + lane; k < b; k += 32) {
        float x = v.values ? v.values[k] : 1.f;
        const bool cond_col = v.indices[k] >= n_items;
        float xn = cond_col ? x : x / denom;
        if (drop && !cond_col) {
            bool keep;
            if (keep_tape) {
                keep = keep_tape[o + (k - a)] != 0;
            } else {
                uint64_t grow = (uint64_t)(gr + row_offset);
                uint4 ctr = make_uint4((uint32_t)grow, (uint32_t)(grow >> 32), (uint32_t)v.indices[k],
                                       (uint32_t)step);
                uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x0D0Du);
                uint4 r = philox4x32(ctr, key);
                keep = u32_to_unit(r.x) > p;
            }
            xn = keep ? xn * inv_keep : 0.f;
        }
        xt[o + (k - a)] = xn;
        // item rows of encoder layer 0 that this step reads (gather) and whose gradient it writes (scatter):
        // every other row has an exactly-zero gradient, which lets Adam update it off the critical path
        if (mark && xn != 0.f) mark[v.indices[k]] = mark_step;
    }
}

int launch_batch_prep(Ctx* c, const BatchView& in, float p, uint64_t seed, uint64_t step,
                      int64_t row_offset, const uint8_t* keep_tape, bool train, float* xt, float* row_sum_out,
                      int32_t* mark, int32_t mark_step, cudaStream_t s, bool with_scan) {
    if (in.B == 0) return 0;
    int threads = 256;
    int blocks = (int)cdiv((int64_t)in.B * 32, threads);
    const int n_items = c->n_items > 0 ? c->n_items : INT32_MAX;
    // with_scan: the view's bp / sp have NOT been computed (make_view(..., defer_scan)); this launch writes them
    if (with_scan)
        B200_CUDA_OK(launch_pdl(k_batch_prep<true>, dim3(blocks), dim3(threads), 0, s, in, p, seed, step, row_offset, keep_tape,
                                train ? 1 : 0, c->cfg.max_batch_nnz, xt, row_sum_out, mark, mark_step, n_items,
                                const_cast<int64_t*>(in.bp), const_cast<int32_t*>(in.sp), c->d_err));
    else
        B200_CUDA_OK(launch_pdl(k_batch_prep<false>, dim3(blocks), dim3(threads), 0, s, in, p, seed, step, row_offset, keep_tape,
                                train ? 1 : 0, c->cfg.max_batch_nnz, xt, row_sum_out, mark, mark_step, n_items,
                                (int64_t*)nullptr, (int32_t*)nullptr, c->d_err));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void k_row_sums(BatchView v, float* __restrict__ out) {
    pdl_sync();
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= v.B) return;
    int64_t gr = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
    int64_t a = v.indptr[gr], b = v.indptr[gr + 1];
    float sacc = 0.f;
    for (int64_t k = a + lane; k < b; k += 32) sacc += v.values ? v.values[k] : 1.f;
    sacc = warp_sum(sacc);
    if (lane == 0) out[warp] = sacc;
}

int launch_row_sums(Ctx* c, const BatchView& v, float* out, cudaStream_t s) {
    if (v.B == 0) return 0;
    int threads = 256;
    B200_CUDA_OK(launch_pdl(k_row_sums, dim3((unsigned)cdiv((int64_t)v.B * 32, threads)), dim3(threads), 0, s, v, out));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// K2 spmm_gather: out[r,:] = act(bias + sum_k vals[bp[r]+k] * Wt[col_k,:])
//
// Work unit = (row, segment of <= SPMM_SEG non-zeros): histories are log-normal with a 20x tail, so a
// CTA per row would leave the longest user running alone.  CTA b finds its row by a binary search in
// the segment pointer sp.  Rows with one segment write their result directly; longer rows reduce
// their partial sums with fp32 reductions in L2 into a zeroed accumulator and the CTA that finishes
// last (per-row ticket) applies bias + activation and re-zeroes the accumulator for the next call.
// VEC=4 -> float4 lanes across H (H % 4 == 0, 16 B aligned rows).  Entries whose value is 0
// (dropped by nn.Dropout) are skipped without touching their weight row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_row(const int32_t* __restrict__ sp, int B, int b) {
    int lo = 0, hi = B;          // largest r with sp[r] <= b
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (sp[mid] <= b) lo = mid; else hi = mid;
    }
    return lo;
}

template <int VEC>
__global__ void __launch_bounds__(256)
k_spmm_gather(BatchView v, const float* __restrict__ vals, const float* __restrict__ Wt, int H,
              const float* __restrict__ bias, int act, float* __restrict__ out, float* __restrict__ acc_ws,
              int* __restrict__ ticket, __half* __restrict__ out16, int64_t ld16, const __half* __restrict__ Wt16, int mod_n,
              int64_t rows_per, int det) {
    pdl_sync();
    // det: one CTA per row walks every non-zero in index order (no cross-segment reduction, so no atomics)
    if (!det && (int)blockIdx.x >= v.sp[v.B]) return;
    const int r = det ? (int)blockIdx.x : find_row(v.sp, v.B, blockIdx.x);
    const int seg = det ? 0 : blockIdx.x - v.sp[r];
    const int nseg = det ? 1 : v.sp[r + 1] - v.sp[r];
    const int64_t gr = v.row_ids ? (int64_t)v.row_ids[r] : (int64_t)r;
    const int64_t a = v.indptr[gr];
    const int len = (int)(v.indptr[gr + 1] - a);
    const int k0 = seg * SPMM_SEG, k1 = det ? len : min(len, k0 + SPMM_SEG);
    const int64_t o = v.bp[r];
    const int32_t* cols = v.indices + a;
    const float* xv = vals ? vals + o : nullptr;
    const float* raw = v.values ? v.values + a : nullptr;
    __shared__ int s_last;
    for (int h0 = threadIdx.x * VEC; h0 < H; h0 += blockDim.x * VEC) {
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        int k = k0;
        for (; k + 8 <= k1; k += 8) {   // up to 8 independent row loads in flight per thread
            float w[8][VEC];
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                x[u] = xv ? xv[k + u] : (raw ? raw[k + u] : 1.f);
#pragma unroll
                for (int i = 0; i < VEC; ++i) w[u][i] = 0.f;
                if (x[u] != 0.f) {
                    const int j = cols[k + u];
                    if (Wt16) {
                        const __half* row = Wt16 + ((int64_t)(j % mod_n) * rows_per + j / mod_n) * H + h0;
                        if (VEC == 4) {
                            const uint2 t = __ldg(reinterpret_cast<const uint2*>(row));
                            const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
                            const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
                            w[u][0] = a2.x; w[u][1 % VEC] = a2.y; w[u][2 % VEC] = b2.x; w[u][3 % VEC] = b2.y;
                        } else {
                            w[u][0] = __half2float(row[0]);
                        }
                    } else {
                        const float* row = Wt + (int64_t)j * H + h0;
                        if (VEC == 4) {
                            float4 t = __ldg(reinterpret_cast<const float4*>(row));
                            w[u][0] = t.x; w[u][1 % VEC] = t.y; w[u][2 % VEC] = t.z; w[u][3 % VEC] = t.w;
                        } else {
                            w[u][0] = __ldg(row);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(x[u], w[u][i], acc[i]);
        }
        for (; k < k1; ++k) {
            float x = xv ? xv[k] : (raw ? raw[k] : 1.f);
            if (x == 0.f) continue;
            const int j = cols[k];
            if (Wt16) {
                const __half* row = Wt16 + ((int64_t)(j % mod_n) * rows_per + j / mod_n) * H + h0;
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(x, __half2float(row[i]), acc[i]);
            } else {
                const float* row = Wt + (int64_t)j * H + h0;
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] = fmaf(x, __ldg(row + i), acc[i]);
            }
        }
        if (nseg == 1) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float y = acc[i] + (bias ? bias[h0 + i] : 0.f);
                if (act) y = tanhf(y);
                out[(int64_t)r * H + h0 + i] = y;
                if (out16) out16[(int64_t)r * ld16 + h0 + i] = __float2half_rn(f16_clamp(y));
            }
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) atomicAdd(acc_ws + (int64_t)r * H + h0 + i, acc[i]);
        }
    }
    if (nseg > 1) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(ticket + r, 1) == nseg - 1);
        __syncthreads();
        if (s_last) {
            __threadfence();
            for (int h = threadIdx.x; h < H; h += blockDim.x) {
                float y = __ldcg(acc_ws + (int64_t)r * H + h) + (bias ? bias[h] : 0.f);
                if (act) y = tanhf(y);
                out[(int64_t)r * H + h] = y;
                if (out16) out16[(int64_t)r * ld16 + h] = __float2half_rn(f16_clamp(y));
                acc_ws[(int64_t)r * H + h] = 0.f;
            }
            if (threadIdx.x == 0) ticket[r] = 0;
        }
    }
}

static int spmm_grid(Ctx* c, const BatchView& v) {
    // upper bound of the number of segments: every row has >= 1, plus nnz_cap / SEG
    int64_t g = (int64_t)v.B + c->cfg.max_batch_nnz / SPMM_SEG + 1;
    return (int)std::min<int64_t>(g, 1 << 30);
}

int launch_spmm_gather(Ctx* c, const BatchView& v, const float* vals, const float* Wt, int H,
                       const float* bias, int act, float* out, cudaStream_t s, __half* out16, int64_t ld16,
                       const __half* Wt16, int mod_n, int64_t rows_per) {
    if (v.B == 0) return 0;
    bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(Wt) & 15) == 0);
    const int det = c->deterministic ? 1 : 0;
    const int grid = det ? v.B : spmm_grid(c, v);
    if (vec) {
        int threads = (int)std::min<int64_t>(256, round_up(cdiv(H, 4), 32));
        B200_CUDA_OK(launch_pdl(k_spmm_gather<4>, dim3(grid), dim3(threads), 0, s, v, vals, Wt, H, bias, act, out, c->spmm_acc,
                                c->spmm_ticket, out16, ld16, Wt16, mod_n, rows_per, det));
    } else {
        int threads = (int)std::min<int64_t>(256, round_up(H, 32));
        B200_CUDA_OK(launch_pdl(k_spmm_gather<1>, dim3(grid), dim3(threads), 0, s, v, vals, Wt, H, bias, act, out, c->spmm_acc,
                                c->spmm_ticket, out16, ld16, Wt16, mod_n, rows_per, det));
    }
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// K7 spmm_scatter: dWt[col_k,:] += scale * vals[k] * dY[r,:]      (fp32 reductions in L2)
// same (row, segment) work units as the gather
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
k_spmm_scatter(BatchView v, const float* __restrict__ vals, float scale, const float* __restrict__ dY,
               int H, float* __restrict__ dWt, float* __restrict__ db, int mod_n, int mod_r) {
    pdl_sync();
    if ((int)blockIdx.x >= v.sp[v.B]) return;
    const int r = find_row(v.sp, v.B, blockIdx.x);
    const int seg = blockIdx.x - v.sp[r];
    const int64_t gr = v.row_ids ? (int64_t)v.row_ids[r] : (int64_t)r;
    const int64_t a = v.indptr[gr];
    const int len = (int)(v.indptr[gr + 1] - a);
    const int k0 = seg * SPMM_SEG, k1 = min(len, k0 + SPMM_SEG);
    const int64_t o = v.bp[r];
    const int32_t* cols = v.indices + a;
    const float* xv = vals ? vals + o : nullptr;
    const float* raw = v.values ? v.values + a : nullptr;
    if (db) {   // db[col_k] += scale * value_k   (sparse part of the bias gradient, same CSR walk)
        for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
            float x = xv ? xv[k] : (raw ? raw[k] : 1.f);
            if (x != 0.f) atomicAdd(db + cols[k], scale * x);
        }
    }
    for (int h0 = threadIdx.x * VEC; h0 < H; h0 += blockDim.x * VEC) {
        float d[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) d[i] = dY[(int64_t)r * H + h0 + i] * scale;
        for (int k = k0; k < k1; ++k) {
            float x = xv ? xv[k] : (raw ? raw[k] : 1.f);
            if (x == 0.f) continue;   // dropped entries contribute nothing
            if (mod_n > 1 && cols[k] % mod_n != mod_r) continue;   // another rank owns this item row
            float* row = dWt + (int64_t)cols[k] * H + h0;
            if (VEC == 4) {
                atomicAdd(reinterpret_cast<float4*>(row),
                          make_float4(x * d[0], x * d[1 % VEC], x * d[2 % VEC], x * d[3 % VEC]));
            } else {
                atomicAdd(row, x * d[0]);
            }
        }
    }
}

// Sharded variant (data parallelism, this rank owns the item rows j % mod_n == mod_r): one CTA per USER row keeps the
// row's delta in registers and walks all of its non-zeros, touching only the owned items -- 1/mod_n of the reductions
// of the global batch, and one delta load per user instead of one per 32-non-zero segment.
__global__ void __launch_bounds__(256)
k_spmm_scatter_owned(BatchView v, const float* __restrict__ vals, float scale, const float* __restrict__ dY, int H,
                     float* __restrict__ dWt, int mod_n, int mod_r) {
    pdl_sync();
    const int r = blockIdx.x;
    const int64_t gr = v.row_ids ? (int64_t)v.row_ids[r] : (int64_t)r;
    const int64_t a = v.indptr[gr];
    const int len = (int)(v.indptr[gr + 1] - a);
    const int64_t o = v.bp[r];
    const int32_t* cols = v.indices + a;
    const float* xv = vals ? vals + o : nullptr;
    const float* raw = v.values ? v.values + a : nullptr;
    for (int h0 = threadIdx.x * 4; h0 < H; h0 += blockDim.x * 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(dY + (int64_t)r * H + h0);
        const float d[4] = {d4.x * scale, d4.y * scale, d4.z * scale, d4.w * scale};
        for (int k = 0; k < len; ++k) {
            const int j = cols[k];
            if (j % mod_n != mod_r) continue;
            const float x = xv ? xv[k] : (raw ? raw[k] : 1.f);
            if (x == 0.f) continue;
            atomicAdd(reinterpret_cast<float4*>(dWt + (int64_t)j * H + h0), make_float4(x * d[0], x * d[1], x * d[2], x * d[3]));
        }
    }
}


// ------------------------------------------------------------------------------------------
// Deterministic scatter (b200vae_set_deterministic): the batch is transposed into per-item lists (integer counting,
// a single-CTA scan, a cursor fill whose order is then fixed by ranking the (row, position) keys), and one warp per
// item adds its contributions in ascending (row, position) order with plain stores.  Bit-identical from run to run;
// several times slower than the atomic version, so it is opt-in.
// ------------------------------------------------------------------------------------------
__global__ void k_det_count(BatchView v, const float* __restrict__ vals, int n_rows, int* __restrict__ count) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= v.B) return;
    const int64_t gr = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
    const int64_t a = v.indptr[gr];
    const int len = (int)(v.indptr[gr + 1] - a);
    const float* xv = vals ? vals + v.bp[warp] : nullptr;
    const float* raw = v.values ? v.values + a : nullptr;
    for (int k = lane; k < len; k += 32) {
        const float x = xv ? xv[k] : (raw ? raw[k] : 1.f);
        const int j = v.indices[a + k];
        if (x != 0.f && j < n_rows) atomicAdd(count + j, 1);   // integer: the RESULT does not depend on the order
    }
}

__global__ void __launch_bounds__(1024) k_det_scan(const int* __restrict__ count, int n_rows, int* __restrict__ off) {
    __shared__ int s_part[1024];
    const int per = (n_rows + 1023) / 1024;
    const int lo = min(n_rows, (int)threadIdx.x * per), hi = min(n_rows, lo + per);
    int sum = 0;
    for (int j = lo; j < hi; ++j) sum += count[j];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        int t = (threadIdx.x >= d) ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += t;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int j = lo; j < hi; ++j) { off[j] = run; run += count[j]; }
    if (threadIdx.x == 1023) off[n_rows] = s_part[1023];
}

__global__ void k_det_fill(BatchView v, const float* __restrict__ vals, int n_rows, const int* __restrict__ off,
                           int* __restrict__ cursor, int64_t* __restrict__ ent) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= v.B) return;
    const int64_t gr = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
    const int64_t a = v.indptr[gr];
    const int len = (int)(v.indptr[gr + 1] - a);
    const float* xv = vals ? vals + v.bp[warp] : nullptr;
    const float* raw = v.values ? v.values + a : nullptr;
    for (int k = lane; k < len; k += 32) {
        const float x = xv ? xv[k] : (raw ? raw[k] : 1.f);
        const int j = v.indices[a + k];
        if (x != 0.f && j < n_rows) ent[off[j] + atomicAdd(cursor + j, 1)] = ((int64_t)warp << 32) | (uint32_t)k;
    }
}

__global__ void __launch_bounds__(256)
k_det_reduce(BatchView v, const float* __restrict__ vals, float scale, const float* __restrict__ dY, int H, int n_rows,
             const int* __restrict__ off, const int64_t* __restrict__ ent, int64_t* __restrict__ sorted,
             float* __restrict__ dWt, float* __restrict__ db, int mod_n, int mod_r) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= n_rows) return;
    if (mod_n > 1 && j % mod_n != mod_r) return;
    const int o = off[j], n = off[j + 1] - o;
    if (n == 0) return;
    // fix the order: rank of each key among the item's keys (keys are distinct)
    for (int i = lane; i < n; i += 32) {
        const int64_t key = ent[o + i];
        int rank = 0;
        for (int q = 0; q < n; ++q) rank += (ent[o + q] < key);
        sorted[o + rank] = key;
    }
    __syncwarp();
    float bsum = 0.f;
    for (int h0 = lane; h0 < H; h0 += 32 * 8) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int q = 0; q < n; ++q) {
            const int64_t key = sorted[o + q];
            const int r = (int)(key >> 32), k = (int)(key & 0xffffffff);
            const int64_t gr = v.row_ids ? (int64_t)v.row_ids[r] : (int64_t)r;
            const float x = vals ? vals[v.bp[r] + k] : (v.values ? v.values[v.indptr[gr] + k] : 1.f);
            if (h0 == lane) bsum += scale * x;
            const float* d = dY + (int64_t)r * H;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int h = h0 + 32 * i;
                if (h < H) acc[i] += x * (d[h] * scale);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int h = h0 + 32 * i;
            if (h < H) dWt[(int64_t)j * H + h] += acc[i];
        }
    }
    if (db && lane == 0) db[j] += bsum;
}

static int launch_spmm_scatter_det(Ctx* c, const BatchView& v, const float* vals, float scale, const float* dY, int H,
                                   float* dWt, float* db, cudaStream_t s, int mod_n, int mod_r) {
    const int n_rows = c->det_rows;
    B200_CUDA_OK(cudaMemsetAsync(c->det_count, 0, (size_t)(n_rows + 1) * sizeof(int), s));
    B200_CUDA_OK(cudaMemsetAsync(c->det_cursor, 0, (size_t)(n_rows + 1) * sizeof(int), s));
    const int wgrid = (int)cdiv((int64_t)v.B * 32, 256);
    k_det_count<<<wgrid, 256, 0, s>>>(v, vals, n_rows, c->det_count);
    k_det_scan<<<1, 1024, 0, s>>>(c->det_count, n_rows, c->det_off);
    k_det_fill<<<wgrid, 256, 0, s>>>(v, vals, n_rows, c->det_off, c->det_cursor, c->det_ent);
    k_det_reduce<<<(int)cdiv((int64_t)n_rows * 32, 256), 256, 0, s>>>(v, vals, scale, dY, H, n_rows, c->det_off, c->det_ent,
                                                                   c->det_sorted, dWt, db, mod_n, mod_r);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_spmm_scatter(Ctx* c, const BatchView& v, const float* vals, float scale, const float* dY,
                        int H, float* dWt, cudaStream_t s, int mod_n, int mod_r) {
    return launch_spmm_scatter_bias(c, v, vals, scale, dY, H, dWt, nullptr, s, mod_n, mod_r);
}

int launch_spmm_scatter_bias(Ctx* c, const BatchView& v, const float* vals, float scale, const float* dY,
                             int H, float* dWt, float* db, cudaStream_t s, int mod_n, int mod_r) {
    if (v.B == 0) return 0;
    if (c->deterministic) return launch_spmm_scatter_det(c, v, vals, scale, dY, H, dWt, db, s, mod_n, mod_r);
    bool vec = (H % 4 == 0) && ((reinterpret_cast<uintptr_t>(dWt) & 15) == 0);
    if (mod_n > 1 && vec && !db && ((reinterpret_cast<uintptr_t>(dY) & 15) == 0)) {
        const int threads = (int)std::min<int64_t>(256, round_up(cdiv(H, 4), 32));
        B200_CUDA_OK(launch_pdl(k_spmm_scatter_owned, dim3(v.B), dim3(threads), 0, s, v, vals, scale, dY, H, dWt, mod_n, mod_r));
        note(c, "launch_spmm_scatter_owned", s);
        return 0;
    }
    const int grid = spmm_grid(c, v);
    if (vec) {
        int threads = (int)std::min<int64_t>(256, round_up(cdiv(H, 4), 32));
        B200_CUDA_OK(launch_pdl(k_spmm_scatter<4>, dim3(grid), dim3(threads), 0, s, v, vals, scale, dY, H, dWt, db, mod_n, mod_r));
    } else {
        int threads = (int)std::min<int64_t>(256, round_up(H, 32));
        B200_CUDA_OK(launch_pdl(k_spmm_scatter<1>, dim3(grid), dim3(threads), 0, s, v, vals, scale, dY, H, dWt, db, mod_n, mod_r));
    }
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// dense <-> CSR
// ------------------------------------------------------------------------------------------
__global__ void k_dense_count(const float* __restrict__ dense, int B, int I, int64_t* __restrict__ lens) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* row = dense + (int64_t)warp * I;
    int cnt = 0;
    for (int j = lane; j < I; j += 32) cnt += (row[j] != 0.f);
    cnt = (int)warp_sum((float)cnt);   // exact for counts < 2^24
    if (lane == 0) lens[warp] = cnt;
}

int launch_dense_count(Ctx* c, const float* dense, int B, int I, int64_t* lens, cudaStream_t s) {
    int threads = 256;
    k_dense_count<<<(int)cdiv((int64_t)B * 32, threads), threads, 0, s>>>(dense, B, I, lens);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void k_dense_fill(const float* __restrict__ dense, int B, int I,
                             const int64_t* __restrict__ indptr, int64_t cap,
                             int32_t* __restrict__ indices, float* __restrict__ values) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* row = dense + (int64_t)warp * I;
    int64_t o = indptr[warp];
    if (indptr[warp + 1] > cap) return;
    for (int j0 = 0; j0 < I; j0 += 32) {
        int j = j0 + lane;
        float x = (j < I) ? row[j] : 0.f;
        unsigned m = __ballot_sync(0xffffffffu, x != 0.f);
        if (x != 0.f) {
            int pos = __popc(m & ((1u << lane) - 1u));
            indices[o + pos] = j;
            values[o + pos] = x;
        }
        o += __popc(m);
    }
}

int launch_dense_fill(Ctx* c, const float* dense, int B, int I, const int64_t* indptr,
                      int32_t* indices, float* values, cudaStream_t s) {
    int threads = 256;
    k_dense_fill<<<(int)cdiv((int64_t)B * 32, threads), threads, 0, s>>>(
        dense, B, I, indptr, c->cfg.max_batch_nnz, indices, values);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void k_expand(BatchView v, int I, float* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= v.B) return;
    int64_t gr = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
    int64_t a = v.indptr[gr], b = v.indptr[gr + 1];
    for (int64_t k = a + lane; k < b; k += 32)
        out[(int64_t)warp * I + v.indices[k]] = v.values ? v.values[k] : 1.f;
}

int launch_expand(Ctx* c, const BatchView& v, int I, float* out, cudaStream_t s) {
    if (v.B == 0) return 0;
    B200_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)v.B * I * sizeof(float), s));
    int threads = 256;
    k_expand<<<(int)cdiv((int64_t)v.B * 32, threads), threads, 0, s>>>(v, I, out);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void k_mask_seen(BatchView v, int I, float* __restrict__ scores) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= v.B) return;
    int64_t gr = v.row_ids ? (int64_t)v.row_ids[warp] : (int64_t)warp;
    int64_t a = v.indptr[gr], b = v.indptr[gr + 1];
    for (int64_t k = a + lane; k < b; k += 32) {
        float x = v.values ? v.values[k] : 1.f;
        if (x != 0.f && v.indices[k] < I) scores[(int64_t)warp * I + v.indices[k]] = -INFINITY;   // cond columns are not items
    }
}

int launch_mask_seen(Ctx* c, const BatchView& v, int I, float* scores, cudaStream_t s) {
    if (v.B == 0) return 0;
    int threads = 256;
    k_mask_seen<<<(int)cdiv((int64_t)v.B * 32, threads), threads, 0, s>>>(v, I, scores);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// conditioned batches (ConditionedDataSampler.__iter__, samplers.py:187-232): example = (row, cond)
//   slot 0 internal batch: CSR-0 row + the condition as column n_items + cond (value 1)
//   slot 1 internal batch: CSR-1 row restricted to the items whose condition mask has bit `cond` (any bit if cond < 0)
// one warp per example; count -> scan -> fill
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool cond_pass(uint64_t m, int cond) {
    return cond < 0 ? (m != 0ull) : ((m >> cond) & 1ull) != 0ull;
}
__global__ void k_cond_count(const int64_t* __restrict__ ip0, const int64_t* __restrict__ ip1,
                             const int32_t* __restrict__ ix1, const int32_t* __restrict__ rows,
                             const int32_t* __restrict__ conds, int B, const uint64_t* __restrict__ mask,
                             int64_t* __restrict__ len0, int64_t* __restrict__ len1) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int64_t r = rows[warp];
    const int cond = conds[warp];
    int cnt = 0;
    for (int64_t k = ip1[r] + lane; k < ip1[r + 1]; k += 32) cnt += cond_pass(mask[ix1[k]], cond) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        len0[warp] = ip0[r + 1] - ip0[r] + (cond >= 0 ? 1 : 0);
        len1[warp] = cnt;
    }
}
__global__ void k_cond_fill(const int64_t* __restrict__ ip0, const int32_t* __restrict__ ix0, const float* __restrict__ v0,
                            const int64_t* __restrict__ ip1, const int32_t* __restrict__ ix1, const float* __restrict__ v1,
                            const int32_t* __restrict__ rows, const int32_t* __restrict__ conds, int B,
                            const uint64_t* __restrict__ mask, int n_items, int64_t cap,
                            const int64_t* __restrict__ op0, int32_t* __restrict__ ox0, float* __restrict__ ov0,
                            const int64_t* __restrict__ op1, int32_t* __restrict__ ox1, float* __restrict__ ov1) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    if (op0[B] > cap || op1[B] > cap) return;          // capacity overflow flagged by the scans
    const int64_t r = rows[warp];
    const int cond = conds[warp];
    const int64_t a0 = ip0[r], n0 = ip0[r + 1] - a0, o0 = op0[warp];
    for (int64_t k = lane; k < n0; k += 32) {
        ox0[o0 + k] = ix0[a0 + k];
        ov0[o0 + k] = v0 ? v0[a0 + k] : 1.f;
    }
    if (cond >= 0 && lane == 0) {                      // columns stay sorted: every item id < n_items + cond
        ox0[o0 + n0] = n_items + cond;
        ov0[o0 + n0] = 1.f;
    }
    // filtered target row, order preserved: warp-wide compaction 32 entries at a time
    const int64_t a1 = ip1[r], n1 = ip1[r + 1] - a1;
    int64_t w = op1[warp];
    for (int64_t k0 = 0; k0 < n1; k0 += 32) {
        const int64_t k = k0 + lane;
        bool keep = false;
        int32_t col = 0;
        if (k < n1) { col = ix1[a1 + k]; keep = cond_pass(mask[col], cond); }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int64_t pos = w + __popc(bal & ((1u << lane) - 1u));
            ox1[pos] = col;
            ov1[pos] = v1 ? v1[a1 + k] : 1.f;
        }
        w += __popc(bal);
    }
}

int launch_build_cond_batch(Ctx* c, const int32_t* ex_rows, const int32_t* ex_conds, int B,
                            const uint64_t* item_cond_mask, cudaStream_t s) {
    CsrSlot& S0 = c->slot[0];
    CsrSlot& S1 = c->slot[1];
    const int threads = 256;
    const int blocks = (int)cdiv((int64_t)B * 32, threads);
    int64_t* len0 = c->lens_tmp;
    int64_t* len1 = c->lens_tmp2;
    k_cond_count<<<blocks, threads, 0, s>>>(S0.indptr, S1.indptr, S1.indices, ex_rows, ex_conds, B, item_cond_mask, len0, len1);
    note(c, "cond_count", s);
    B200_CHECK(launch_scan_i64(c, len0, B, c->cfg.max_batch_nnz, S0.int_indptr, s));
    B200_CHECK(launch_scan_i64(c, len1, B, c->cfg.max_batch_nnz, S1.int_indptr, s));
    k_cond_fill<<<blocks, threads, 0, s>>>(S0.indptr, S0.indices, S0.values, S1.indptr, S1.indices, S1.values, ex_rows,
                                           ex_conds, B, item_cond_mask, c->n_items, c->cfg.max_batch_nnz,
                                           S0.int_indptr, S0.int_indices, S0.int_values,
                                           S1.int_indptr, S1.int_indices, S1.int_values);
    note(c, "cond_fill", s);
    B200_CUDA_OK(cudaGetLastError());
    S0.int_has_values = true;
    S1.int_has_values = true;
    return 0;
}

// ------------------------------------------------------------------------------------------
// row_loss: loss_r = T_r * lse_r - h_r . g_r - sum_k t_k * b[col_k]       (models.py:813)
// one warp per row
// ------------------------------------------------------------------------------------------
__global__ void k_row_loss(BatchView tgt, const float* __restrict__ h, const float* __restrict__ gvec,
                           int H, const float* __restrict__ bias, const float* __restrict__ pmax,
                           const float* __restrict__ psum, int n_tiles, float* __restrict__ lse,
                           const float* __restrict__ T, float inv_Bg, float* __restrict__ loss_row,
                           float* __restrict__ rowscale) {
    pdl_sync();
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= tgt.B) return;
    const int M = tgt.B;
    // merge of the per-tile (max, sum exp) partials of the fused decoder kernel -> lse_u
    float l;
    if (pmax) {
        float mx = -INFINITY;
        for (int t = lane; t < n_tiles; t += 32) mx = fmaxf(mx, pmax[(int64_t)t * M + warp]);
        mx = warp_max(mx);
        float sacc = 0.f;
        for (int t = lane; t < n_tiles; t += 32) {
            float pm = pmax[(int64_t)t * M + warp];
            if (pm != -INFINITY) sacc += psum[(int64_t)t * M + warp] * expf(pm - mx);
        }
        sacc = warp_sum(sacc);
        l = mx + logf(sacc);
        if (lane == 0) lse[warp] = l;
    } else {
        l = lse[warp];
    }
    float dot = 0.f, tb = 0.f;
    if (gvec) {     // fp32 path: sum_j t_j logit_j = h . g + sum_j t_j b_j
        for (int i = lane; i < H; i += 32) dot = fmaf(h[(int64_t)warp * H + i], gvec[(int64_t)warp * H + i], dot);
        int64_t gr = tgt.row_ids ? (int64_t)tgt.row_ids[warp] : (int64_t)warp;
        int64_t a = tgt.indptr[gr], b = tgt.indptr[gr + 1];
        for (int64_t k = a + lane; k < b; k += 32)
            tb = fmaf(tgt.values ? tgt.values[k] : 1.f, bias[tgt.indices[k]], tb);
        dot = warp_sum(dot);
        tb = warp_sum(tb);
    }
    if (lane == 0) {
        if (gvec) loss_row[warp] = T[warp] * l - dot - tb;
        if (rowscale) rowscale[warp] = T[warp] * inv_Bg;     // dlogits = softmax * T/B - t/B
    }
}

// Tensor-core path: after the recompute kernel stored P~^T[j,u] = softmax_uj * 2^S (fp16, the dense part of
// dlogits up to the per-user factor T_u/B that travels with the other GEMM operand), walk the TARGET non-zeros once:
//   loss_u   = -sum_j t_uj * log softmax_uj                 (== sum_j t_uj (lse_u - logit_uj), models.py:813)
//   P~^T[j,u] -= (t_uj / T_u) * 2^S                          (sparse part of dlogits: softmax*T/B - t/B =
//                                                             (T/B) * (softmax - t/T))
// so dW_d, db_d and dh come out of the two GEMMs complete: no gather of W_d rows for the loss and no
// scatter into dW_d.  One CTA per user row; ~nnz_B scattered 2-byte read-modify-writes.
// log softmax is read back from P~ where that is a normal fp16 number (>= 2^-14, i.e. softmax >= 3.7e-9: 11
// significant bits like every other tensor-core operand here); below that it is recomputed exactly from the
// operands, h_u . W_d[j] + b_j - lse_u, by the thread that found it (rare: a target item the model gives ~0).
__global__ void __launch_bounds__(128)
k_target_fixup(BatchView tgt, __half* __restrict__ PT, int64_t ldp, const float* __restrict__ T,
               const float* __restrict__ lse, const __half* __restrict__ h16, int64_t ldh,
               const __half* __restrict__ W16, int64_t ldw, const float* __restrict__ bias, int H,
               float log2_scale, float* __restrict__ loss_row, int* __restrict__ err, LossTail lt) {
    __shared__ float sh[4];
    __shared__ float s1[128], s2[128];
    __shared__ int s_last;
    pdl_sync();
    const int u = blockIdx.x;
    const float Tu = T[u];
    const float sub = (Tu != 0.f) ? exp2f(log2_scale) / Tu : 0.f;
    int64_t gr = tgt.row_ids ? (int64_t)tgt.row_ids[u] : (int64_t)u;
    int64_t a = tgt.indptr[gr], b = tgt.indptr[gr + 1];
    float acc = 0.f;
    for (int64_t k = a + threadIdx.x; k < b; k += blockDim.x) {
        const float t = tgt.values ? tgt.values[k] : 1.f;
        const int j = tgt.indices[k];
        __half* q = PT + (int64_t)j * ldp + u;
        const float p = __half2float(*q);
        float lsm;
        if (p >= 6.103515625e-05f) {
            lsm = (log2f(p) - log2_scale) * 0.6931471805599453f;
        } else {
            float dot = bias ? bias[j] : 0.f;
            const __half* hr = h16 + (int64_t)u * ldh;
            const __half* wr = W16 + (int64_t)j * ldw;
            for (int i = 0; i < H; ++i) dot = fmaf(__half2float(hr[i]), __half2float(wr[i]), dot);
            lsm = dot - lse[u];
        }
        acc = fmaf(t, lsm, acc);
        float nv = p - t * sub;
        if (!(fabsf(nv) <= 65504.f)) {      // |t / T_u| > ~4: not a multinomial target (mixed-sign ratings)
            atomicOr(err, 2);
            nv = fminf(fmaxf(nv, -65504.f), 65504.f);
        }
        *q = __float2half_rn(nv);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) loss_row[u] = -((sh[0] + sh[1]) + (sh[2] + sh[3]));
    if (!lt.loss_out) return;
    // the CTA that finishes last closes the loss (what k_loss_final does in its own launch): fixed-order tree over the
    // per-row terms, so the value does not depend on which CTA that is
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(lt.ticket, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float a1 = 0.f, a2 = 0.f;
    for (int i = threadIdx.x; i < lt.B; i += 128) {
        a1 += __ldcg(loss_row + i);
        if (lt.kl_row) a2 += __ldcg(lt.kl_row + i);
    }
    s1[threadIdx.x] = a1;
    s2[threadIdx.x] = a2;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float nll = s1[0] * lt.inv_Bg, kld = s2[0] * lt.inv_Bg;
        float reg = 0.f;
        if (lt.norms)
            for (int t = 0; t < lt.n_tensors; ++t) reg += lt.norms[t];
        lt.loss_out[0] = nll + lt.beta * kld + lt.lam * reg;
        lt.loss_out[1] = nll;
        lt.loss_out[2] = kld;
        lt.loss_out[3] = reg;
        *lt.ticket = 0;
    }
}

int launch_target_fixup(Ctx* c, const BatchView& tgt, __half* PT, int64_t ldp, const float* T, const float* lse,
                        const __half* h16, int64_t ldh, const __half* W16, int64_t ldw, const float* bias, int H,
                        float* loss_row, cudaStream_t s, const LossTail* tail) {
    if (tgt.B == 0) return 0;
    LossTail lt = {};
    if (tail) lt = *tail;
    B200_CUDA_OK(launch_pdl(k_target_fixup, dim3(tgt.B), dim3(128), 0, s, tgt, PT, ldp, T, lse, h16, ldh, W16, ldw, bias, H,
                            PROB_LOG2_SCALE, loss_row, c->d_err, lt));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_row_loss(Ctx* c, const BatchView& tgt, const float* h, const float* gvec, int H,
                    const float* bias, const float* pmax, const float* psum, int n_tiles, float* lse,
                    const float* T, float inv_Bg, float* loss_row, float* rowscale, cudaStream_t s) {
    if (tgt.B == 0) return 0;
    int threads = 256;
    B200_CUDA_OK(launch_pdl(k_row_loss, dim3((unsigned)cdiv((int64_t)tgt.B * 32, threads)), dim3(threads), 0, s, tgt, h, gvec,
                            H, bias, pmax, psum, n_tiles, lse, T, inv_Bg, loss_row, rowscale));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200
