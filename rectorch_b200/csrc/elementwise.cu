// elementwise.cu -- reparameterisation + KL, loss assembly, parameter norms and the fused Adam.
#include <cuda_fp16.h>
#include <algorithm>
#include "ctx.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------
// K3b: z = mu + eps * exp(0.5*logvar) (train) | mu (eval)            nets.py:317-320, 407-411
//      kl_row[r] = -0.5 * sum_l (1 + logvar - mu^2 - exp(logvar))    models.py:814
// enc_out [B x 2L] = [mu | logvar];  one warp per row.
// eps: tape (parity mode) or Philox + Box-Muller keyed by the GLOBAL row id, so the draw
// does not depend on how rows are sharded over ranks.
// ------------------------------------------------------------------------------------------
__global__ void k_reparam_kl(const float* __restrict__ enc_out, int B, int L, int train,
                             const float* __restrict__ eps_tape, uint64_t seed, uint64_t step,
                             int64_t row_offset, const int32_t* __restrict__ row_ids,
                             float* __restrict__ z, float* __restrict__ eps_out,
                             float* __restrict__ kl_row, __half* __restrict__ z16, int64_t ldz16) {
    pdl_sync();
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const float* mu = enc_out + (int64_t)warp * 2 * L;
    const float* lv = mu + L;
    uint64_t grow = (uint64_t)((row_ids ? (int64_t)row_ids[warp] : (int64_t)warp) + row_offset);
    float kl = 0.f;
    for (int l = lane; l < L; l += 32) {
        float m_ = mu[l], v_ = lv[l];
        float ev = expf(v_);
        kl += 1.f + v_ - m_ * m_ - ev;
        float zz = m_;
        if (train) {
            float e;
            if (eps_tape) {
                e = eps_tape[(int64_t)warp * L + l];
            } else {
                uint4 ctr = make_uint4((uint32_t)grow, (uint32_t)(grow >> 32), (uint32_t)l, (uint32_t)step);
                uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0xE9E9u);
                uint4 r = philox4x32(ctr, key);
                float u1 = u32_to_unit(r.x), u2 = u32_to_unit(r.y);
                e = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
            }
            eps_out[(int64_t)warp * L + l] = e;
            zz = m_ + e * expf(0.5f * v_);
        }
        z[(int64_t)warp * L + l] = zz;
        if (z16) z16[(int64_t)warp * ldz16 + l] = __float2half_rn(f16_clamp(zz));
    }
    kl = warp_sum(kl);
    if (lane == 0) kl_row[warp] = -0.5f * kl;
}

int launch_reparam_kl(Ctx* c, const float* enc_out, int B, int L, bool train, const float* eps_tape,
                      uint64_t seed, uint64_t step, int64_t row_offset, const int32_t* row_ids,
                      float* z, float* eps_out, float* kl_row, __half* z16, int64_t ldz16, cudaStream_t s) {
    if (B == 0) return 0;
    int threads = 256;
    B200_CUDA_OK(launch_pdl(k_reparam_kl, dim3((unsigned)cdiv((int64_t)B * 32, threads)), dim3(threads), 0, s,
                            enc_out, B, L, train ? 1 : 0, eps_tape, seed, step, row_offset, row_ids, z, eps_out, kl_row, z16,
                            ldz16));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// d(enc_out) from dz:  dmu = dz + beta*mu/B ;  dlogvar = dz*eps*0.5*std + beta*0.5*(exp(lv)-1)/B
__global__ void k_dz_to_denc(const float* __restrict__ dz, const float* __restrict__ enc_out,
                             const float* __restrict__ eps, int B, int L, float beta_over_B, int train,
                             float* __restrict__ denc, __half* __restrict__ denc16, int64_t ld16) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * L) return;
    int r = (int)(i / L), l = (int)(i % L);
    float mu = enc_out[(int64_t)r * 2 * L + l], lv = enc_out[(int64_t)r * 2 * L + L + l];
    float d = dz[i];
    float dlv = beta_over_B * 0.5f * (expf(lv) - 1.f);
    if (train) dlv += d * eps[i] * 0.5f * expf(0.5f * lv);
    const float dmu = d + beta_over_B * mu;
    denc[(int64_t)r * 2 * L + l] = dmu;
    denc[(int64_t)r * 2 * L + L + l] = dlv;
    if (denc16) {
        denc16[(int64_t)r * ld16 + l] = __float2half_rn(f16_clamp(dmu));
        denc16[(int64_t)r * ld16 + L + l] = __float2half_rn(f16_clamp(dlv));
    }
}

int launch_dz_to_denc(Ctx* c, const float* dz, const float* enc_out, const float* eps, int B, int L,
                      float beta_over_B, bool train, float* denc, __half* denc16, int64_t ld16, cudaStream_t s) {
    int64_t n = (int64_t)B * L;
    if (n == 0) return 0;
    B200_CUDA_OK(launch_pdl(k_dz_to_denc, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, s, dz, enc_out, eps, B, L, beta_over_B,
                            train ? 1 : 0, denc, denc16, ld16));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// loss_out = {total, nll, kld, reg}; nll = sum(loss_row)/Bg; kld = sum(kl_row)/Bg; reg = sum norms.
// Single CTA, fixed-order tree reduction -> deterministic.
__global__ void k_loss_final(const float* __restrict__ loss_row, const float* __restrict__ kl_row, int B,
                             float inv_Bg, float beta, float lam, const float* __restrict__ norms,
                             int n_tensors, float* __restrict__ loss_out) {
    __shared__ float s1[256], s2[256];
    pdl_sync();
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) {
        a += loss_row[i];
        if (kl_row) b += kl_row[i];
    }
    s1[threadIdx.x] = a;
    s2[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float nll = s1[0] * inv_Bg, kld = s2[0] * inv_Bg, reg = 0.f;
        if (norms)
            for (int t = 0; t < n_tensors; ++t) reg += norms[t];
        loss_out[0] = nll + beta * kld + lam * reg;
        loss_out[1] = nll;
        loss_out[2] = kld;
        loss_out[3] = reg;
    }
}

int launch_loss_final(Ctx* c, const float* loss_row, const float* kl_row, int B, float inv_Bg,
                      float beta, float lam, const float* norms, int n_tensors, float* loss_out,
                      cudaStream_t s) {
    B200_CUDA_OK(launch_pdl(k_loss_final, dim3(1), dim3(256), 0, s, loss_row, kl_row, B, inv_Bg, beta, lam, norms, n_tensors,
                            loss_out));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// per-tensor L2 norms (MultiDAE regulariser, models.py:702-704): two fixed-order levels
// ------------------------------------------------------------------------------------------
__global__ void k_norm_partial(const float* __restrict__ w, const int64_t* __restrict__ offs,
                               const int64_t* __restrict__ lens, float* __restrict__ partial) {
    __shared__ float sh[256];
    int t = blockIdx.y;
    const float* p = w + offs[t];
    int64_t n = lens[t];
    int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(n, lo + chunk);
    float acc = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc = fmaf(p[i], p[i], acc);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[t * gridDim.x + blockIdx.x] = sh[0];
}
__global__ void k_norm_final(const float* __restrict__ partial, int n_parts, float* __restrict__ norms) {
    int t = blockIdx.x;
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int i = 0; i < n_parts; ++i) acc += partial[t * n_parts + i];
        norms[t] = sqrtf(acc);
    }
}

int launch_tensor_norms(Ctx* c, const float* w, const int64_t* offs, const int64_t* lens, int n_tensors,
                        float* partial, float* norms, cudaStream_t s) {
    dim3 grid(NORM_PARTS, n_tensors);     // enough CTAs on the two item-sized tensors to stream at HBM rate
    k_norm_partial<<<grid, 256, 0, s>>>(w, offs, lens, partial);
    k_norm_final<<<n_tensors, 32, 0, s>>>(partial, NORM_PARTS, norms);
    note(c, __func__, s); c->launches++;
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// K8 fused Adam: one streaming pass, 16 B/param read (w,g,m,v) + 12 B/param written (w,m,v).
// Arithmetic follows torch/optim/adam.py::_single_tensor_adam + the ATen CPU kernels:
//   g  = g + wd*w (+ lam*w/||w||  -- MultiDAE regulariser gradient)
//   m  = fma(1-b1, g-m, m)                       (Tensor.lerp_)
//   v  = v*b2 + ((1-b2)*g)*g                     (mul_ ; addcmul_)
//   w  = w + ((-lr/bc1)*m) / (sqrt(v)/sqrt(bc2) + eps)       (addcdiv_)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_one(float& w, float g, float& m, float& v, float neg_step,
                                         float b1c, float b2, float b2c, float bc2_sqrt, float eps,
                                         float wd, float reg) {
    g = __fadd_rn(g, __fadd_rn(__fmul_rn(wd, w), __fmul_rn(reg, w)));
    m = fmaf(b1c, g - m, m);
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(b2c, g), g));
    float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), eps);
    w = __fadd_rn(w, __fdiv_rn(__fmul_rn(neg_step, m), denom));
}

// FILTER (see AdamOpt): rows of the re-zero window [z_lo, z_hi) are selected by their step mark; elements outside
// the window are always processed.  z_lo and row_len are multiples of 4 when FILTER != 0, so a float4 never
// straddles two rows.
template <bool EXTRAS, int FILTER>
__global__ void __launch_bounds__(256)
k_adam(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
       int64_t n, float neg_step, float b1c, float b2, float b2c, float bc2_sqrt, float eps, float wd,
       float lam, const float* __restrict__ norm_ptr, __half* __restrict__ shadow, int64_t sh_lo, int64_t sh_hi,
       int64_t z_lo, int64_t z_hi, const int32_t* __restrict__ mark, int32_t mark_step, int row_len,
       __half* __restrict__ shadow2, int64_t s2_lo, int64_t s2_hi, AdamW1 w1) {
    pdl_sync();
    float reg = 0.f;
    if (EXTRAS && norm_ptr) {
        float nrm = *norm_ptr;
        reg = lam / nrm;
    }
    int64_t n4 = n >> 2;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    float4* w4 = reinterpret_cast<float4*>(w);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t e0 = i << 2;
        const bool in_z = e0 >= z_lo && e0 < z_hi;
        bool zero_g = false;
        int64_t w1_pos = -1;            // ADAM_ROWS_MOD: where this float4 goes in the gathered encoder-0 weight
        if (FILTER == ADAM_ROWS_MOD) {
            if (in_z) {
                const int64_t row = (e0 - z_lo) / row_len;
                if ((int)(row % w1.mod_n) != w1.mod_r) continue;      // another rank's row
                w1_pos = ((int64_t)w1.mod_r * w1.rows_per + row / w1.mod_n) * row_len + ((e0 - z_lo) - row * row_len);
            }
        } else if (FILTER != ADAM_ROWS_ALL && in_z) {
            const bool marked = mark[(e0 - z_lo) / row_len] == mark_step;
            if (FILTER == ADAM_ROWS_MARKED && !marked) continue;
            if (FILTER == ADAM_ROWS_UNMARKED) {
                if (marked) continue;
                zero_g = true;          // no scatter reached this row: its gradient is +0 and stays in memory as such
            }
        }
        float4 ww, mm, vv;
        if (w1.stream) { ww = __ldcs(w4 + i); mm = __ldcs(m4 + i); vv = __ldcs(v4 + i); }
        else           { ww = w4[i]; mm = m4[i]; vv = v4[i]; }
        float4 gg = zero_g ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcs(g4 + i);
        adam_one(ww.x, gg.x, mm.x, vv.x, neg_step, b1c, b2, b2c, bc2_sqrt, eps, EXTRAS ? wd : 0.f, reg);
        adam_one(ww.y, gg.y, mm.y, vv.y, neg_step, b1c, b2, b2c, bc2_sqrt, eps, EXTRAS ? wd : 0.f, reg);
        adam_one(ww.z, gg.z, mm.z, vv.z, neg_step, b1c, b2, b2c, bc2_sqrt, eps, EXTRAS ? wd : 0.f, reg);
        adam_one(ww.w, gg.w, mm.w, vv.w, neg_step, b1c, b2, b2c, bc2_sqrt, eps, EXTRAS ? wd : 0.f, reg);
        if (w1.stream) { __stcs(w4 + i, ww); __stcs(m4 + i, mm); __stcs(v4 + i, vv); }
        else           { w4[i] = ww; m4[i] = mm; v4[i] = vv; }
        if (FILTER == ADAM_ROWS_MOD && w1_pos >= 0) {       // fp16 image of the updated row, into this rank's block
            const __half2 lo = __floats2half2_rn(f16_clamp(ww.x), f16_clamp(ww.y));
            const __half2 hi = __floats2half2_rn(f16_clamp(ww.z), f16_clamp(ww.w));
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(w1.w1g + w1_pos) = pk;
        }
        // the encoder-0 gradient is accumulated by sparse scatters into an all-zero buffer: restore the zeros
        // here, touching only the (few) rows that actually received a gradient
        if (in_z && !zero_g && (gg.x != 0.f || gg.y != 0.f || gg.z != 0.f || gg.w != 0.f))
            g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (shadow && e0 >= sh_lo && e0 < sh_hi) {   // fp16 image of W_d for the tensor-core GEMMs
            const __half2 lo = __floats2half2_rn(f16_clamp(ww.x), f16_clamp(ww.y));
            const __half2 hi = __floats2half2_rn(f16_clamp(ww.z), f16_clamp(ww.w));
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(shadow + (e0 - sh_lo)) = pk;
        }
        if (shadow2 && e0 >= s2_lo && e0 < s2_hi) {  // fp16 image of the hidden-layer tensors
            const __half2 lo = __floats2half2_rn(f16_clamp(ww.x), f16_clamp(ww.y));
            const __half2 hi = __floats2half2_rn(f16_clamp(ww.z), f16_clamp(ww.w));
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(shadow2 + (e0 - s2_lo)) = pk;
        }
    }
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool in_z = i >= z_lo && i < z_hi;
        bool zero_g = false;
        if (FILTER == ADAM_ROWS_MOD) {
            if (in_z) continue;         // rows are whole float4s (row_len % 4 == 0): the window has no scalar tail
        } else if (FILTER != ADAM_ROWS_ALL && in_z) {
            const bool marked = mark[(i - z_lo) / row_len] == mark_step;
            if (FILTER == ADAM_ROWS_MARKED && !marked) continue;
            if (FILTER == ADAM_ROWS_UNMARKED) {
                if (marked) continue;
                zero_g = true;
            }
        }
        float ww = w[i], mm = m[i], vv = v[i];
        const float gi = zero_g ? 0.f : g[i];
        adam_one(ww, gi, mm, vv, neg_step, b1c, b2, b2c, bc2_sqrt, eps, EXTRAS ? wd : 0.f, reg);
        if (in_z && gi != 0.f) g[i] = 0.f;
        w[i] = ww;
        m[i] = mm;
        v[i] = vv;
        if (shadow && i >= sh_lo && i < sh_hi) shadow[i - sh_lo] = __float2half_rn(f16_clamp(ww));
        if (shadow2 && i >= s2_lo && i < s2_hi) shadow2[i - s2_lo] = __float2half_rn(f16_clamp(ww));
    }
}

int launch_adam(Ctx* c, float* w, float* g, float* m, float* v, int64_t n, float lr_over_bc1,
                float beta1, float beta2, float bc2_sqrt, float eps, float wd, float lam,
                const float* norm_ptr, __half* shadow, int64_t sh_lo, int64_t sh_hi, int64_t z_lo, int64_t z_hi,
                const AdamOpt& opt, cudaStream_t s) {
    // second image window (hidden-layer tensors): absolute arena positions translated to this launch's base
    __half* shadow2 = opt.shadow2;
    const int64_t s2_lo = opt.s2_lo, s2_hi = opt.s2_hi;
    if (n == 0) return 0;
    B200_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(m) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
                 B200VAE_EINVAL, "adam: arenas must be 16-byte aligned");
    const int filter = (z_hi > z_lo) ? opt.filter : ADAM_ROWS_ALL;
    if (filter != ADAM_ROWS_ALL)
        B200_REQUIRE((opt.mark || filter == ADAM_ROWS_MOD) && opt.row_len > 0 && opt.row_len % 4 == 0 && z_lo % 4 == 0 &&
                     (z_hi - z_lo) % 4 == 0, B200VAE_EINVAL, "adam: the row filter needs step marks and 16-byte aligned rows");
    if (filter == ADAM_ROWS_MOD)
        B200_REQUIRE(opt.w1g && opt.mod_n > 1 && opt.mod_r >= 0 && opt.mod_r < opt.mod_n && opt.rows_per > 0, B200VAE_EINVAL,
                     "adam: bad encoder-0 shard description");
    const AdamW1 w1 = {opt.w1g, opt.mod_n, opt.mod_r, opt.rows_per, opt.stream ? 1 : 0};
    // persistent-style grid: a multiple of the SM count (default 8 CTAs of 256 threads per SM)
    const int threads = std::min(256, std::max(32, opt.threads & ~31));
    int64_t want = cdiv(std::max<int64_t>(n >> 2, 1), threads);
    int blocks = (int)std::min<int64_t>(want, (int64_t)c->num_sms * std::max(1, opt.ctas_per_sm));
    bool extras = (wd != 0.f) || (lam != 0.f && norm_ptr);
    float b1c = 1.f - beta1, b2c = 1.f - beta2;
#define ADAM_LAUNCH(EX, FI)                                                                                               \
    B200_CUDA_OK(launch_pdl(k_adam<EX, FI>, dim3(blocks), dim3(threads), 0, s, w, g, m, v, n, -lr_over_bc1, b1c, beta2, b2c,   \
                            bc2_sqrt, eps, EX ? wd : 0.f, EX ? lam : 0.f, EX ? norm_ptr : (const float*)nullptr, shadow,  \
                            sh_lo, sh_hi, z_lo, z_hi, opt.mark, opt.mark_step, opt.row_len, shadow2, s2_lo, s2_hi, w1))
    if (extras) {
        if (filter == ADAM_ROWS_MARKED) ADAM_LAUNCH(true, ADAM_ROWS_MARKED);
        else if (filter == ADAM_ROWS_UNMARKED) ADAM_LAUNCH(true, ADAM_ROWS_UNMARKED);
        else if (filter == ADAM_ROWS_MOD) ADAM_LAUNCH(true, ADAM_ROWS_MOD);
        else ADAM_LAUNCH(true, ADAM_ROWS_ALL);
    } else {
        if (filter == ADAM_ROWS_MARKED) ADAM_LAUNCH(false, ADAM_ROWS_MARKED);
        else if (filter == ADAM_ROWS_UNMARKED) ADAM_LAUNCH(false, ADAM_ROWS_UNMARKED);
        else if (filter == ADAM_ROWS_MOD) ADAM_LAUNCH(false, ADAM_ROWS_MOD);
        else ADAM_LAUNCH(false, ADAM_ROWS_ALL);
    }
#undef ADAM_LAUNCH
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// Drop the L2 lines of a consumed buffer without writing them back (discard.global.L2): the decoder-output gradient is
// produced chunk by chunk, read once by Adam while still in L2, and overwritten next step -- it never needs to reach HBM.
__global__ void k_discard_l2(const char* __restrict__ p, int64_t n_lines) {
    pdl_sync();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_lines; i += stride) asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + i * 128) : "memory");
}
int launch_discard_l2(Ctx* c, const float* p, int64_t n, cudaStream_t s) {
    // whole 128-byte lines inside [p, p + n)
    const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127;
    const uintptr_t b = reinterpret_cast<uintptr_t>(p + n) & ~(uintptr_t)127;
    if (b <= a) return 0;
    const int64_t n_lines = (int64_t)((b - a) >> 7);
    const int blocks = (int)std::min<int64_t>(cdiv(n_lines, 256), (int64_t)c->num_sms * 4);
    B200_CUDA_OK(launch_pdl(k_discard_l2, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const char*>(a), n_lines));
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

// y[r * ldy + c] = fp16(x[r * cols + c]) (round to nearest, clamped to the finite fp16 range): operand images for
// the tensor cores (W_d after anything but Adam wrote it; the last hidden activation in predict)
__global__ void k_to_f16(const float* __restrict__ x, __half* __restrict__ y, int64_t rows, int cols, int64_t ldy) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n = rows * cols;
    if (ldy == cols) {
        for (; i < n; i += stride) y[i] = __float2half_rn(f16_clamp(x[i]));
    } else {
        for (; i < n; i += stride) y[(i / cols) * ldy + (i % cols)] = __float2half_rn(f16_clamp(x[i]));
    }
}
int launch_to_f16(Ctx* c, const float* x, __half* y, int64_t rows, int cols, int64_t ldy, cudaStream_t s) {
    const int64_t n = rows * cols;
    if (n == 0) return 0;
    int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)c->num_sms * 8);
    k_to_f16<<<blocks, 256, 0, s>>>(x, y, rows, cols, ldy);
    note(c, __func__, s);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace b200

// ------------------------------------------------------------------------------------------
// dense-API helpers for MultiVAE/MultiDAE.loss_function(recon_x, x, ...) called with explicit
// [B x n_items] tensors (models.py:701, 813-814): row-wise -sum(log_softmax(l) * x) and KL rows.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_nll_rows(const float* __restrict__ logits, const float* __restrict__ target, int I, float* __restrict__ out) {
    __shared__ float sh[8];
    __shared__ float s_bcast;
    const float* l = logits + (int64_t)blockIdx.x * I;
    const float* t = target + (int64_t)blockIdx.x * I;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < I; j += 256) mx = fmaxf(mx, l[j]);
    mx = b200::warp_max(mx);
    if (lane == 0) sh[wid] = mx;
    __syncthreads();
    if (threadIdx.x == 0) { float m = sh[0]; for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]); s_bcast = m; }
    __syncthreads();
    mx = s_bcast;
    float se = 0.f, tl = 0.f, ts = 0.f;
    for (int j = threadIdx.x; j < I; j += 256) {
        se += expf(l[j] - mx);
        tl = fmaf(t[j], l[j], tl);
        ts += t[j];
    }
    __syncthreads();
    se = b200::warp_sum(se); if (lane == 0) sh[wid] = se; __syncthreads();
    if (threadIdx.x == 0) { float a = 0.f; for (int i = 0; i < 8; ++i) a += sh[i]; s_bcast = a; }
    __syncthreads(); se = s_bcast; __syncthreads();
    tl = b200::warp_sum(tl); if (lane == 0) sh[wid] = tl; __syncthreads();
    if (threadIdx.x == 0) { float a = 0.f; for (int i = 0; i < 8; ++i) a += sh[i]; s_bcast = a; }
    __syncthreads(); tl = s_bcast; __syncthreads();
    ts = b200::warp_sum(ts); if (lane == 0) sh[wid] = ts; __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f; for (int i = 0; i < 8; ++i) a += sh[i];
        out[blockIdx.x] = a * (mx + logf(se)) - tl;     // sum_j t_j (lse - l_j)
    }
}
__global__ void k_kl_rows(const float* __restrict__ mu, const float* __restrict__ lv, int B, int L, float* __restrict__ out) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    float kl = 0.f;
    for (int l = lane; l < L; l += 32) {
        float m_ = mu[(int64_t)warp * L + l], v_ = lv[(int64_t)warp * L + l];
        kl += 1.f + v_ - m_ * m_ - expf(v_);
    }
    kl = b200::warp_sum(kl);
    if (lane == 0) out[warp] = -0.5f * kl;
}

extern "C" int b200vae_multinomial_nll_rows(const float* logits, const float* target, int32_t B, int32_t n_items,
                                            float* out_rows, void* stream) {
    B200_REQUIRE(logits && target && out_rows && B >= 1 && n_items >= 1, B200VAE_EINVAL, "bad argument");
    k_nll_rows<<<B, 256, 0, (cudaStream_t)stream>>>(logits, target, n_items, out_rows);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" int b200vae_kl_rows(const float* mu, const float* logvar, int32_t B, int32_t L, float* out_rows, void* stream) {
    B200_REQUIRE(mu && logvar && out_rows && B >= 1 && L >= 1, B200VAE_EINVAL, "bad argument");
    k_kl_rows<<<(int)b200::cdiv((int64_t)B * 32, 256), 256, 0, (cudaStream_t)stream>>>(mu, logvar, B, L, out_rows);
    B200_CUDA_OK(cudaGetLastError());
    return 0;
}
