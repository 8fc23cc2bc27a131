// K3/K4/K5  HBM-bound vector kernels of the data-consistency step.
//
//   cg_update_xr : alpha = rr/pd;  x += alpha p;  r -= alpha d;  partial ||r||^2
//                  (reference src/utils/cg.py:29-33; the direction update p = r + beta p of :35-38 is
//                  fused into the pack pass of the projector, fp_march.cu)
//   tweedie_rhs  : xhat0 = (x - s*std_t)/mean_t;  b = xhat0 + gamma*atb
//                  (reference src/samplers/utils.py:370-378 and :197)
//   ddim         : DDPM branch of ddim() (reference src/samplers/utils.py:356-368)
//
// Layout: grid = (blocks per sample, batch); every block streams a contiguous
// slice of one sample with 128-bit accesses when the sample size allows it.
// Per-sample scalars (alpha, beta, schedule coefficients) are recomputed by
// each block from device-resident partial sums / the alpha-bar table: nothing
// goes through the host, so the whole step is CUDA-graph capturable and the
// per-sample time step may differ between samples.
//
// tweedie_rhs and ddim use explicit round-to-nearest intrinsics in the
// reference's operation order (no FMA contraction) so that, given the same
// inputs, they reproduce the eager PyTorch arithmetic of the reference bit for
// bit.
#include "scd_internal.cuh"

#define VEC_THREADS 256
#define VEC_U 4          /* float4 per thread and array in flight: the reads of a batch are issued together, then its writes */

int scd_vec_blocks_per_sample(int64_t numel, int batch)
{
    // >= ~4 CTAs per SM over the whole batch (148 SMs), at least one float4 per thread,
    // at most 64 blocks per sample (the consumers sum the per-block partials serially)
    int64_t nb = (4 * 148 + batch - 1) / (batch > 0 ? batch : 1);
    const int64_t cap = (numel + 4 * VEC_THREADS - 1) / (4 * VEC_THREADS);
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    if (nb > 64) nb = 64;
    return (int)nb;
}

__device__ __forceinline__ float block_sum(float v, float *red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (VEC_THREADS >> 5)) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    }
    return t;   // valid in thread 0 (whole warp 0)
}

// Sum n partials (n <= a few hundred) in a fixed order; result broadcast via smem.
__device__ __forceinline__ float sum_partials(const float *part, int n, float *slot)
{
    if (threadIdx.x < 32) {
        float v = 0.f;
        for (int i = threadIdx.x; i < n; i += 32) v += part[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0) *slot = v;
    }
    __syncthreads();
    return *slot;
}

__device__ __forceinline__ void slice_of(int64_t numel, int64_t &lo, int64_t &hi)
{
    // contiguous slice of this block, aligned to 4 elements
    const int64_t per = ((numel + gridDim.x - 1) / gridDim.x + 3) & ~(int64_t)3;
    lo = (int64_t)blockIdx.x * per;
    hi = lo + per;
    if (hi > numel) hi = numel;
    if (lo > numel) lo = numel;
}

template <bool VEC4>
__global__ void __launch_bounds__(VEC_THREADS)
cg_update_xr_kernel(const float *x_in, float *x, float *__restrict__ r,
                    const float *__restrict__ p, const float *__restrict__ d,
                    const float *__restrict__ rr_part, int rr_n,
                    const float *__restrict__ pd_part, int pd_n, int part_stride,
                    float *__restrict__ rr_new_part, int64_t numel)
{
    scd_pdl_wait();                               // predecessor complete, its writes visible
    scd_pdl_trigger();
    __shared__ float red[VEC_THREADS / 32];
    __shared__ float sc[2];
    const int b = blockIdx.y;
    int64_t lo, hi;
    slice_of(numel, lo, hi);
    const size_t base = (size_t)b * numel;
    const float4 *xi4 = reinterpret_cast<const float4 *>(x_in + base);
    float4 *x4 = reinterpret_cast<float4 *>(x + base);
    float4 *r4 = reinterpret_cast<float4 *>(r + base);
    const float4 *p4 = reinterpret_cast<const float4 *>(p + base);
    const float4 *d4 = reinterpret_cast<const float4 *>(d + base);
    // the first batch of vector loads goes out before the scalar phase (its latency hides theirs)
    float4 xv[VEC_U], rv[VEC_U], pv[VEC_U], dv[VEC_U];
    const int64_t i_end = hi >> 2;
    int64_t i0 = (lo >> 2) + threadIdx.x;
    auto load_batch = [&](int64_t i) {
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            const int64_t k = i + (int64_t)u * VEC_THREADS;
            if (k < i_end) { xv[u] = xi4[k]; rv[u] = r4[k]; pv[u] = p4[k]; dv[u] = d4[k]; }
        }
    };
    if (VEC4) load_batch(i0);
    // ||r||^2 and <p,d> from their per-block partials: warp 0 and warp 1 add them concurrently (each in
    // the fixed lane-strided + butterfly order of sum_partials), one barrier for both
    if (threadIdx.x < 64) {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const float *part = (w == 0 ? rr_part : pd_part) + (size_t)b * part_stride;
        const int n = w == 0 ? rr_n : pd_n;
        float v = 0.f;
        for (int i = lane; i < n; i += 32) v += part[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sc[w] = v;
    }
    __syncthreads();
    const float rr = sc[0], pd = sc[1];
    const float alpha = __fdiv_rn(rr, pd);      // no guard: same as the reference
    float acc = 0.f;
    if (VEC4) {
        // per thread the elements are visited in the same order as a VEC_THREADS-strided loop, so the
        // partial sums do not depend on VEC_U
        for (int64_t i = i0; i < i_end; ) {
#pragma unroll
            for (int u = 0; u < VEC_U; ++u) {
                const int64_t k = i + (int64_t)u * VEC_THREADS;
                if (k < i_end) {
                    float4 xo = xv[u], ro = rv[u];
                    xo.x = fmaf(alpha, pv[u].x, xo.x); xo.y = fmaf(alpha, pv[u].y, xo.y);
                    xo.z = fmaf(alpha, pv[u].z, xo.z); xo.w = fmaf(alpha, pv[u].w, xo.w);
                    ro.x = fmaf(-alpha, dv[u].x, ro.x); ro.y = fmaf(-alpha, dv[u].y, ro.y);
                    ro.z = fmaf(-alpha, dv[u].z, ro.z); ro.w = fmaf(-alpha, dv[u].w, ro.w);
                    x4[k] = xo; r4[k] = ro;
                    acc = fmaf(ro.x, ro.x, acc); acc = fmaf(ro.y, ro.y, acc);
                    acc = fmaf(ro.z, ro.z, acc); acc = fmaf(ro.w, ro.w, acc);
                }
            }
            i += (int64_t)VEC_U * VEC_THREADS;
            if (i < i_end) load_batch(i);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += VEC_THREADS) {
            const float xv = fmaf(alpha, p[base + i], x_in[base + i]);
            const float rv = fmaf(-alpha, d[base + i], r[base + i]);
            x[base + i] = xv; r[base + i] = rv;
            acc = fmaf(rv, rv, acc);
        }
    }
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) rr_new_part[(size_t)b * part_stride + blockIdx.x] = tot;
}

// ---- schedule look-up: abar[t+1] exactly like DDPM._compute_alpha_cumprod ----
__device__ __forceinline__ float abar_at(const float *abar, int n_table, float t)
{
    long long idx = (long long)t + 1;           // Tensor.long() truncates toward zero
    if (idx < 0) idx = 0;
    if (idx >= n_table) idx = n_table - 1;
    return abar[idx];
}

template <bool VEC4>
__global__ void __launch_bounds__(VEC_THREADS)
tweedie_rhs_kernel(const float *__restrict__ x, const float *__restrict__ s,
                   const float *__restrict__ atb, const float *__restrict__ t,
                   const float *__restrict__ abar, int n_table, float gamma,
                   float *__restrict__ xhat0, float *__restrict__ bvec, int64_t numel)
{
    scd_pdl_wait();                               // predecessor complete, its writes visible
    scd_pdl_trigger();
    const int b = blockIdx.y;
    int64_t lo, hi;
    slice_of(numel, lo, hi);
    const size_t base = (size_t)b * numel;
    const float4 *x4 = reinterpret_cast<const float4 *>(x + base);
    const float4 *s4 = reinterpret_cast<const float4 *>(s + base);
    const float4 *a4 = reinterpret_cast<const float4 *>(atb + base);
    float4 xv[VEC_U], sv[VEC_U], av[VEC_U];
    const int64_t i_end = hi >> 2;
    const int64_t i0 = (lo >> 2) + threadIdx.x;
    auto load_batch = [&](int64_t i) {
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            const int64_t k = i + (int64_t)u * VEC_THREADS;
            av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < i_end) { xv[u] = x4[k]; sv[u] = s4[k]; if (bvec) av[u] = a4[k]; }
        }
    };
    if (VEC4) load_batch(i0);                     // in flight while the schedule look-up resolves
    const float ab = abar_at(abar, n_table, t[b]);
    const float mean = __fsqrt_rn(ab);                       // bar_a.pow(.5)
    const float stdv = __fsqrt_rn(__fsub_rn(1.0f, ab));      // (1 - bar_a).pow(.5)
    const float div = __fdiv_rn(1.0f, mean);                 // mean.pow(-1)
#define TW_ONE(X, S_, A, XH, BV)                                                   \
    {                                                                              \
        const float u__ = __fsub_rn((X), __fmul_rn((S_), stdv));                   \
        (XH) = __fmul_rn(u__, div);                                                \
        (BV) = __fadd_rn((XH), __fmul_rn(gamma, (A)));                             \
    }
    if (VEC4) {
        float4 *h4 = reinterpret_cast<float4 *>(xhat0 + base);
        float4 *b4 = bvec ? reinterpret_cast<float4 *>(bvec + base) : nullptr;
        for (int64_t i = i0; i < i_end; ) {
#pragma unroll
            for (int u = 0; u < VEC_U; ++u) {
                const int64_t k = i + (int64_t)u * VEC_THREADS;
                if (k < i_end) {
                    float4 hv, bv;
                    TW_ONE(xv[u].x, sv[u].x, av[u].x, hv.x, bv.x) TW_ONE(xv[u].y, sv[u].y, av[u].y, hv.y, bv.y)
                    TW_ONE(xv[u].z, sv[u].z, av[u].z, hv.z, bv.z) TW_ONE(xv[u].w, sv[u].w, av[u].w, hv.w, bv.w)
                    h4[k] = hv;
                    if (b4) b4[k] = bv;
                }
            }
            i += (int64_t)VEC_U * VEC_THREADS;
            if (i < i_end) load_batch(i);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += VEC_THREADS) {
            float hv, bv;
            const float av = bvec ? atb[base + i] : 0.f;
            TW_ONE(x[base + i], s[base + i], av, hv, bv)
            xhat0[base + i] = hv;
            if (bvec) bvec[base + i] = bv;
        }
    }
#undef TW_ONE
}

template <bool VEC4>
__global__ void __launch_bounds__(VEC_THREADS)
ddim_kernel(const float *__restrict__ xhat, const float *__restrict__ s,
            const float *__restrict__ eps, const float *__restrict__ t,
            const float *__restrict__ tp, const float *__restrict__ abar, int n_table,
            float eta, float eta2, float *__restrict__ out, int64_t numel)
{
    scd_pdl_wait();                               // predecessor complete, its writes visible
    scd_pdl_trigger();
    const int b = blockIdx.y;
    int64_t lo, hi;
    slice_of(numel, lo, hi);
    const size_t base = (size_t)b * numel;
    const float4 *x4 = reinterpret_cast<const float4 *>(xhat + base);
    const float4 *s4 = reinterpret_cast<const float4 *>(s + base);
    const float4 *e4 = reinterpret_cast<const float4 *>(eps + base);
    float4 xv[VEC_U], sv[VEC_U], ev[VEC_U];
    const int64_t i_end = hi >> 2;
    const int64_t i0 = (lo >> 2) + threadIdx.x;
    auto load_batch = [&](int64_t i) {
#pragma unroll
        for (int u = 0; u < VEC_U; ++u) {
            const int64_t k = i + (int64_t)u * VEC_THREADS;
            if (k < i_end) { xv[u] = x4[k]; sv[u] = s4[k]; ev[u] = e4[k]; }
        }
    };
    if (VEC4) load_batch(i0);                     // in flight while the coefficients are formed
    // mean_t, mean_tminus1 and tbeta in the reference's operation order
    const float m_t = __fsqrt_rn(abar_at(abar, n_table, t[b]));
    const float m_p = __fsqrt_rn(abar_at(abar, n_table, tp[b]));
    const float mp2 = __fmul_rn(m_p, m_p), mt2 = __fmul_rn(m_t, m_t);
    const float q1 = __fsqrt_rn(__fdiv_rn(__fsub_rn(1.0f, mp2), __fsub_rn(1.0f, mt2)));
    const float q2 = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(mt2, __fdiv_rn(1.0f, mp2))));
    float tbeta = __fmul_rn(q1, q2);
    if (tbeta != tbeta) tbeta = 0.f;                         // isnan -> 0
    const float cdet = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, mp2),
                                            __fmul_rn(__fmul_rn(tbeta, tbeta), eta2)));
    const float csto = __fmul_rn(eta, tbeta);
#define DD_ONE(XH, S_, E)                                                          \
    __fadd_rn(__fadd_rn(__fmul_rn((XH), m_p), __fmul_rn(cdet, (S_))), __fmul_rn(csto, (E)))
    if (VEC4) {
        float4 *o4 = reinterpret_cast<float4 *>(out + base);
        for (int64_t i = i0; i < i_end; ) {
#pragma unroll
            for (int u = 0; u < VEC_U; ++u) {
                const int64_t k = i + (int64_t)u * VEC_THREADS;
                if (k < i_end) {
                    float4 ov;
                    ov.x = DD_ONE(xv[u].x, sv[u].x, ev[u].x); ov.y = DD_ONE(xv[u].y, sv[u].y, ev[u].y);
                    ov.z = DD_ONE(xv[u].z, sv[u].z, ev[u].z); ov.w = DD_ONE(xv[u].w, sv[u].w, ev[u].w);
                    o4[k] = ov;
                }
            }
            i += (int64_t)VEC_U * VEC_THREADS;
            if (i < i_end) load_batch(i);
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += VEC_THREADS)
            out[base + i] = DD_ONE(xhat[base + i], s[base + i], eps[base + i]);
    }
#undef DD_ONE
}

// ------------------------------------------------------------- host side ---
static inline bool vec4_ok(int64_t numel, const void *a, const void *b = nullptr,
                           const void *c = nullptr, const void *d = nullptr,
                           const void *e = nullptr)
{
    if (numel & 3) return false;
    const void *ptrs[5] = {a, b, c, d, e};
    for (const void *q : ptrs)
        if (q && ((uintptr_t)q & 15)) return false;
    return true;
}

int scd_launch_cg_update_xr(const float *x_in, float *x, float *r, const float *p, const float *d,
                            const float *rr_part, int rr_n, const float *pd_part, int pd_n,
                            int part_stride, float *rr_new_part, int batch, int64_t numel,
                            cudaStream_t st)
{
    if (batch <= 0 || numel <= 0) return 0;
    dim3 grid(scd_vec_blocks_per_sample(numel, batch), batch);
    if (vec4_ok(numel, x, r, p, d, x_in))
        SCD_CUDA(scd_launch_kernel(cg_update_xr_kernel<true>, grid, dim3(VEC_THREADS), 0, st, 0, x_in, x, r, p, d, rr_part, rr_n, pd_part, pd_n,
                                                               part_stride, rr_new_part, numel));
    else
        SCD_CUDA(scd_launch_kernel(cg_update_xr_kernel<false>, grid, dim3(VEC_THREADS), 0, st, 0, x_in, x, r, p, d, rr_part, rr_n, pd_part, pd_n,
                                                                part_stride, rr_new_part, numel));
    SCD_LAUNCH_CHECK("cg_update_xr_kernel");
    return 0;
}

int scd_launch_tweedie_rhs(const float *x, const float *s, const float *atb, const float *t,
                           const float *abar, int n_table, float gamma, float *xhat0, float *b,
                           int batch, int64_t numel, cudaStream_t st)
{
    if (batch <= 0 || numel <= 0) return 0;
    dim3 grid(scd_vec_blocks_per_sample(numel, batch), batch);
    if (vec4_ok(numel, x, s, atb, xhat0, b))
        SCD_CUDA(scd_launch_kernel(tweedie_rhs_kernel<true>, grid, dim3(VEC_THREADS), 0, st, 0, x, s, atb, t, abar, n_table, gamma, xhat0, b, numel));
    else
        SCD_CUDA(scd_launch_kernel(tweedie_rhs_kernel<false>, grid, dim3(VEC_THREADS), 0, st, 0, x, s, atb, t, abar, n_table, gamma, xhat0, b, numel));
    SCD_LAUNCH_CHECK("tweedie_rhs_kernel");
    return 0;
}

int scd_launch_ddim(const float *xhat, const float *s, const float *eps, const float *t,
                    const float *t_prev, const float *abar, int n_table, float eta, float eta2,
                    float *out, int batch, int64_t numel, cudaStream_t st)
{
    if (batch <= 0 || numel <= 0) return 0;
    dim3 grid(scd_vec_blocks_per_sample(numel, batch), batch);
    if (vec4_ok(numel, xhat, s, eps, out))
        SCD_CUDA(scd_launch_kernel(ddim_kernel<true>, grid, dim3(VEC_THREADS), 0, st, 0, xhat, s, eps, t, t_prev, abar, n_table, eta, eta2, out, numel));
    else
        SCD_CUDA(scd_launch_kernel(ddim_kernel<false>, grid, dim3(VEC_THREADS), 0, st, 0, xhat, s, eps, t, t_prev, abar, n_table, eta, eta2, out, numel));
    SCD_LAUNCH_CHECK("ddim_kernel");
    return 0;
}
