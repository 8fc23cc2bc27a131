// K8  vector kernels on sample-interleaved images ("il image": img_il[group][k0][k1][SB], SB >= 4 samples of
// a group interleaved per pixel, no pads).
//
// For batches of >= 3 samples the CG vectors x, r, p, d, b live in this layout for the whole solve: the
// projector fetches its strips from it with tensor copies (fp_march.cu), the backprojector's epilogue reads and
// writes it with one 16-byte access per lane (bp_tile.cu), and the recurrences below are element-wise.  Only the
// two ends of a data-consistency step touch the reference's [B][n0][n1] layout:
//
//   il_pack       user -> il (one or two arrays)                         entry of scd_cg / scd_fp
//   il_unpack     il -> user                                             exit of scd_cg
//   tweedie_il    xhat0 = (x - s*std_t)/mean_t (user), CG start iterate and b = xhat0 + gamma*atb (il)
//                 (reference src/samplers/utils.py:370-378 and :197) -- same arithmetic, bit for bit, as
//                 tweedie_rhs_kernel (vec_ops.cu)
//   ddim_il       DDPM branch of ddim() (reference src/samplers/utils.py:356-368) reading the CG result in il layout
//   cg_update_xr_il  alpha = rr/pd; x += alpha p; r -= alpha d; partial ||r||^2   (reference src/utils/cg.py:29-33)
//
// The layout changes go through a shared-memory tile of one image row x 128 pixels x SB samples: user-layout
// accesses are 512-byte runs per sample, il accesses are contiguous 16-byte vectors.
#include "scd_internal.cuh"

#define IL_THREADS 256
#define IL_TP      128          /* pixels of a tile (one image row) */

template <int SB> struct IlTile {
    static constexpr int STRIDE = IL_TP + 32 / SB;     // row stride = 2 / 4 / 8 mod 32: conflict-free vector phase
    static constexpr int Q = SB / 4;                   // float4 per pixel
};

__device__ __forceinline__ float il_abar_at(const float *abar, int n_table, float t)
{
    long long idx = (long long)t + 1;           // Tensor.long() truncates toward zero
    if (idx < 0) idx = 0;
    if (idx >= n_table) idx = n_table - 1;
    return abar[idx];
}

// tile[s][pix] -> il image (vector phase): thread = (pixel, quad of samples)
template <int SB>
__device__ __forceinline__ void il_store_tile(const float *tile, float *dst_il, int n1, size_t row_off, int K1, int tw)
{
    constexpr int ST = IlTile<SB>::STRIDE, Q = IlTile<SB>::Q;
    float4 *dst = reinterpret_cast<float4 *>(dst_il + (row_off + (size_t)K1) * SB);
    (void)n1;
    for (int i = threadIdx.x; i < tw * Q; i += IL_THREADS) {
        const int pix = i / Q, q = i - pix * Q;
        const float *t = tile + (q * 4) * ST + pix;
        dst[i] = make_float4(t[0], t[ST], t[2 * ST], t[3 * ST]);
    }
}

template <int SB>
__device__ __forceinline__ void il_load_tile(float *tile, const float *src_il, size_t row_off, int K1, int tw)
{
    constexpr int ST = IlTile<SB>::STRIDE, Q = IlTile<SB>::Q;
    const float4 *src = reinterpret_cast<const float4 *>(src_il + (row_off + (size_t)K1) * SB);
    for (int i = threadIdx.x; i < tw * Q; i += IL_THREADS) {
        const int pix = i / Q, q = i - pix * Q;
        const float4 v = src[i];
        float *t = tile + (q * 4) * ST + pix;
        t[0] = v.x; t[ST] = v.y; t[2 * ST] = v.z; t[3 * ST] = v.w;
    }
}

// grid = (pixel tiles of a row, image rows, groups)
template <int SB>
__global__ void __launch_bounds__(IL_THREADS)
il_pack_kernel(const float *__restrict__ a_user, float *__restrict__ a_il, const float *__restrict__ b_user,
               float *__restrict__ b_il, int n0, int n1, int batch)
{
    constexpr int ST = IlTile<SB>::STRIDE;
    __shared__ float ta[SB * ST], tb[SB * ST];
    scd_pdl_wait();
    scd_pdl_trigger();
    const int K1 = blockIdx.x * IL_TP, k0 = blockIdx.y, grp = blockIdx.z;
    const int tw = min(IL_TP, n1 - K1);
    const size_t isz = (size_t)n0 * n1;
    const int px = threadIdx.x & (IL_TP - 1), ps = threadIdx.x >> 7;       // 2 samples per round
    // all loads of the thread in flight before the first one is consumed (the kernel is pure data movement)
    float va[SB / 2], vb[SB / 2];
#pragma unroll
    for (int i = 0; i < SB / 2; ++i) {
        const int b = grp * SB + 2 * i + ps;
        va[i] = vb[i] = 0.f;
        if (b < batch && px < tw) {
            const size_t o = (size_t)b * isz + (size_t)k0 * n1 + K1 + px;
            va[i] = __ldg(a_user + o);
            if (b_user) vb[i] = __ldg(b_user + o);
        }
    }
#pragma unroll
    for (int i = 0; i < SB / 2; ++i) {
        ta[(2 * i + ps) * ST + px] = va[i];
        if (b_user) tb[(2 * i + ps) * ST + px] = vb[i];
    }
    __syncthreads();
    const size_t row_off = ((size_t)grp * n0 + k0) * n1;
    il_store_tile<SB>(ta, a_il, n1, row_off, K1, tw);
    if (b_user) il_store_tile<SB>(tb, b_il, n1, row_off, K1, tw);
}

template <int SB>
__global__ void __launch_bounds__(IL_THREADS)
il_unpack_kernel(const float *__restrict__ a_il, float *__restrict__ a_user, int n0, int n1, int batch)
{
    constexpr int ST = IlTile<SB>::STRIDE;
    __shared__ float ta[SB * ST];
    scd_pdl_wait();
    scd_pdl_trigger();
    const int K1 = blockIdx.x * IL_TP, k0 = blockIdx.y, grp = blockIdx.z;
    const int tw = min(IL_TP, n1 - K1);
    const size_t isz = (size_t)n0 * n1;
    il_load_tile<SB>(ta, a_il, ((size_t)grp * n0 + k0) * n1, K1, tw);
    __syncthreads();
    const int px = threadIdx.x & (IL_TP - 1), ps = threadIdx.x >> 7;
#pragma unroll 4
    for (int s0 = 0; s0 < SB; s0 += 2) {
        const int s = s0 + ps, b = grp * SB + s;
        if (b < batch && px < tw) a_user[(size_t)b * isz + (size_t)k0 * n1 + K1 + px] = ta[s * ST + px];
    }
}

template <int SB>
__global__ void __launch_bounds__(IL_THREADS)
tweedie_il_kernel(const float *__restrict__ x, const float *__restrict__ sc, const float *__restrict__ atb,
                  const float *__restrict__ t, const float *__restrict__ abar, int n_table, float gamma,
                  float *__restrict__ xhat0, float *__restrict__ x_il, float *__restrict__ b_il,
                  int n0, int n1, int batch)
{
    constexpr int ST = IlTile<SB>::STRIDE;
    __shared__ float ta[SB * ST], tb[SB * ST];
    __shared__ float coef[SB][2];
    scd_pdl_wait();
    scd_pdl_trigger();
    const int K1 = blockIdx.x * IL_TP, k0 = blockIdx.y, grp = blockIdx.z;
    const int tw = min(IL_TP, n1 - K1);
    const size_t isz = (size_t)n0 * n1;
    if (threadIdx.x < SB) {
        const int b = grp * SB + threadIdx.x;
        float stdv = 0.f, div = 0.f;
        if (b < batch) {
            const float ab = il_abar_at(abar, n_table, t[b]);
            const float mean = __fsqrt_rn(ab);                     // bar_a.pow(.5)
            stdv = __fsqrt_rn(__fsub_rn(1.0f, ab));                // (1 - bar_a).pow(.5)
            div = __fdiv_rn(1.0f, mean);                           // mean.pow(-1)
        }
        coef[threadIdx.x][0] = stdv; coef[threadIdx.x][1] = div;
    }
    const int px = threadIdx.x & (IL_TP - 1), ps = threadIdx.x >> 7;
    // loads of four rounds in flight before the coefficients are needed
    float xv[SB / 2], sv[SB / 2], av[SB / 2];
#pragma unroll
    for (int i = 0; i < SB / 2; ++i) {
        const int s = 2 * i + ps, b = grp * SB + s;
        xv[i] = sv[i] = av[i] = 0.f;
        if (b < batch && px < tw) {
            const size_t o = (size_t)b * isz + (size_t)k0 * n1 + K1 + px;
            xv[i] = __ldg(x + o); sv[i] = __ldg(sc + o); av[i] = __ldg(atb + o);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SB / 2; ++i) {
        const int s = 2 * i + ps, b = grp * SB + s;
        float hv = 0.f, bv = 0.f;
        if (b < batch && px < tw) {
            const float u = __fsub_rn(xv[i], __fmul_rn(sv[i], coef[s][0]));
            hv = __fmul_rn(u, coef[s][1]);
            bv = __fadd_rn(hv, __fmul_rn(gamma, av[i]));
            xhat0[(size_t)b * isz + (size_t)k0 * n1 + K1 + px] = hv;
        }
        ta[s * ST + px] = hv;
        tb[s * ST + px] = bv;
    }
    __syncthreads();
    const size_t row_off = ((size_t)grp * n0 + k0) * n1;
    il_store_tile<SB>(ta, x_il, n1, row_off, K1, tw);
    il_store_tile<SB>(tb, b_il, n1, row_off, K1, tw);
}

template <int SB>
__global__ void __launch_bounds__(IL_THREADS)
ddim_il_kernel(const float *__restrict__ xh_il, const float *__restrict__ sc, const float *__restrict__ eps,
               const float *__restrict__ t, const float *__restrict__ tp, const float *__restrict__ abar, int n_table,
               float eta, float eta2, float *__restrict__ out, int n0, int n1, int batch)
{
    constexpr int ST = IlTile<SB>::STRIDE;
    __shared__ float ta[SB * ST];
    __shared__ float coef[SB][3];
    scd_pdl_wait();
    scd_pdl_trigger();
    const int K1 = blockIdx.x * IL_TP, k0 = blockIdx.y, grp = blockIdx.z;
    const int tw = min(IL_TP, n1 - K1);
    const size_t isz = (size_t)n0 * n1;
    const int px = threadIdx.x & (IL_TP - 1), ps = threadIdx.x >> 7;
    float sv[SB / 2], ev[SB / 2];
#pragma unroll
    for (int i = 0; i < SB / 2; ++i) {
        const int s = 2 * i + ps, b = grp * SB + s;
        sv[i] = ev[i] = 0.f;
        if (b < batch && px < tw) {
            const size_t o = (size_t)b * isz + (size_t)k0 * n1 + K1 + px;
            sv[i] = __ldg(sc + o); ev[i] = __ldg(eps + o);
        }
    }
    il_load_tile<SB>(ta, xh_il, ((size_t)grp * n0 + k0) * n1, K1, tw);
    if (threadIdx.x < SB) {
        const int b = grp * SB + threadIdx.x;
        float m_p = 0.f, cdet = 0.f, csto = 0.f;
        if (b < batch) {
            // mean_t, mean_tminus1 and tbeta in the reference's operation order (as ddim_kernel, vec_ops.cu)
            const float m_t = __fsqrt_rn(il_abar_at(abar, n_table, t[b]));
            m_p = __fsqrt_rn(il_abar_at(abar, n_table, tp[b]));
            const float mp2 = __fmul_rn(m_p, m_p), mt2 = __fmul_rn(m_t, m_t);
            const float q1 = __fsqrt_rn(__fdiv_rn(__fsub_rn(1.0f, mp2), __fsub_rn(1.0f, mt2)));
            const float q2 = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(mt2, __fdiv_rn(1.0f, mp2))));
            float tbeta = __fmul_rn(q1, q2);
            if (tbeta != tbeta) tbeta = 0.f;                         // isnan -> 0
            cdet = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, mp2), __fmul_rn(__fmul_rn(tbeta, tbeta), eta2)));
            csto = __fmul_rn(eta, tbeta);
        }
        coef[threadIdx.x][0] = m_p; coef[threadIdx.x][1] = cdet; coef[threadIdx.x][2] = csto;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SB / 2; ++i) {
        const int s = 2 * i + ps, b = grp * SB + s;
        if (b < batch && px < tw) {
            const float xh = ta[s * ST + px];
            out[(size_t)b * isz + (size_t)k0 * n1 + K1 + px] =
                __fadd_rn(__fadd_rn(__fmul_rn(xh, coef[s][0]), __fmul_rn(coef[s][1], sv[i])), __fmul_rn(coef[s][2], ev[i]));
        }
    }
}

// grid = (blocks per group, groups); a thread's float4 always holds the same four samples of the group
#define ILV_U 4
template <int SB>
__global__ void __launch_bounds__(IL_THREADS)
cg_update_xr_il_kernel(const float *x_in, float *x, float *__restrict__ r, const float *__restrict__ p, const float *__restrict__ d,
                       const float *__restrict__ rr_part, int rr_n, const float *__restrict__ pd_part, int pd_n,
                       int part_stride, float *__restrict__ rr_new_part, size_t group_f4, int batch)
{
    constexpr int Q = SB / 4;
    __shared__ float alpha_s[SB];
    __shared__ float red[IL_THREADS / 32][SB];
    scd_pdl_wait();
    scd_pdl_trigger();
    const int grp = blockIdx.y;
    const size_t base = (size_t)grp * group_f4;
    const float4 *xi4 = reinterpret_cast<const float4 *>(x_in) + base;
    float4 *x4 = reinterpret_cast<float4 *>(x) + base;
    float4 *r4 = reinterpret_cast<float4 *>(r) + base;
    const float4 *p4 = reinterpret_cast<const float4 *>(p) + base;
    const float4 *d4 = reinterpret_cast<const float4 *>(d) + base;
    const size_t step = (size_t)gridDim.x * IL_THREADS;
    size_t i0 = (size_t)blockIdx.x * IL_THREADS + threadIdx.x;
    float4 xv[ILV_U], rv[ILV_U], pv[ILV_U], dv[ILV_U];
    auto load_batch = [&](size_t i) {
#pragma unroll
        for (int u = 0; u < ILV_U; ++u) {
            const size_t k = i + (size_t)u * step;
            if (k < group_f4) { xv[u] = xi4[k]; rv[u] = r4[k]; pv[u] = p4[k]; dv[u] = d4[k]; }
        }
    };
    load_batch(i0);                                // in flight while the per-sample scalars are formed
    {
        // alpha of the group's samples: warp w adds the partials of samples w, w + 8 (fixed order)
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int s = w; s < SB; s += IL_THREADS / 32) {
            const int b = grp * SB + s;
            float rr = 0.f, pd = 0.f;
            if (b < batch) {
                for (int i = lane; i < rr_n; i += 32) rr += rr_part[(size_t)b * part_stride + i];
                for (int i = lane; i < pd_n; i += 32) pd += pd_part[(size_t)b * part_stride + i];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                rr += __shfl_xor_sync(0xffffffffu, rr, off);
                pd += __shfl_xor_sync(0xffffffffu, pd, off);
            }
            if (lane == 0) alpha_s[s] = b < batch ? __fdiv_rn(rr, pd) : 0.f;     // no guard: same as the reference
        }
    }
    __syncthreads();
    const int q = threadIdx.x % Q;
    const float a0 = alpha_s[q * 4], a1 = alpha_s[q * 4 + 1], a2 = alpha_s[q * 4 + 2], a3 = alpha_s[q * 4 + 3];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (size_t i = i0; i < group_f4; ) {
#pragma unroll
        for (int u = 0; u < ILV_U; ++u) {
            const size_t k = i + (size_t)u * step;
            if (k < group_f4) {
                float4 xo = xv[u], ro = rv[u];
                xo.x = fmaf(a0, pv[u].x, xo.x); xo.y = fmaf(a1, pv[u].y, xo.y);
                xo.z = fmaf(a2, pv[u].z, xo.z); xo.w = fmaf(a3, pv[u].w, xo.w);
                ro.x = fmaf(-a0, dv[u].x, ro.x); ro.y = fmaf(-a1, dv[u].y, ro.y);
                ro.z = fmaf(-a2, dv[u].z, ro.z); ro.w = fmaf(-a3, dv[u].w, ro.w);
                x4[k] = xo; r4[k] = ro;
                acc[0] = fmaf(ro.x, ro.x, acc[0]); acc[1] = fmaf(ro.y, ro.y, acc[1]);
                acc[2] = fmaf(ro.z, ro.z, acc[2]); acc[3] = fmaf(ro.w, ro.w, acc[3]);
            }
        }
        i += (size_t)ILV_U * step;
        if (i < group_f4) load_batch(i);
    }
    // lanes with equal lane % Q hold the same samples
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float v = acc[j];
#pragma unroll
        for (int off = 16; off >= Q; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane < Q) red[w][lane * 4 + j] = v;
    }
    __syncthreads();
    if (threadIdx.x < SB) {
        const int b = grp * SB + threadIdx.x;
        if (b < batch) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < IL_THREADS / 32; ++k) s += red[k][threadIdx.x];
            rr_new_part[(size_t)b * part_stride + blockIdx.x] = s;
        }
    }
}

// ------------------------------------------------------------- host side ---
int scd_il_vec_blocks(const scd_geom *g, int batch)
{
    const int SB = scd_group_samples(g, batch);
    const int groups = (batch + SB - 1) / SB;
    const size_t f4 = (size_t)g->n0 * g->n1 * SB / 4;
    long nb = (4L * g->sm_count + groups - 1) / groups;
    const long cap = (long)((f4 + IL_THREADS - 1) / IL_THREADS);
    if (nb > cap) nb = cap;
    if (nb > g->sm_count) nb = g->sm_count;      // one block per SM for a single group; the consumers add the partials
    if (nb < 1) nb = 1;
    return (int)nb;
}

#define IL_DISPATCH(SBV, CALL)                                                     \
    switch (SBV) {                                                                 \
    case 4:  { constexpr int SB_ = 4;  CALL; } break;                              \
    case 8:  { constexpr int SB_ = 8;  CALL; } break;                              \
    case 16: { constexpr int SB_ = 16; CALL; } break;                              \
    default: scd_set_error("il image: unsupported group size %d", SBV); return SCD_E_INVALID; }

static int il_grid(const scd_geom *g, int batch, dim3 &grid, int &SB)
{
    SB = scd_group_samples(g, batch);
    const int groups = (batch + SB - 1) / SB;
    if (g->n0 > 65535 || groups > 65535) { scd_set_error("il image: image / batch too large"); return SCD_E_INVALID; }
    grid = dim3((g->n1 + IL_TP - 1) / IL_TP, g->n0, groups);
    return 0;
}

int scd_launch_il_pack(const scd_geom *g, const float *a_user, float *a_il, const float *b_user, float *b_il,
                       int batch, cudaStream_t st)
{
    if (batch <= 0) return 0;
    dim3 grid; int SB, rc;
    if ((rc = il_grid(g, batch, grid, SB))) return rc;
    if (SB == 1) {                                  // one sample per group: the two layouts coincide
        const size_t bytes = (size_t)batch * g->n0 * g->n1 * 4;
        if (a_il != a_user) SCD_CUDA(cudaMemcpyAsync(a_il, a_user, bytes, cudaMemcpyDeviceToDevice, st));
        if (b_user && b_il != b_user) SCD_CUDA(cudaMemcpyAsync(b_il, b_user, bytes, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    IL_DISPATCH(SB, SCD_CUDA(scd_launch_kernel(il_pack_kernel<SB_>, grid, dim3(IL_THREADS), 0, st, 0, a_user, a_il, b_user, b_il,
                                               g->n0, g->n1, batch)))
    SCD_LAUNCH_CHECK("il_pack_kernel");
    return 0;
}

int scd_launch_il_unpack(const scd_geom *g, const float *a_il, float *a_user, int batch, cudaStream_t st)
{
    if (batch <= 0) return 0;
    dim3 grid; int SB, rc;
    if ((rc = il_grid(g, batch, grid, SB))) return rc;
    if (SB == 1) {
        if (a_il != a_user) SCD_CUDA(cudaMemcpyAsync(a_user, a_il, (size_t)batch * g->n0 * g->n1 * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    IL_DISPATCH(SB, SCD_CUDA(scd_launch_kernel(il_unpack_kernel<SB_>, grid, dim3(IL_THREADS), 0, st, 0, a_il, a_user, g->n0, g->n1, batch)))
    SCD_LAUNCH_CHECK("il_unpack_kernel");
    return 0;
}

int scd_launch_tweedie_il(const scd_geom *g, const float *x, const float *s, const float *atb, const float *t,
                          const float *abar, int n_table, float gamma, float *xhat0_user, float *x_il, float *b_il,
                          int batch, cudaStream_t st)
{
    if (batch <= 0) return 0;
    dim3 grid; int SB, rc;
    if ((rc = il_grid(g, batch, grid, SB))) return rc;
    IL_DISPATCH(SB, SCD_CUDA(scd_launch_kernel(tweedie_il_kernel<SB_>, grid, dim3(IL_THREADS), 0, st, 0, x, s, atb, t, abar, n_table, gamma,
                                               xhat0_user, x_il, b_il, g->n0, g->n1, batch)))
    SCD_LAUNCH_CHECK("tweedie_il_kernel");
    return 0;
}

int scd_launch_ddim_il(const scd_geom *g, const float *xh_il, const float *s, const float *eps, const float *t,
                       const float *t_prev, const float *abar, int n_table, float eta, float eta2, float *out,
                       int batch, cudaStream_t st)
{
    if (batch <= 0) return 0;
    dim3 grid; int SB, rc;
    if ((rc = il_grid(g, batch, grid, SB))) return rc;
    IL_DISPATCH(SB, SCD_CUDA(scd_launch_kernel(ddim_il_kernel<SB_>, grid, dim3(IL_THREADS), 0, st, 0, xh_il, s, eps, t, t_prev, abar, n_table,
                                               eta, eta2, out, g->n0, g->n1, batch)))
    SCD_LAUNCH_CHECK("ddim_il_kernel");
    return 0;
}

int scd_launch_cg_update_xr_il(const scd_geom *g, const float *x_in, float *x, float *r, const float *p, const float *d,
                               const float *rr_part, int rr_n, const float *pd_part, int pd_n, int part_stride,
                               float *rr_new_part, int batch, cudaStream_t st)
{
    if (batch <= 0) return 0;
    const int SB = scd_group_samples(g, batch);
    const int groups = (batch + SB - 1) / SB;
    const size_t f4 = (size_t)g->n0 * g->n1 * SB / 4;
    dim3 grid(scd_il_vec_blocks(g, batch), groups);
    IL_DISPATCH(SB, SCD_CUDA(scd_launch_kernel(cg_update_xr_il_kernel<SB_>, grid, dim3(IL_THREADS), 0, st, 0, x_in, x, r, p, d, rr_part, rr_n,
                                               pd_part, pd_n, part_stride, rr_new_part, f4, batch)))
    SCD_LAUNCH_CHECK("cg_update_xr_il_kernel");
    return 0;
}
