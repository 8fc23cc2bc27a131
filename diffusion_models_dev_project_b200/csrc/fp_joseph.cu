// K1  fp_joseph -- ray-driven Joseph forward projector A, batched over samples.
//
// Replaces SimpleTrafo.trafo (reference src/physics/trafo.py:58), which reaches
// ASTRA's par_fp through ODL one image at a time.  Arithmetic follows
// march along the dominant axis, linear interpolation
// across it, weight dx/max(|cos|,|sin|), zero outside the image.
//
// Work decomposition
//   CTA      = (chunk of NA same-class angles, group of S samples)
//   thread   = RPT rays (detector bins) of the chunk, S samples each; the
//              accumulators live in registers for the whole march, so there is
//              no cross-CTA reduction and every sinogram entry is written once.
//   smem     = strip of TR marching rows x full interpolation axis, S planes,
//              zero-padded columns left/right (so edge taps need no branches);
//              class-1 angles read the image transposed while filling, so the
//              march itself is class-agnostic.
//   lanes    = adjacent detector bins -> at a fixed row they read addresses
//              1..1.41 floats apart: shared-memory wavefronts <= 2 per load.
//
// Binding resource: shared-memory bandwidth (8 B per ray-step and sample) and
// issue slots, not HBM -- see DESIGN.md "fp_joseph".
#include "scd_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

struct FpRun { int cls, first, count, cta0; };

// Packed image ("tile-ready") layout written by fp_pack_kernel and consumed by the
// march kernel with plain byte copies:
//   packed[group][cls][row][col][s]      s < S samples interleaved per pixel
//   cls 0: row = k0, col = k1            cls 1: row = k1, col = k0 (transposed)
//   each row has `pitch` = ncols + 3 pixels: 1 zero pad left, 2 right; rows are padded
//   with zero rows up to a multiple of TR, so a strip of TR rows is one contiguous,
//   16-byte aligned range = exactly the shared-memory tile.
struct FpLayout {
    int S, TR;
    int pitch[2];            // pixels per packed row, per class
    int rows[2];             // packed rows per class (multiple of TR)
    size_t cls_off[2];       // float offset of class c inside a group
    size_t group_floats;     // floats per sample group
};

struct FpParams {
    const float   *img;
    float         *sino;
    float         *packed;
    const FpAngle *fp;
    const int     *order;
    int n0, n1, n_angles, n_det, batch;
    int NA;                 // angles per CTA
    int G;                  // threads per angle group (multiple of 32)
    int nbuf;               // strips in flight (shared-memory ring depth)
    FpLayout L;
    int n_runs;
    FpRun runs[8];
};

#define SCD_MAGIC      12582912.0f      /* 1.5 * 2^23: float add rounds to integer */
#define SCD_MAGIC_BITS 0x4B400000

struct __align__(16) FpAngSmem { double a, b, c; float scale; int id; };

// ------------------------------------------------------------------ pack ---
// One block = one 32x32 image tile of one sample group; it writes the tile into BOTH packed
// orientations (class 0 directly, class 1 through a shared-memory transpose), including the
// zero pad pixels and the zero rows that round the row count up to a multiple of TR.
// Optional prologue (the producer of the image is fused into the pack pass):
//   mode 1  CG direction update   p = r + beta p        (reference src/utils/cg.py:35-38)
//   mode 2  Tweedie + CG rhs      xhat0, b              (reference src/samplers/utils.py:370-378, :197)
// grid = (tiles over axis 1, tiles over axis 0, groups), both tile ranges cover the padded rows.
template <int S, int MODE>
__global__ void __launch_bounds__(256)
fp_pack_kernel(const FpParams P, const FpPrologue Q)
{
    __shared__ float tr[S][32][33];
    __shared__ float coef[S][3];
    const int grp = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int K0 = blockIdx.y * 32, K1 = blockIdx.x * 32;
    const size_t isz = (size_t)P.n0 * P.n1;
    float *dst0 = P.packed + (size_t)grp * P.L.group_floats + P.L.cls_off[0];
    float *dst1 = P.packed + (size_t)grp * P.L.group_floats + P.L.cls_off[1];
    const int pitch0 = P.L.pitch[0], pitch1 = P.L.pitch[1];

    if (MODE != 0) {
        // per-sample scalars, computed by one warp per sample
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (w < S) {
            const int b = grp * S + w;
            if (b < P.batch) {
                if (MODE == 1) {
                    float rn = 0.f, ro = 0.f;
                    for (int i = lane; i < Q.rr_new_n; i += 32) rn += Q.rr_new_part[(size_t)b * Q.part_stride + i];
                    for (int i = lane; i < Q.rr_old_n; i += 32) ro += Q.rr_old_part[(size_t)b * Q.part_stride + i];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        rn += __shfl_xor_sync(0xffffffffu, rn, off);
                        ro += __shfl_xor_sync(0xffffffffu, ro, off);
                    }
                    if (lane == 0) coef[w][0] = __fdiv_rn(rn, ro);           // beta
                } else if (lane == 0) {
                    long long idx = (long long)Q.t[b] + 1;                   // Tensor.long() + 1
                    idx = idx < 0 ? 0 : (idx >= Q.n_table ? Q.n_table - 1 : idx);
                    const float ab = Q.abar[idx];
                    const float mean = __fsqrt_rn(ab);
                    coef[w][0] = __fsqrt_rn(__fsub_rn(1.0f, ab));            // std_t
                    coef[w][1] = __fdiv_rn(1.0f, mean);                      // mean_t^-1
                }
            }
        }
        __syncthreads();
    }

    // ---- produce the tile (prologue), plain outputs, class-0 packed rows ----
    for (int i = ty; i < 32; i += 8) {
        const int k0 = K0 + i, k1 = K1 + tx;
        const bool in_img = k0 < P.n0 && k1 < P.n1;
        float v[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int b = grp * S + s;
            v[s] = 0.f;
            if (in_img && b < P.batch) {
                const size_t o = b * isz + (size_t)k0 * P.n1 + k1;
                if (MODE == 0) {
                    v[s] = __ldg(P.img + o);
                } else if (MODE == 1) {
                    v[s] = fmaf(coef[s][0], Q.p[o], Q.r[o]);
                    Q.p[o] = v[s];
                } else {
                    const float u = __fsub_rn(Q.x[o], __fmul_rn(Q.s[o], coef[s][0]));
                    v[s] = __fmul_rn(u, coef[s][1]);
                    Q.xhat0[o] = v[s];
                    Q.b[o] = __fadd_rn(v[s], __fmul_rn(Q.gamma, Q.atb[o]));
                }
            }
            tr[s][i][tx] = v[s];
        }
        if (k0 < P.L.rows[0] && k1 < P.n1) {
            float *q = dst0 + ((size_t)k0 * pitch0 + 1 + k1) * S;
            if (S == 4) *reinterpret_cast<float4 *>(q) = make_float4(v[0], v[1], v[2], v[3]);
            else if (S == 2) *reinterpret_cast<float2 *>(q) = make_float2(v[0], v[1]);
            else q[0] = v[0];
            if (k1 == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s) q[s - S] = 0.f;
            }
            if (k1 == P.n1 - 1) {
#pragma unroll
                for (int s = 0; s < 2 * S; ++s) q[S + s] = 0.f;
            }
        }
    }
    __syncthreads();
    // ---- class-1 packed rows: row = k1, col = k0 (coalesced along k0) -------
    for (int i = ty; i < 32; i += 8) {
        const int k1 = K1 + i, k0 = K0 + tx;
        if (k1 < P.L.rows[1] && k0 < P.n0) {
            float *q = dst1 + ((size_t)k1 * pitch1 + 1 + k0) * S;
            if (S == 4) *reinterpret_cast<float4 *>(q) = make_float4(tr[0][tx][i], tr[1 % S][tx][i], tr[2 % S][tx][i], tr[3 % S][tx][i]);
            else if (S == 2) *reinterpret_cast<float2 *>(q) = make_float2(tr[0][tx][i], tr[1 % S][tx][i]);
            else q[0] = tr[0][tx][i];
            if (k0 == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s) q[s - S] = 0.f;
            }
            if (k0 == P.n0 - 1) {
#pragma unroll
                for (int s = 0; s < 2 * S; ++s) q[S + s] = 0.f;
            }
        }
    }
}

// ----------------------------------------------------------------- march ---
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int S> struct VecS;
template <> struct VecS<1> { typedef float  T; };
template <> struct VecS<2> { typedef float2 T; };
template <> struct VecS<4> { typedef float4 T; };

__device__ __forceinline__ void tap_acc(float (&p)[1], float l, float r, float wl, float w)
{ p[0] = fmaf(r, w, fmaf(l, wl, p[0])); }
__device__ __forceinline__ void tap_acc(float (&p)[2], float2 l, float2 r, float wl, float w)
{ p[0] = fmaf(r.x, w, fmaf(l.x, wl, p[0])); p[1] = fmaf(r.y, w, fmaf(l.y, wl, p[1])); }
__device__ __forceinline__ void tap_acc(float (&p)[4], float4 l, float4 r, float wl, float w)
{
    p[0] = fmaf(r.x, w, fmaf(l.x, wl, p[0])); p[1] = fmaf(r.y, w, fmaf(l.y, wl, p[1]));
    p[2] = fmaf(r.z, w, fmaf(l.z, wl, p[2])); p[3] = fmaf(r.w, w, fmaf(l.w, wl, p[3]));
}

// S  = samples marched together by one thread (interleaved per pixel: one LDS.(32*S) per tap)
// TR = marching rows per shared-memory strip
#define FP_MAX_NBUF 8
template <int S, int TR>
__global__ void __launch_bounds__(512)
fp_joseph_kernel(const FpParams P)
{
    typedef typename VecS<S>::T V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthr = blockDim.x;

    // locate this CTA's run without indexing the parameter array dynamically
    FpRun R = P.runs[0];
#pragma unroll
    for (int k = 1; k < 8; ++k)
        if (k < P.n_runs && (int)blockIdx.x >= P.runs[k].cta0) R = P.runs[k];
    const int pos0 = R.first + ((int)blockIdx.x - R.cta0) * P.NA;
    const int na = min(P.NA, R.first + R.count - pos0);
    const int cls = R.cls;
    const int nrows = cls == 0 ? P.n0 : P.n1;   // marching axis
    const int ncols = cls == 0 ? P.n1 : P.n0;   // interpolation axis
    const int pitch = P.L.pitch[cls];
    const int n_det = P.n_det;
    const int b0 = blockIdx.y * S;
    const unsigned strip_bytes = (unsigned)(TR * pitch * S * 4);

    // shared memory: tile[nbuf] | mbarriers | angle table | u0[NA*n_det] | acc[NA*n_det][S]
    const int NBUF = P.nbuf;
    unsigned char *tile0 = smem_raw;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw + NBUF * (size_t)strip_bytes);
    FpAngSmem *ang = reinterpret_cast<FpAngSmem *>(bars + FP_MAX_NBUF);
    float *u0s = reinterpret_cast<float *>(ang + P.NA);
    float *acc = u0s + ((P.NA * n_det + 3) & ~3);

    const float *src = P.packed + (size_t)blockIdx.y * P.L.group_floats + P.L.cls_off[cls];
    const int nstrips = (nrows + TR - 1) / TR;

    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        for (int i = 0; i < NBUF; ++i)
            if (i < nstrips) {
                mbar_expect_tx(&bars[i], strip_bytes);
                bulk_g2s(tile0 + (size_t)i * strip_bytes, src + (size_t)i * TR * pitch * S, strip_bytes, &bars[i]);
            }
    }
    for (int i = tid; i < na; i += nthr) {
        const int id = P.order[pos0 + i];
        const FpAngle f = P.fp[id];
        FpAngSmem a;
        a.a = f.a; a.b = f.b;
        a.c = f.c + 1.0;                 // zf = u + 1 (left pad pixel); floor(zf) = pixel of the left tap
        a.scale = f.scale; a.id = id;
        ang[i] = a;
    }
    __syncthreads();
    for (int e = tid; e < na * n_det; e += nthr) {
        const int ai = e / n_det, j = e - ai * n_det;
        u0s[e] = (float)(ang[ai].a * (double)j + ang[ai].c);
#pragma unroll
        for (int s = 0; s < S; ++s) acc[(size_t)e * S + s] = 0.f;
    }
    __syncthreads();                              // mbarrier init + tables visible

    const int G = P.G;
    const int grp = tid / G, gl = tid - grp * G, ngrp = nthr / G;
    const float zmax = (float)(ncols + 1);        // u = ncols: both taps on the right pads
    const float zmin = 0.0f;                      // u = -1:    left tap on the left pad, weight of the right tap 0

    for (int st = 0; st < nstrips; ++st) {
        const int bi = st % NBUF;
        mbar_wait(&bars[bi], (unsigned)((st / NBUF) & 1));
        const unsigned char *buf = tile0 + (size_t)bi * strip_bytes;
        const float r0f = (float)(st * TR);
        for (int ai = grp; ai < na; ai += ngrp) {
            const float bf = (float)ang[ai].b;
            const float bspan = bf * (float)(TR - 1);
            for (int j0 = 0; j0 < n_det; j0 += G) {
                const int j = j0 + gl;
                const bool live = j < n_det;
                const int e = ai * n_det + (live ? j : n_det - 1);
                const float z0 = fmaf(r0f, bf, u0s[e]);
                // whole warp outside the image for every row of this strip -> nothing to add
                const float za = z0, zb = z0 + bspan;
                const bool outside = !live || fmaxf(za, zb) <= zmin || fminf(za, zb) >= zmax;
                if (__all_sync(0xffffffffu, outside)) continue;
                float part[S];
#pragma unroll
                for (int s = 0; s < S; ++s) part[s] = 0.f;
#pragma unroll
                for (int rr = 0; rr < TR; ++rr) {
                    // clamped position: outside the image both taps land on zero pad pixels
                    const float z = fminf(fmaxf(fmaf((float)rr, bf, z0), zmin), zmax);
                    const float t = __fadd_rd(z, SCD_MAGIC);           // floor(z) in the mantissa
                    const float w = z - (t - SCD_MAGIC);
                    const float wl = 1.0f - w;
                    const V *q = reinterpret_cast<const V *>(buf + (size_t)rr * pitch * S * 4) +
                                 (__float_as_int(t) - SCD_MAGIC_BITS);
                    tap_acc(part, q[0], q[1], wl, w);
                }
                if (live) {
                    float *ap = acc + (size_t)e * S;
#pragma unroll
                    for (int s = 0; s < S; ++s) ap[s] += part[s];
                }
            }
        }
        __syncthreads();                          // every thread is done with buffer bi
        if (tid == 0 && st + NBUF < nstrips) {
            mbar_expect_tx(&bars[bi], strip_bytes);
            bulk_g2s(tile0 + (size_t)bi * strip_bytes, src + (size_t)(st + NBUF) * TR * pitch * S, strip_bytes, &bars[bi]);
        }
    }

    // ---- write the line integrals (coalesced along the detector) ----------
    const size_t sino_sz = (size_t)P.n_angles * n_det;
    for (int e = tid; e < na * n_det; e += nthr) {
        const int ai = e / n_det, j = e - ai * n_det;
        const float sc = ang[ai].scale;
        const size_t o = (size_t)ang[ai].id * n_det + j;
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (b0 + s < P.batch) P.sino[(size_t)(b0 + s) * sino_sz + o] = acc[(size_t)e * S + s] * sc;
    }
}

// ------------------------------------------------------------- host side ---
struct FpConfig { int S, TR, NA, G, threads, nbuf; FpLayout L; size_t smem; int groups; size_t scratch_bytes; };

static FpLayout fp_layout(const scd_geom *g, int S, int TR)
{
    FpLayout L;
    L.S = S; L.TR = TR;
    const int nr[2] = {g->n0, g->n1}, nc[2] = {g->n1, g->n0};
    size_t off = 0;
    for (int c = 0; c < 2; ++c) {
        L.pitch[c] = nc[c] + 3;
        L.rows[c] = ((nr[c] + TR - 1) / TR) * TR;
        L.cls_off[c] = off;
        off += (size_t)L.rows[c] * L.pitch[c] * S;
        off = (off + 31) & ~(size_t)31;            // keep every class 128-byte aligned
    }
    L.group_floats = off;
    return L;
}

static size_t fp_smem_bytes(const scd_geom *g, const FpConfig &c)
{
    const size_t strip = (size_t)c.TR * std::max(c.L.pitch[0], c.L.pitch[1]) * c.S * 4;
    const size_t nray = ((size_t)c.NA * g->n_det + 3) & ~(size_t)3;
    return c.nbuf * strip + FP_MAX_NBUF * 8 + sizeof(FpAngSmem) * c.NA + 4 * nray + 4 * nray * c.S + 16;
}

static FpConfig fp_choose(const scd_geom *g, int batch, int n_sel_angles)
{
    FpConfig c;
    // Large batches: 4 samples per thread (one LDS.128 per tap), 5 angles per CTA and 8-row
    // strips keep two 512-thread CTAs per SM (~103 KB each at 256x256).  Small batches trade
    // that efficiency for CTAs: one angle and fewer samples per CTA.
    c.S = batch >= 32 ? 4 : (batch >= 4 ? 2 : 1);
    if (g->tune_fp_samples) c.S = g->tune_fp_samples;
    if (c.S != 1 && c.S != 2 && c.S != 4) c.S = 1;
    c.groups = (batch + c.S - 1) / c.S;
    const int n_cta_angles = n_sel_angles;
    // One angle per CTA measured fastest at every batch size on B200 (more, smaller CTAs; the
    // packed strips are re-read from L2, which sustains it); larger chunks remain selectable.
    int na = 1;
    (void)n_cta_angles;
    c.NA = g->tune_fp_angles ? g->tune_fp_angles : na;
    c.NA = std::max(1, std::min(c.NA, n_sel_angles));
    c.TR = g->tune_fp_rows ? g->tune_fp_rows : (c.S == 4 ? 8 : 16);
    if (c.TR != 8 && c.TR != 16 && c.TR != 32) c.TR = 16;
    c.G = 128;
    int thr = g->tune_fp_threads ? g->tune_fp_threads : (c.NA >= 4 ? 512 : 384);
    if (c.NA == 1) c.G = ((std::min(thr, g->n_det) + 31) / 32) * 32;   // one group spans the detector
    thr = std::max(c.G, (thr / c.G) * c.G);
    c.threads = std::min(thr, 512);
    if (c.threads < c.G) c.G = c.threads;
    // ring depth: with few rays per CTA a strip is consumed faster than a bulk copy lands, so
    // keep several strips in flight; budget ~100 KB (2 CTAs/SM), whole opt-in smem if the grid
    // has at most one CTA per SM anyway
    c.nbuf = 2;
    for (;;) {
        c.L = fp_layout(g, c.S, c.TR);
        c.smem = fp_smem_bytes(g, c);
        if (c.smem <= (size_t)g->smem_optin) break;
        if (c.NA > 1 && (size_t)c.NA * g->n_det * (c.S + 1) * 4 > c.smem / 2) c.NA = (c.NA + 1) / 2;
        else if (c.TR > 8) c.TR >>= 1;
        else if (c.S > 1) c.S >>= 1;
        else if (c.NA > 1) c.NA = (c.NA + 1) / 2;
        else break;
    }
    {
        const int nr = std::max(g->n0, g->n1);
        const int nstrips = (nr + c.TR - 1) / c.TR;
        const long ctas = (long)((batch + c.S - 1) / c.S) * ((n_sel_angles + c.NA - 1) / c.NA);
        const size_t budget = ctas <= g->sm_count ? (size_t)g->smem_optin
                              : (ctas <= 2L * g->sm_count ? (size_t)110 * 1024 : (size_t)72 * 1024);
        int want = g->tune_fp_nbuf ? g->tune_fp_nbuf : 3;
        want = std::max(2, std::min(std::min(want, FP_MAX_NBUF), nstrips));
        while (c.nbuf < want) {
            c.nbuf++;
            if (fp_smem_bytes(g, c) > (g->tune_fp_nbuf ? (size_t)g->smem_optin : budget)) { c.nbuf--; break; }
        }
        c.smem = fp_smem_bytes(g, c);
    }
    c.groups = (batch + c.S - 1) / c.S;
    c.scratch_bytes = (size_t)c.groups * c.L.group_floats * 4;
    return c;
}

size_t scd_fp_scratch_need_v3(const scd_geom *g, int batch)
{
    if (!g || batch <= 0) return 0;
    // the configuration may be overridden by tuning: size for the worst case (S = 1 packs least densely)
    size_t need = 0;
    for (int S = 1; S <= 4; S <<= 1)
        for (int TR = 8; TR <= 32; TR <<= 1) {
            const FpLayout L = fp_layout(g, S, TR);
            need = std::max(need, (size_t)((batch + S - 1) / S) * L.group_floats * 4);
        }
    return need + 256;
}

template <int S, int TR>
static int fp_launch_t(const FpParams &P, dim3 grid, int threads, size_t smem, cudaStream_t st)
{
    static int configured_smem = 0;     // per instantiation
    if ((int)smem > configured_smem) {
        SCD_CUDA(cudaFuncSetAttribute(fp_joseph_kernel<S, TR>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_smem = (int)smem;
    }
    fp_joseph_kernel<S, TR><<<grid, threads, smem, st>>>(P);
    SCD_LAUNCH_CHECK("fp_joseph_kernel");
    return 0;
}

template <int S>
static int fp_pack_launch(const FpParams &P, const FpPrologue &Q, dim3 pg, cudaStream_t st)
{
    if (Q.mode == 0) fp_pack_kernel<S, 0><<<pg, 256, 0, st>>>(P, Q);
    else if (Q.mode == 1) fp_pack_kernel<S, 1><<<pg, 256, 0, st>>>(P, Q);
    else fp_pack_kernel<S, 2><<<pg, 256, 0, st>>>(P, Q);
    SCD_LAUNCH_CHECK("fp_pack_kernel");
    return 0;
}

int scd_launch_fp_v3(const scd_geom *g, const float *img, float *sino, int batch,
                  int angle_lo, int angle_hi, void *scratch, size_t scratch_bytes, cudaStream_t st,
                     const FpPrologue *prologue)
{
    FpPrologue Q;
    if (prologue) Q = *prologue; else { memset(&Q, 0, sizeof(Q)); }
    if (!g || (!img && Q.mode == 0) || !sino) { scd_set_error("scd_fp: null argument"); return SCD_E_INVALID; }
    if (batch < 0 || angle_lo < 0 || angle_hi > g->n_angles || angle_lo > angle_hi) {
        scd_set_error("scd_fp: bad batch/angle range (batch=%d, angles [%d,%d) of %d)",
                      batch, angle_lo, angle_hi, g->n_angles);
        return SCD_E_INVALID;
    }
    if (batch == 0 || angle_lo == angle_hi) return 0;

    FpConfig c = fp_choose(g, batch, angle_hi - angle_lo);
    if (c.smem > (size_t)g->smem_optin) { scd_set_error("scd_fp: image too large for the shared-memory strip"); return SCD_E_INVALID; }
    const uintptr_t sp = ((uintptr_t)scratch + 127) & ~(uintptr_t)127;
    if (!scratch || sp + c.scratch_bytes > (uintptr_t)scratch + scratch_bytes) {
        scd_set_error("scd_fp: scratch too small (%zu bytes given, %zu needed; see scd_fp_scratch_bytes)",
                      scratch_bytes, c.scratch_bytes + 128);
        return SCD_E_WORKSPACE;
    }
    FpParams P;
    P.img = img; P.sino = sino; P.packed = (float *)sp; P.fp = g->d_fp; P.order = g->d_order;
    P.n0 = g->n0; P.n1 = g->n1; P.n_angles = g->n_angles; P.n_det = g->n_det; P.batch = batch;
    P.NA = c.NA; P.G = c.G; P.nbuf = c.nbuf; P.L = c.L; P.n_runs = 0;
    const unsigned gy = (unsigned)c.groups;
    if (gy > 32767u) { scd_set_error("scd_fp: batch too large"); return SCD_E_INVALID; }

    // ---- pack: image -> tile-ready layout (both orientations), producer fused in ----
    {
        const int t0 = (std::max(g->n0, c.L.rows[0]) + 31) / 32, t1 = (std::max(g->n1, c.L.rows[1]) + 31) / 32;
        dim3 pg(t1, t0, gy);
        int rc;
        if (c.S == 1) rc = fp_pack_launch<1>(P, Q, pg, st);
        else if (c.S == 2) rc = fp_pack_launch<2>(P, Q, pg, st);
        else rc = fp_pack_launch<4>(P, Q, pg, st);
        if (rc) return rc;
    }

    // runs of order[] positions whose angle lies in [angle_lo, angle_hi), per class; a
    // launch takes up to 8 runs (monotone angle lists need at most 3), more runs -> more launches
    int pos = 0;
    while (pos < g->n_angles) {
        P.n_runs = 0;
        int cta = 0;
        while (pos < g->n_angles && P.n_runs < 8) {
            const int a = g->h_order[pos];
            if (a < angle_lo || a >= angle_hi) { ++pos; continue; }
            const int cls = g->h_fp[a].cls;
            int end = pos + 1;
            while (end < g->n_angles && g->h_order[end] >= angle_lo && g->h_order[end] < angle_hi &&
                   g->h_fp[g->h_order[end]].cls == cls) ++end;
            FpRun &r = P.runs[P.n_runs++];
            r.cls = cls; r.first = pos; r.count = end - pos; r.cta0 = cta;
            cta += (r.count + c.NA - 1) / c.NA;
            pos = end;
        }
        if (P.n_runs == 0) break;
        dim3 grid(cta, gy);
        int rc = SCD_E_INVALID;
#define FP_CASE(SS, TT) if (c.S == SS && c.TR == TT) rc = fp_launch_t<SS, TT>(P, grid, c.threads, c.smem, st);
        FP_CASE(1, 8) FP_CASE(1, 16) FP_CASE(1, 32)
        FP_CASE(2, 8) FP_CASE(2, 16) FP_CASE(2, 32)
        FP_CASE(4, 8) FP_CASE(4, 16) FP_CASE(4, 32)
#undef FP_CASE
        if (rc) return rc;
    }
    return 0;
}
