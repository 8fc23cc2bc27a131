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

struct FpRun { int cls, first, count, cta0; };

struct FpParams {
    const float   *img;
    float         *sino;
    const FpAngle *fp;
    const int     *order;
    int n0, n1, n_angles, n_det, batch;
    int NA, TR, pitch;
    int n_runs;
    FpRun runs[8];
};

#define SCD_MAGIC      12582912.0f      /* 1.5 * 2^23: float add rounds to integer */
#define SCD_MAGIC_BITS 0x4B400000

template <int S, int RPT>
__global__ void __launch_bounds__(512)
fp_joseph_kernel(const FpParams P)
{
    extern __shared__ float tile[];           // [S][TR][pitch]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;

    // locate this CTA's run without indexing the parameter array dynamically
    FpRun R = P.runs[0];
#pragma unroll
    for (int k = 1; k < 8; ++k)
        if (k < P.n_runs && (int)blockIdx.x >= P.runs[k].cta0) R = P.runs[k];
    const int pos0 = R.first + ((int)blockIdx.x - R.cta0) * P.NA;
    const int na = min(P.NA, R.first + R.count - pos0);
    // cls < 0: unsorted fallback (one angle per CTA, positions are angle ids)
    const int cls = R.cls < 0 ? P.fp[pos0].cls : R.cls;
    const int nrows = cls == 0 ? P.n0 : P.n1;   // marching axis
    const int ncols = cls == 0 ? P.n1 : P.n0;   // interpolation axis
    const int b0 = blockIdx.y * S;
    const int TR = P.TR, pitch = P.pitch;
    const int plane = TR * pitch;

    // ---- per-ray setup (fp64 for the affine start position) ---------------
    const int nrays = na * P.n_det;
    float u0[RPT], bb[RPT], sc[RPT], acc[RPT][S];
    int   oidx[RPT];
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
        const int ray = tid + m * nthr;
        u0[m] = -1.0e30f; bb[m] = 0.f; sc[m] = 0.f; oidx[m] = -1;
        if (ray < nrays) {
            const int ai = ray / P.n_det;
            const int j = ray - ai * P.n_det;
            const int ang = R.cls < 0 ? pos0 + ai : P.order[pos0 + ai];
            const FpAngle f = P.fp[ang];
            // z' = u + 1 (left pad column) - 0.5 (round-to-nearest == floor)
            u0[m] = (float)(f.a * (double)j + f.c + 0.5);
            bb[m] = (float)f.b;
            sc[m] = f.scale;
            oidx[m] = ang * P.n_det + j;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) acc[m][s] = 0.f;
    }

    // ---- zero the pad columns once ----------------------------------------
    for (int row = tid; row < S * TR; row += nthr) {
        float *q = tile + row * pitch;
        q[0] = 0.f;
        for (int c = ncols + 1; c < pitch; ++c) q[c] = 0.f;
    }

    const float zlim = (float)ncols + 0.5f;
    for (int r0 = 0; r0 < nrows; r0 += TR) {
        __syncthreads();                       // previous strip fully consumed
        // ---- fill the strip ------------------------------------------------
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int b = b0 + s;
            const bool bok = b < P.batch;
            const float *src = P.img + (size_t)(bok ? b : 0) * P.n0 * P.n1;
            float *dst = tile + s * plane + 1;
            if (cls == 0) {
                // tile row rr <- image row r0+rr (contiguous): coalesced both sides
                for (int rr = warp; rr < TR; rr += nwarps) {
                    const int r = r0 + rr;
                    const bool ok = bok && r < nrows;
                    const float *g = src + (size_t)(ok ? r : 0) * P.n1;
                    float *d = dst + rr * pitch;
#pragma unroll 4
                    for (int c = lane; c < ncols; c += 32)
                        d[c] = ok ? __ldg(g + c) : 0.f;
                }
            } else {
                // transposed: tile row rr <- image column r0+rr.  Lanes run along
                // rr (global-contiguous k1); smem stride = pitch (odd) -> no conflicts.
                const int lpr = TR < 32 ? TR : 32;       // lanes along rr
                const int cpw = 32 / lpr;                // columns per warp pass
                const int lr = lane % lpr, lc = lane / lpr;
                for (int c = warp * cpw + lc; c < ncols; c += nwarps * cpw) {
                    const float *g = src + (size_t)c * P.n1 + r0;
#pragma unroll 2
                    for (int rr = lr; rr < TR; rr += lpr) {
                        const bool ok = bok && (r0 + rr) < nrows;
                        dst[rr * pitch + c] = ok ? __ldg(g + rr) : 0.f;
                    }
                }
            }
        }
        __syncthreads();

        // ---- march ---------------------------------------------------------
        const float r0f = (float)r0;
#pragma unroll
        for (int m = 0; m < RPT; ++m) {
            if (oidx[m] < 0) continue;
            const float bm = bb[m];
            const float z0 = fmaf(r0f, bm, u0[m]);
#pragma unroll 8
            for (int rr = 0; rr < TR; ++rr) {
                const float z = fmaf((float)rr, bm, z0);
                const float t = z + SCD_MAGIC;
                const float kf = t - SCD_MAGIC;
                const float w = (z - kf) + 0.5f;
                const int k = __float_as_int(t) - SCD_MAGIC_BITS;
                if (z >= -0.5f && z < zlim) {
                    const float *q = tile + rr * pitch + k;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const float f0 = q[s * plane];
                        const float f1 = q[s * plane + 1];
                        acc[m][s] += fmaf(w, f1 - f0, f0);
                    }
                }
            }
        }
    }

    // ---- write the line integrals -----------------------------------------
    const size_t sino_sz = (size_t)P.n_angles * P.n_det;
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
        if (oidx[m] < 0) continue;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int b = b0 + s;
            if (b < P.batch) P.sino[(size_t)b * sino_sz + oidx[m]] = acc[m][s] * sc[m];
        }
    }
}

// ------------------------------------------------------------- host side ---
struct FpConfig { int S, RPT, NA, TR, threads, pitch; size_t smem; };

static FpConfig fp_choose(const scd_geom *g, int batch, int n_sel_angles)
{
    FpConfig c;
    const int nmax = std::max(g->n0, g->n1);
    c.pitch = nmax + 2;
    if ((c.pitch & 1) == 0) c.pitch += 1;
    // samples per thread: amortise the index arithmetic when the batch is large
    c.S = batch >= 64 ? 4 : (batch >= 16 ? 2 : 1);
    if (g->tune_fp_samples) c.S = g->tune_fp_samples;
    c.NA = g->tune_fp_angles ? g->tune_fp_angles : 2;
    c.NA = std::max(1, std::min(c.NA, n_sel_angles));
    c.TR = g->tune_fp_rows ? g->tune_fp_rows : 32;
    // keep the tile inside the opt-in shared memory limit
    while ((size_t)c.S * c.TR * c.pitch * 4 > (size_t)g->smem_optin && c.TR > 8) c.TR >>= 1;
    while ((size_t)c.S * c.TR * c.pitch * 4 > (size_t)g->smem_optin && c.S > 1) c.S >>= 1;
    // threads own RPT <= 4 rays each; at most 512 threads per CTA
    int want = g->tune_fp_threads ? g->tune_fp_threads : 384;
    want = std::max(64, std::min(want, 512));
    while (c.NA > 1 && c.NA * g->n_det > 4 * 512) --c.NA;
    const int nrays = c.NA * g->n_det;
    c.RPT = 1;
    while (c.RPT < 4 && (nrays + c.RPT - 1) / c.RPT > want) c.RPT <<= 1;
    int thr = (nrays + c.RPT - 1) / c.RPT;
    thr = ((thr + 31) / 32) * 32;
    c.threads = std::max(thr, 64);
    c.smem = (size_t)c.S * c.TR * c.pitch * 4;
    return c;
}

static FpConfig fp_choose_na1(const scd_geom *g, FpConfig c)
{
    c.NA = 1;
    c.RPT = 1;
    while (c.RPT < 4 && (g->n_det + c.RPT - 1) / c.RPT > 512) c.RPT <<= 1;
    c.threads = std::max(64, ((g->n_det + c.RPT - 1) / c.RPT + 31) / 32 * 32);
    return c;
}

template <int S, int RPT>
static int fp_launch_t(const FpParams &P, dim3 grid, int threads, size_t smem, cudaStream_t st)
{
    static int configured_smem = 0;     // per instantiation
    if ((int)smem > configured_smem) {
        SCD_CUDA(cudaFuncSetAttribute(fp_joseph_kernel<S, RPT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_smem = (int)smem;
    }
    fp_joseph_kernel<S, RPT><<<grid, threads, smem, st>>>(P);
    SCD_LAUNCH_CHECK("fp_joseph_kernel");
    return 0;
}

int scd_launch_fp(const scd_geom *g, const float *img, float *sino, int batch,
                  int angle_lo, int angle_hi, cudaStream_t st)
{
    if (!g || !img || !sino) { scd_set_error("scd_fp: null argument"); return SCD_E_INVALID; }
    if (batch < 0 || angle_lo < 0 || angle_hi > g->n_angles || angle_lo > angle_hi) {
        scd_set_error("scd_fp: bad batch/angle range (batch=%d, angles [%d,%d) of %d)",
                      batch, angle_lo, angle_hi, g->n_angles);
        return SCD_E_INVALID;
    }
    if (batch == 0 || angle_lo == angle_hi) return 0;
    if (g->n_det > 2048) { scd_set_error("scd_fp: n_det > 2048 unsupported"); return SCD_E_INVALID; }

    FpConfig c = fp_choose(g, batch, angle_hi - angle_lo);
    FpParams P;
    P.img = img; P.sino = sino; P.fp = g->d_fp; P.order = g->d_order;
    P.n0 = g->n0; P.n1 = g->n1; P.n_angles = g->n_angles; P.n_det = g->n_det; P.batch = batch;
    P.TR = c.TR; P.pitch = c.pitch;

    // runs of order[] positions whose angle lies in [angle_lo, angle_hi), per class
    P.NA = c.NA;
    P.n_runs = 0;
    int cta = 0;
    bool overflow = false;
    int pos = 0;
    while (pos < g->n_angles) {
        const int a = g->h_order[pos];
        if (a < angle_lo || a >= angle_hi) { ++pos; continue; }
        const int cls = g->h_fp[a].cls;
        int end = pos + 1;
        while (end < g->n_angles && g->h_order[end] >= angle_lo && g->h_order[end] < angle_hi &&
               g->h_fp[g->h_order[end]].cls == cls) ++end;
        if (P.n_runs == 8) { overflow = true; break; }
        FpRun &r = P.runs[P.n_runs++];
        r.cls = cls; r.first = pos; r.count = end - pos; r.cta0 = cta;
        cta += (r.count + c.NA - 1) / c.NA;
        pos = end;
    }
    if (overflow) {
        // angle list interleaves the two classes too often: one angle per CTA,
        // class looked up per angle (cls = -1), positions are angle ids
        c = fp_choose_na1(g, c);
        P.NA = 1; P.n_runs = 1;
        P.runs[0].cls = -1; P.runs[0].first = angle_lo; P.runs[0].count = angle_hi - angle_lo;
        P.runs[0].cta0 = 0;
        cta = angle_hi - angle_lo;
    }
    dim3 grid(cta, (batch + c.S - 1) / c.S);
    if (grid.y > 65535) { scd_set_error("scd_fp: batch too large"); return SCD_E_INVALID; }
#define FP_CASE(SS, RR) if (c.S == SS && c.RPT == RR) return fp_launch_t<SS, RR>(P, grid, c.threads, c.smem, st);
    FP_CASE(1, 1) FP_CASE(1, 2) FP_CASE(1, 4)
    FP_CASE(2, 1) FP_CASE(2, 2) FP_CASE(2, 4)
    FP_CASE(4, 1) FP_CASE(4, 2) FP_CASE(4, 4)
#undef FP_CASE
    scd_set_error("scd_fp: unsupported config S=%d RPT=%d", c.S, c.RPT);
    return SCD_E_INVALID;
}
