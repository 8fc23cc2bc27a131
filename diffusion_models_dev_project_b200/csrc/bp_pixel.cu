// K2  bp_pixel -- pixel-driven linear-interpolation backprojector A*, batched,
// with the CG axpy and dot product fused into its epilogue.
//
// Replaces SimpleTrafo.trafo_adjoint (reference src/physics/trafo.py:61 ->
// ODL -> ASTRA par_bp), the `x + gamma*A*(A x)` axpy of `op`
// (src/samplers/utils.py:188-189) and the <p,d> / ||r||^2 reductions of cg
// (src/utils/cg.py:22,27).  Arithmetic: SURVEY.md Appendix A, "A*".
//
// Work decomposition
//   CTA    = (TH x 32 pixel tile, group of S samples); TH = 4*WY
//   thread = 4 pixels in a column of the tile (k0 .. k0+3, same k1), S samples;
//            lanes run along k1 (memory-contiguous) so stores are coalesced and
//            lanes read detector bins <= 1.01 apart (conflict-free / broadcast).
//   smem   = for a chunk of AC angles the detector segment the tile projects
//            onto ([jlo_i, jlo_i+SEG) per angle, zero outside the detector) and
//            the per-angle constants re-centred on the tile (fp64 on entry), so
//            the fp32 index arithmetic works on small magnitudes.
//   angle range [lo,hi): the angle-sharded variant (multi-GPU config 4) is the
//            same kernel; partial images are summed by the caller (NCCL).
#include "scd_internal.cuh"
#include <algorithm>
#include <cmath>

#define SCD_MAGIC      12582912.0f
#define SCD_MAGIC_BITS 0x4B400000

struct BpParams {
    const float   *sino;
    float         *out;
    const BpAngle *bp;
    int n0, n1, n_angles, n_det, batch;
    int angle_lo, angle_hi;
    int AC, SEG;
    BpEpilogue ep;
};

template <int S> struct BpVec;
template <> struct BpVec<1> { typedef float  T; };
template <> struct BpVec<2> { typedef float2 T; };
template <> struct BpVec<4> { typedef float4 T; };

__device__ __forceinline__ void bp_tap(float (&a)[1], float l, float r, float wl, float w)
{ a[0] = fmaf(r, w, fmaf(l, wl, a[0])); }
__device__ __forceinline__ void bp_tap(float (&a)[2], float2 l, float2 r, float wl, float w)
{ a[0] = fmaf(r.x, w, fmaf(l.x, wl, a[0])); a[1] = fmaf(r.y, w, fmaf(l.y, wl, a[1])); }
__device__ __forceinline__ void bp_tap(float (&a)[4], float4 l, float4 r, float wl, float w)
{
    a[0] = fmaf(r.x, w, fmaf(l.x, wl, a[0])); a[1] = fmaf(r.y, w, fmaf(l.y, wl, a[1]));
    a[2] = fmaf(r.z, w, fmaf(l.z, wl, a[2])); a[3] = fmaf(r.w, w, fmaf(l.w, wl, a[3]));
}

template <int S, int WY>
__global__ void __launch_bounds__(32 * WY)
bp_pixel_kernel(const BpParams P)
{
    typedef typename BpVec<S>::T V;
    constexpr int TH = 4 * WY;
    constexpr int NT = 32 * WY;
    extern __shared__ __align__(16) float smem[];
    // layout: consts[AC] (float4) | jlo[AC] (int, padded to 4) | seg[AC][SEG][S] (samples interleaved)
    float4 *cst = reinterpret_cast<float4 *>(smem);
    int *jlo = reinterpret_cast<int *>(smem + 4 * P.AC);
    float *seg = smem + 4 * P.AC + ((P.AC + 3) & ~3);
    __shared__ float red[S][WY];

    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
    const int K0 = blockIdx.y * TH, K1 = blockIdx.x * 32;
    const int b0 = blockIdx.z * S;
    const int AC = P.AC, SEG = P.SEG;
    const size_t sino_sz = (size_t)P.n_angles * P.n_det;

    float acc[4][S];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int s = 0; s < S; ++s) acc[m][s] = 0.f;

    const float lx = (float)lane;
    const float ky0 = (float)(4 * wy);

    for (int a0 = P.angle_lo; a0 < P.angle_hi; a0 += AC) {
        const int nac = min(AC, P.angle_hi - a0);
        __syncthreads();
        // ---- per-angle constants re-centred on this tile (fp64) -------------
        for (int i = tid; i < nac; i += NT) {
            const BpAngle A = P.bp[a0 + i];
            const double v00 = A.ci * (double)K0 + A.si * (double)K1 + A.oi;
            const double vmin = v00 + fmin(0.0, A.ci * (double)(TH - 1)) + fmin(0.0, A.si * 31.0);
            const int j0 = (int)floor(vmin) - 1;   // one bin of slack below (fp32 rounding)
            jlo[i] = j0;
            // zf = v - j0 >= 1: floor(zf) = segment index of the left tap
            cst[i] = make_float4((float)A.ci, (float)A.si, (float)(v00 - (double)j0), 0.f);
        }
        __syncthreads();
        // ---- stage the detector segments, samples interleaved per bin ----------
        for (int row = wy; row < S * nac; row += WY) {
            const int s = row / nac, i = row - s * nac;
            const int b = b0 + s;
            const int j0 = jlo[i];
            const float *src = P.sino + (size_t)(b < P.batch ? b : 0) * sino_sz +
                               (size_t)(a0 + i) * P.n_det;
            float *dst = seg + (size_t)i * SEG * S + s;
            for (int e = lane; e < SEG; e += 32) {
                const int j = j0 + e;
                dst[e * S] = (b < P.batch && j >= 0 && j < P.n_det) ? __ldg(src + j) : 0.f;
            }
        }
        __syncthreads();
        // ---- accumulate ----------------------------------------------------
#pragma unroll 4
        for (int i = 0; i < nac; ++i) {
            const float4 c = cst[i];
            const float vb = fmaf(lx, c.y, c.z);
            const V *row = reinterpret_cast<const V *>(seg) + (size_t)i * SEG;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float z = fmaf(ky0 + (float)m, c.x, vb);
                const float t = __fadd_rd(z, SCD_MAGIC);                 // floor(z) in the mantissa
                const float w = z - (t - SCD_MAGIC);
                const float wl = 1.0f - w;
                const V *q = row + (__float_as_int(t) - SCD_MAGIC_BITS);
                bp_tap(acc[m], q[0], q[1], wl, w);
            }
        }
    }

    // ---- epilogue: axpy, second output, dot-product partials ---------------
    const BpEpilogue &E = P.ep;
    const int k1 = K1 + lane;
    float dsum[S];
#pragma unroll
    for (int s = 0; s < S; ++s) dsum[s] = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int b = b0 + s;
        if (b >= P.batch) continue;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int k0 = K0 + 4 * wy + m;
            if (k0 < P.n0 && k1 < P.n1) {
                const size_t o = ((size_t)b * P.n0 + k0) * P.n1 + k1;
                float v = E.c_acc * acc[m][s];
                float a1 = 0.f;
                if (E.add1) { a1 = E.add1[o]; v = fmaf(E.c1, a1, v); }
                if (E.add2) v = fmaf(E.c2, E.add2[o], v);
                P.out[o] = v;
                if (E.out2) E.out2[o] = v;
                dsum[s] += v * (E.dot_with_add1 ? a1 : v);
            }
        }
    }
    if (E.dot_part) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float v = dsum[s];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) red[s][wy] = v;
        }
        __syncthreads();
        if (tid < S) {
            const int b = b0 + tid;
            if (b < P.batch) {
                float v = 0.f;
                for (int w = 0; w < WY; ++w) v += red[tid][w];
                E.dot_part[(size_t)b * E.dot_stride + blockIdx.y * gridDim.x + blockIdx.x] = v;
            }
        }
    }
}

// ------------------------------------------------------------- host side ---
struct BpConfig { int S, WY, AC, SEG; size_t smem; dim3 grid; };

static BpConfig bp_choose(const scd_geom *g, int batch, int angle_lo, int angle_hi)
{
    BpConfig c;
    // measured on B200 (tools/kbench.py --kernel bp --sweep): 64 x 32 tiles are fastest at every
    // batch size; 4 interleaved samples per thread once the batch provides enough CTAs
    c.S = batch >= 32 ? 4 : (batch >= 16 ? 2 : 1);
    if (g->tune_bp_samples) c.S = g->tune_bp_samples;
    if (c.S != 1 && c.S != 2 && c.S != 4) c.S = 1;
    c.WY = 16;                                 // 64 x 32 tile, 512 threads
    if (g->tune_bp_tile) c.WY = g->tune_bp_tile / 4;
    if (c.WY != 4 && c.WY != 8 && c.WY != 16) c.WY = 16;
    const int TH = 4 * c.WY;
    double span = 0.0;
    for (int i = angle_lo; i < angle_hi; ++i)
        span = std::max(span, std::fabs(g->h_bp[i].ci) * (TH - 1) + std::fabs(g->h_bp[i].si) * 31.0);
    c.SEG = (int)std::ceil(span) + 6;         // taps j, j+1 plus one bin of slack either side
    const int na = std::max(1, angle_hi - angle_lo);
    // whole angle range in one chunk when it fits in ~64 KB, else chunks
    c.AC = na;
    const size_t budget = 64 * 1024;
    auto bytes = [&](int ac) { return (size_t)(4 * ac + ((ac + 3) & ~3) + (size_t)ac * c.S * c.SEG) * 4; };
    while (bytes(c.AC) > budget && c.AC > 8) c.AC = (c.AC + 1) / 2;
    c.smem = bytes(c.AC);
    c.grid = dim3((g->n1 + 31) / 32, (g->n0 + TH - 1) / TH, (batch + c.S - 1) / c.S);
    return c;
}

int scd_bp_ctas_per_sample_v1(const scd_geom *g, int batch)
{
    BpConfig c = bp_choose(g, batch, 0, g->n_angles);
    return (int)(c.grid.x * c.grid.y);
}

template <int S, int WY>
static int bp_launch_t(const BpParams &P, const BpConfig &c, cudaStream_t st, int device)
{
    static ScdSmemAttr attr = {};        // per instantiation
    SCD_CUDA(scd_ensure_smem(bp_pixel_kernel<S, WY>, attr, device, c.smem));
    bp_pixel_kernel<S, WY><<<c.grid, 32 * WY, c.smem, st>>>(P);
    SCD_LAUNCH_CHECK("bp_pixel_kernel");
    return 0;
}

int scd_launch_bp_v1(const scd_geom *g, const float *sino, float *out, int batch,
                     int angle_lo, int angle_hi, const BpEpilogue &ep, cudaStream_t st)
{
    if (!g || !sino || !out) { scd_set_error("scd_bp: null argument"); return SCD_E_INVALID; }
    if (batch < 0 || angle_lo < 0 || angle_hi > g->n_angles || angle_lo > angle_hi) {
        scd_set_error("scd_bp: bad batch/angle range (batch=%d, angles [%d,%d) of %d)",
                      batch, angle_lo, angle_hi, g->n_angles);
        return SCD_E_INVALID;
    }
    if (batch == 0) return 0;
    BpConfig c = bp_choose(g, batch, angle_lo, angle_hi);
    if (c.grid.z > 65535) { scd_set_error("scd_bp: batch too large"); return SCD_E_INVALID; }
    if (c.smem > (size_t)g->smem_optin) { scd_set_error("scd_bp: segment does not fit in shared memory"); return SCD_E_INVALID; }
    BpParams P;
    P.sino = sino; P.out = out; P.bp = g->d_bp;
    P.n0 = g->n0; P.n1 = g->n1; P.n_angles = g->n_angles; P.n_det = g->n_det; P.batch = batch;
    P.angle_lo = angle_lo; P.angle_hi = angle_hi; P.AC = c.AC; P.SEG = c.SEG; P.ep = ep;
#define BP_CASE(SS, WW) if (c.S == SS && c.WY == WW) return bp_launch_t<SS, WW>(P, c, st, g->device);
    BP_CASE(1, 4) BP_CASE(1, 8) BP_CASE(1, 16)
    BP_CASE(2, 4) BP_CASE(2, 8) BP_CASE(2, 16)
    BP_CASE(4, 4) BP_CASE(4, 8) BP_CASE(4, 16)
#undef BP_CASE
    scd_set_error("scd_bp: unsupported config S=%d WY=%d", c.S, c.WY);
    return SCD_E_INVALID;
}
