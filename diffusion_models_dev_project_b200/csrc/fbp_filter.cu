// K7  ramp filter of the filtered back-projection.
//
// Replaces the detector-axis filtering of `SimpleTrafo.fbp` (reference src/physics/trafo.py:34,42,67); the
// recipe the reference's own source holds is `filter_sinogram` (src/physics/utils.py:11-33): zero-pad every
// sinogram row to P = max(64, 2^ceil(log2(2 N_s))), multiply its FFT with the "ramp" Fourier filter of
// torch-radon / scikit-image -- 2*Re(FFT(f)) with f the band-limited ramp of Kak & Slaney (eq. 61):
// f[0] = 1/4, f[n] = -1/(pi n)^2 for odd n, 0 for even n, mirrored -- transform back, crop, scale by
// pi/(2 N_theta).  Because P >= 2 N_s the circular convolution never wraps, so the recipe equals the
// LINEAR convolution of the row with h(n) = f[|n|], which is what this kernel evaluates directly in shared
// memory (no FFT library, no padding):
//
//     out[b][i][k] = scale * sum_j sino[b][i][j] * h(|k - j|)
//
// With scale = 1/ds (the recipe assumes unit detector spacing; for a detector cell of ds the discrete ramp
// carries 1/ds^2 and the Riemann sum ds) followed by the backprojector with weight dphi = pi/N_theta this
// is the recipe's 2 * pi/(2 N_theta) * sum_i lerp(...): fbp(A x) ~ x.
#include "scd_internal.cuh"

#define RF_THREADS 256

__global__ void __launch_bounds__(RF_THREADS)
ramp_filter_kernel(const float *__restrict__ sino, float *__restrict__ out, int n_det, float scale)
{
    extern __shared__ double smd[];
    double *h = smd;                                          // [n_det]  h(0 .. n_det-1), fp64 (see below)
    float *row = reinterpret_cast<float *>(smd + n_det);      // [n_det]
    scd_pdl_wait();
    scd_pdl_trigger();
    const size_t base = (size_t)blockIdx.x * n_det;
    for (int j = threadIdx.x; j < n_det; j += RF_THREADS) {
        row[j] = sino[base + j];
        double v = 0.0;
        if (j == 0) v = 0.25;
        else if (j & 1) { const double pn = 3.14159265358979323846 * (double)j; v = -1.0 / (pn * pn); }
        h[j] = v * (double)scale;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_det; k += RF_THREADS) {
        // offsets of the other parity only (h vanishes on even offsets except 0).  The ramp removes the mean of
        // the row, so the sum cancels by three to four orders of magnitude: the taps are kept and the sum is
        // accumulated in fp64 (fp32 taps alone cost 1e-5 of the result at 711 bins); the kernel runs once per
        // image, not per reverse step
        double a0 = (double)row[k] * h[0], a1 = 0.0;
        for (int d = 1; d < n_det; d += 2) {
            const double hv = h[d];
            const int jl = k - d, jr = k + d;
            if (jl >= 0) a0 = fma((double)row[jl], hv, a0);
            if (jr < n_det) a1 = fma((double)row[jr], hv, a1);
        }
        out[base + k] = (float)(a0 + a1);
    }
}

extern "C" int scd_ramp_filter(const scd_geom_t *g, const float *sino, float *out, int batch, void *stream)
{
    if (g && batch == 0) return 0;
    if (!g || !sino || !out) { scd_set_error("scd_ramp_filter: null argument"); return SCD_E_INVALID; }
    if (batch < 0) { scd_set_error("scd_ramp_filter: negative batch"); return SCD_E_INVALID; }
    const long rows = (long)batch * g->n_angles;
    if (rows > 0x7fffffffL) { scd_set_error("scd_ramp_filter: too many rows"); return SCD_E_INVALID; }
    const size_t smem = (size_t)g->n_det * (sizeof(double) + sizeof(float));
    static ScdSmemAttr attr = {};
    SCD_CUDA(scd_ensure_smem(ramp_filter_kernel, attr, g->device, smem));
    SCD_CUDA(scd_launch_kernel(ramp_filter_kernel, dim3((unsigned)rows), dim3(RF_THREADS), smem, (cudaStream_t)stream, 0,
                               sino, out, g->n_det, (float)(1.0 / g->ds)));
    SCD_LAUNCH_CHECK("ramp_filter_kernel");
    return 0;
}
