// K2c  peer-staged reduction of angle-sharded backprojections (multi-GPU config 4 of BASELINE.json).
//
// The reference has no multi-GPU path; north_star partitions one large slice stack
// by ANGLES: every GPU backprojects its angles into a full-size partial image stack and the partials are
// summed.  Besides the NCCL all-reduce (sharding.py, reduce='nccl') the sum can ride on the backprojector
// itself: the image rows are split into one band per GPU, bp_tile's epilogue stores every tile straight
// into the memory of the band's owner (peer memory mapped into this process, stores travel over
// NVLink / NVSwitch while the kernel is still computing other tiles), and the owner adds the P staged
// copies of its band in rank order (deterministic) and stores the result into every GPU's output:
//
//   scd_bp_banded / scd_bp_il_banded   bp_tile with banded (peer) output, no local output
//   scd_band_reduce                    out_p[rows of my band] = c_sum * sum_r stage[r] (+ c_add * addend), all p
//
// Cross-GPU ordering (all ranks have stored band b before its owner reduces it; all owners have stored
// before the result is read) is the caller's: sharding.py uses stream-ordered NCCL barriers.
#include "scd_internal.cuh"
#include <cstring>

#define PR_THREADS 256
#define PR_U 4

struct PrOut { float *p[SCD_MAX_BANDS]; };

__global__ void __launch_bounds__(PR_THREADS)
band_reduce_kernel(const float *stage, int n_src, int64_t slot_stride, int band_rows, int rows,
                   int n1, int row_lo, int n0, PrOut outs, int n_out, int multicast, const float *__restrict__ addend,
                   float c_add, float c_sum, int64_t total)
{
    scd_pdl_wait();
    scd_pdl_trigger();
    const int64_t per_slice = (int64_t)rows * n1;
    const int64_t stride = (int64_t)gridDim.x * PR_THREADS;
    for (int64_t i0 = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x; i0 < total; i0 += stride * PR_U) {
        float sum[PR_U];
        int64_t full[PR_U];
        bool live[PR_U];
        // the loads of the n_src staged copies of PR_U elements go out together
#pragma unroll
        for (int u = 0; u < PR_U; ++u) {
            const int64_t i = i0 + (int64_t)u * stride;
            live[u] = i < total;
            sum[u] = 0.f; full[u] = 0;
            if (live[u]) {
                const int64_t s = i / per_slice, rem = i - s * per_slice;
                const int64_t off = s * (int64_t)band_rows * n1 + rem;           // dense [slice][band_rows][n1]
                full[u] = (s * n0 + row_lo) * (int64_t)n1 + rem;                  // [slice][n0][n1]
                for (int r = 0; r < n_src; ++r) sum[u] += stage[(int64_t)r * slot_stride + off];   // rank order
            }
        }
#pragma unroll
        for (int u = 0; u < PR_U; ++u) {
            if (!live[u]) continue;
            float val = c_sum * sum[u];
            if (addend) val = fmaf(c_add, addend[full[u]], val);
            if (multicast) {
                // one store to the multicast address: the NVSwitch replicates it into every GPU's copy
                asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;\n" :: "l"(outs.p[0] + full[u]), "f"(val) : "memory");
            } else {
#pragma unroll
                for (int p = 0; p < SCD_MAX_BANDS; ++p)
                    if (p < n_out) outs.p[p][full[u]] = val;
            }
        }
    }
}

int scd_launch_band_reduce(const float *stage, int n_src, int64_t slot_stride, int batch, int band_rows, int rows,
                           int n1, int row_lo, int n0, float *const *out_ptrs, int n_out, int multicast,
                           const float *addend, float c_add, float c_sum, cudaStream_t st)
{
    if (batch <= 0 || rows <= 0) return 0;
    if (multicast && n_out != 1) { scd_set_error("scd_band_reduce: a multicast output is a single pointer"); return SCD_E_INVALID; }
    if (!stage || !out_ptrs || n_src <= 0 || n_out <= 0 || n_out > SCD_MAX_BANDS || rows > band_rows ||
        row_lo < 0 || row_lo + rows > n0 || n1 <= 0) {
        scd_set_error("scd_band_reduce: bad argument"); return SCD_E_INVALID;
    }
    PrOut o;
    memset(&o, 0, sizeof(o));
    for (int i = 0; i < n_out; ++i) {
        if (!out_ptrs[i]) { scd_set_error("scd_band_reduce: null output pointer"); return SCD_E_INVALID; }
        o.p[i] = out_ptrs[i];
    }
    const int64_t total = (int64_t)batch * rows * n1;
    int64_t blocks = (total + (int64_t)PR_THREADS * PR_U - 1) / ((int64_t)PR_THREADS * PR_U);
    if (blocks > 148 * 8) blocks = 148 * 8;
    SCD_CUDA(scd_launch_kernel(band_reduce_kernel, dim3((unsigned)blocks), dim3(PR_THREADS), 0, st, 0, stage, n_src, slot_stride,
                               band_rows, rows, n1, row_lo, n0, o, n_out, multicast, addend, c_add, c_sum, total));
    SCD_LAUNCH_CHECK("band_reduce_kernel");
    return 0;
}

static int pr_epilogue(BpEpilogue &e, float c_acc, float *const *band_ptrs, int n_bands, int band_rows)
{
    if (!band_ptrs || n_bands <= 0 || n_bands > SCD_MAX_BANDS) { scd_set_error("scd_bp_banded: 1..%d bands", SCD_MAX_BANDS); return SCD_E_INVALID; }
    e.c_acc = c_acc; e.add1 = nullptr; e.c1 = 0.f; e.add2 = nullptr; e.c2 = 0.f;
    e.out2 = nullptr; e.dot_part = nullptr; e.dot_stride = 0; e.dot_with_add1 = 0;
    e.n_bands = n_bands; e.band_rows = band_rows;
    for (int i = 0; i < n_bands; ++i) e.band_out[i] = band_ptrs[i];
    return 0;
}

extern "C" int scd_bp_il_banded(const scd_geom_t *g, const float *sino_il, int batch, int angle_lo, int angle_hi,
                                float c_acc, float *const *band_ptrs, int n_bands, int band_rows, void *stream)
{
    if (g && batch == 0) return 0;
    if (!g || !sino_il) { scd_set_error("scd_bp_il_banded: null argument"); return SCD_E_INVALID; }
    if (((uintptr_t)sino_il & 127) != 0) { scd_set_error("scd_bp_il_banded: sino_il must be 128-byte aligned"); return SCD_E_INVALID; }
    if (batch < 0 || angle_lo < 0 || angle_hi > g->n_angles || angle_lo > angle_hi) {
        scd_set_error("scd_bp_il_banded: bad batch/angle range"); return SCD_E_INVALID;
    }
    BpEpilogue e;
    const int rc = pr_epilogue(e, c_acc, band_ptrs, n_bands, band_rows);
    if (rc) return rc;
    return scd_launch_bp_il(g, sino_il, nullptr, batch, angle_lo, angle_hi, e, (cudaStream_t)stream);
}

extern "C" int scd_bp_banded(const scd_geom_t *g, const float *sino, int batch, int angle_lo, int angle_hi,
                             float c_acc, float *const *band_ptrs, int n_bands, int band_rows,
                             void *scratch, size_t scratch_bytes, void *stream)
{
    BpEpilogue e;
    const int rc = pr_epilogue(e, c_acc, band_ptrs, n_bands, band_rows);
    if (rc) return rc;
    return scd_launch_bp(g, sino, nullptr, batch, angle_lo, angle_hi, e, scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int scd_band_reduce(const float *stage, int n_src, int64_t slot_stride_floats, int batch, int band_rows,
                               int rows, int n1, int row_lo, int n0, float *const *out_ptrs, int n_out,
                               int out_is_multicast, const float *addend, float c_add, float c_sum, void *stream)
{
    return scd_launch_band_reduce(stage, n_src, slot_stride_floats, batch, band_rows, rows, n1, row_lo, n0, out_ptrs,
                                  n_out, out_is_multicast, addend, c_add, c_sum, (cudaStream_t)stream);
}
