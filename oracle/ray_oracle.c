/*
 * ray_oracle.c -- CPU restatement of the reference's ray-transform arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
 * executed by the product package (diffusion_models_dev_project_b200/); it is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, as the checker and the timed CPU baseline.
 *
 * What it restates.  The reference's projector arithmetic is not in the
 * reference repository: SimpleTrafo (src/physics/trafo.py:17-34, :58, :61)
 * delegates to odl.tomo.RayTransform(impl='astra_cuda') and its .adjoint, i.e.
 * to ASTRA's 2-D CUDA kernels par_fp.cu / par_bp.cu.  Neither odl nor
 * astra-toolbox is vendored, installed or version-pinned by the reference
 * (no requirements file), and the reference ships no tests or golden
 * sinograms -> PARITY UNPINNED for A and A* (SURVEY.md section 8c).  This file
 * follows the published algorithm of those kernels as written out in
 * SURVEY.md Appendix A:
 *   A   Joseph's method: march along the dominant axis, linear interpolation
 *       across it, weight dx/max(|cos|,|sin|), zero outside the image;
 *   A*  pixel-driven backprojection: linear interpolation of each angle's
 *       detector row at t = x cos(phi) + y sin(phi), times adj_scale.
 * Geometry (domain, cell-centred angles, detector partition) follows
 * src/physics/trafo.py:18-27 + odl.tomo.parallel_beam_geometry.
 * It is pinned instead against analytic line integrals (disc, square) and
 * against an independent scipy.sparse restatement (oracle/oracle.py), and the
 * sparse matrices feed the reference's own MatmulRayTrafo / cg / sampler code
 * when the golden fixtures are generated (tests/golden/make_golden.py).
 *
 * All accumulation in double; inputs/outputs float32.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct {
    int32_t n0, n1;
    double x_min, y_min, dx;
    int32_t n_angles;
    const double *angles;
    int32_t n_det;
    double s_min, ds;
    double adj_scale;
} oracle_geom;

static inline double pix(const float *f, int n0, int n1, long k0, long k1)
{
    if (k0 < 0 || k0 >= n0 || k1 < 0 || k1 >= n1) return 0.0;
    return (double)f[(size_t)k0 * n1 + k1];
}

/* sino[b][i][j] = (A img[b])[i][j] */
void oracle_fp(const oracle_geom *g, const float *img, float *sino, int batch)
{
    const int n0 = g->n0, n1 = g->n1;
    const size_t isz = (size_t)n0 * n1, ssz = (size_t)g->n_angles * g->n_det;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int b = 0; b < batch; ++b) {
        for (int i = 0; i < g->n_angles; ++i) {
            const float *f = img + b * isz;
            float *out = sino + b * ssz + (size_t)i * g->n_det;
            const double c = cos(g->angles[i]), s = sin(g->angles[i]);
            for (int j = 0; j < g->n_det; ++j) {
                const double sj = g->s_min + (j + 0.5) * g->ds;
                double acc = 0.0;
                if (fabs(s) > fabs(c)) {
                    /* ray mostly along x: one tap pair per k0 */
                    for (int k0 = 0; k0 < n0; ++k0) {
                        const double x = g->x_min + (k0 + 0.5) * g->dx;
                        const double y = (sj - x * c) / s;
                        const double u = (y - g->y_min) / g->dx - 0.5;
                        const double fl = floor(u);
                        const double w = u - fl;
                        const long k = (long)fl;
                        acc += (1.0 - w) * pix(f, n0, n1, k0, k) + w * pix(f, n0, n1, k0, k + 1);
                    }
                    acc *= g->dx / fabs(s);
                } else {
                    for (int k1 = 0; k1 < n1; ++k1) {
                        const double y = g->y_min + (k1 + 0.5) * g->dx;
                        const double x = (sj - y * s) / c;
                        const double u = (x - g->x_min) / g->dx - 0.5;
                        const double fl = floor(u);
                        const double w = u - fl;
                        const long k = (long)fl;
                        acc += (1.0 - w) * pix(f, n0, n1, k, k1) + w * pix(f, n0, n1, k + 1, k1);
                    }
                    acc *= g->dx / fabs(c);
                }
                out[j] = (float)acc;
            }
        }
    }
}

/* img[b][k0][k1] = adj_scale * sum_i lerp(sino[b][i][.], v_i(k0,k1)), angles in [lo,hi) */
void oracle_bp(const oracle_geom *g, const float *sino, float *img, int batch, int angle_lo, int angle_hi)
{
    const int n0 = g->n0, n1 = g->n1, nd = g->n_det;
    const size_t isz = (size_t)n0 * n1, ssz = (size_t)g->n_angles * nd;
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int b = 0; b < batch; ++b) {
        for (int k0 = 0; k0 < n0; ++k0) {
            const double x = g->x_min + (k0 + 0.5) * g->dx;
            for (int k1 = 0; k1 < n1; ++k1) {
                const double y = g->y_min + (k1 + 0.5) * g->dx;
                double acc = 0.0;
                for (int i = angle_lo; i < angle_hi; ++i) {
                    const double t = x * cos(g->angles[i]) + y * sin(g->angles[i]);
                    const double v = (t - g->s_min) / g->ds - 0.5;
                    const double fl = floor(v);
                    const double w = v - fl;
                    const long j = (long)fl;
                    const float *row = sino + b * ssz + (size_t)i * nd;
                    const double g0 = (j >= 0 && j < nd) ? (double)row[j] : 0.0;
                    const double g1 = (j + 1 >= 0 && j + 1 < nd) ? (double)row[j + 1] : 0.0;
                    acc += (1.0 - w) * g0 + w * g1;
                }
                img[b * isz + (size_t)k0 * n1 + k1] = (float)(g->adj_scale * acc);
            }
        }
    }
}

/* exact transpose of the Joseph matrix: img = J^T sino (for dot tests; scatter, single thread per sample) */
void oracle_fp_transpose(const oracle_geom *g, const float *sino, float *img, int batch)
{
    const int n0 = g->n0, n1 = g->n1;
    const size_t isz = (size_t)n0 * n1, ssz = (size_t)g->n_angles * g->n_det;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < batch; ++b) {
        float *f = img + b * isz;
        for (size_t k = 0; k < isz; ++k) f[k] = 0.f;
        for (int i = 0; i < g->n_angles; ++i) {
            const double c = cos(g->angles[i]), s = sin(g->angles[i]);
            for (int j = 0; j < g->n_det; ++j) {
                const double sj = g->s_min + (j + 0.5) * g->ds;
                const double val = (double)sino[b * ssz + (size_t)i * g->n_det + j];
                if (fabs(s) > fabs(c)) {
                    const double wt = val * g->dx / fabs(s);
                    for (int k0 = 0; k0 < n0; ++k0) {
                        const double x = g->x_min + (k0 + 0.5) * g->dx;
                        const double u = ((sj - x * c) / s - g->y_min) / g->dx - 0.5;
                        const double fl = floor(u), w = u - fl;
                        const long k = (long)fl;
                        if (k >= 0 && k < n1) f[(size_t)k0 * n1 + k] += (float)((1.0 - w) * wt);
                        if (k + 1 >= 0 && k + 1 < n1) f[(size_t)k0 * n1 + k + 1] += (float)(w * wt);
                    }
                } else {
                    const double wt = val * g->dx / fabs(c);
                    for (int k1 = 0; k1 < n1; ++k1) {
                        const double y = g->y_min + (k1 + 0.5) * g->dx;
                        const double u = ((sj - y * s) / c - g->x_min) / g->dx - 0.5;
                        const double fl = floor(u), w = u - fl;
                        const long k = (long)fl;
                        if (k >= 0 && k < n0) f[(size_t)k * n1 + k1] += (float)((1.0 - w) * wt);
                        if (k + 1 >= 0 && k + 1 < n0) f[(size_t)(k + 1) * n1 + k1] += (float)(w * wt);
                    }
                }
            }
        }
    }
}
