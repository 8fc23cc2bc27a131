"""CPU oracle of the data-consistency hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by
diffusion_models_dev_project_b200 (the product has no CPU path).

Contents
  * ctypes wrapper of ray_oracle.c (Joseph A, pixel-driven A*, exact J^T), built
    with gcc into oracle/_build/ (build());
  * an independent, vectorised scipy.sparse restatement of the same two operators
    (joseph_matrix, bp_matrix) -- the two restatements are checked against each
    other and against analytic line integrals in tests/test_oracle.py;
  * OracleRayTrafo: a CPU ray-trafo object holding those sparse matrices as
    torch.sparse tensors.  It is a port of how the reference's only CPU-runnable
    projector works -- MatmulRayTrafo, torch.sparse.mm with a COO matrix
    (reference src/physics/matmul_ray_trafo.py:107-126) -- extended to carry a
    separate backprojection matrix (the reference class multiplies with the exact
    transpose; the production A* is the pixel-driven backprojector);
  * ports of the reference's cg / apTweedy / ddim / DDS predictor on torch CPU
    tensors (ref_port_*), used as the CPU baseline on machines where
    /root/reference does not exist (the GPU box).  tests/test_golden.py pins these
    ports against outputs of the reference's own code (tests/golden/).

PARITY UNPINNED for the projector arithmetic itself: see ray_oracle.c.
"""
import ctypes as C
import os
import subprocess
from math import pi

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, '_build')
SO_PATH = os.path.join(BUILD_DIR, 'libray_oracle.so')


class _Geom(C.Structure):
    _fields_ = [('n0', C.c_int32), ('n1', C.c_int32), ('x_min', C.c_double), ('y_min', C.c_double),
                ('dx', C.c_double), ('n_angles', C.c_int32), ('angles', C.POINTER(C.c_double)),
                ('n_det', C.c_int32), ('s_min', C.c_double), ('ds', C.c_double), ('adj_scale', C.c_double)]


def build(force=False):
    src = os.path.join(HERE, 'ray_oracle.c')
    if not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= os.path.getmtime(src):
        return SO_PATH
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = ['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', '-o', SO_PATH + '.tmp', src, '-lm']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('gcc failed building the oracle:\n' + res.stdout + res.stderr)
    os.replace(SO_PATH + '.tmp', SO_PATH)
    return SO_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_fp.argtypes = [C.POINTER(_Geom), C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_bp.argtypes = [C.POINTER(_Geom), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.oracle_fp_transpose.argtypes = [C.POINTER(_Geom), C.c_void_p, C.c_void_p, C.c_int]
    return _lib


class OracleGeometry:
    """Geometry per reference src/physics/trafo.py:18-27."""

    def __init__(self, im_shape, num_angles, adj_scale=None):
        n0, n1 = int(im_shape[0]), int(im_shape[1])
        self.n0, self.n1 = n0, n1
        self.x_min = float((-n0) // 2)
        self.y_min = float((-n1) // 2)
        x_max, y_max = float(n0 // 2), float(n1 // 2)
        self.dx = (x_max - self.x_min) / n0
        assert abs(self.dx - (y_max - self.y_min) / n1) < 1e-12
        rho = max(np.hypot(x, y) for x in (self.x_min, x_max) for y in (self.y_min, y_max))
        self.n_det = 2 * int(np.ceil(rho / self.dx)) + 1
        self.s_min = -rho
        self.ds = 2 * rho / self.n_det
        self.n_angles = int(num_angles)
        self.angles = (np.arange(num_angles, dtype=np.float64) + 0.5) * (pi / num_angles)
        self.dphi = pi / num_angles
        self.adj_scale = self.dphi if adj_scale is None else float(adj_scale)
        self.im_shape = (n0, n1)
        self.obs_shape = (self.n_angles, self.n_det)

    def _c(self):
        ang = np.ascontiguousarray(self.angles)
        g = _Geom(self.n0, self.n1, self.x_min, self.y_min, self.dx, self.n_angles,
                  ang.ctypes.data_as(C.POINTER(C.c_double)), self.n_det, self.s_min, self.ds, self.adj_scale)
        g._keep = ang
        return g


def _as_batch(a, tail):
    a = np.ascontiguousarray(a, dtype=np.float32)
    lead = a.shape[:-2]
    assert a.shape[-2:] == tuple(tail), (a.shape, tail)
    return a.reshape(-1, *tail), lead


def fp(geom: OracleGeometry, img):
    x, lead = _as_batch(img, geom.im_shape)
    out = np.empty((x.shape[0], *geom.obs_shape), dtype=np.float32)
    g = geom._c()
    _load().oracle_fp(C.byref(g), x.ctypes.data, out.ctypes.data, x.shape[0])
    return out.reshape(*lead, *geom.obs_shape)


def bp(geom: OracleGeometry, sino, angle_range=None):
    y, lead = _as_batch(sino, geom.obs_shape)
    out = np.empty((y.shape[0], *geom.im_shape), dtype=np.float32)
    lo, hi = angle_range if angle_range is not None else (0, geom.n_angles)
    g = geom._c()
    _load().oracle_bp(C.byref(g), y.ctypes.data, out.ctypes.data, y.shape[0], lo, hi)
    return out.reshape(*lead, *geom.im_shape)


def fp_transpose(geom: OracleGeometry, sino):
    y, lead = _as_batch(sino, geom.obs_shape)
    out = np.empty((y.shape[0], *geom.im_shape), dtype=np.float32)
    g = geom._c()
    _load().oracle_fp_transpose(C.byref(g), y.ctypes.data, out.ctypes.data, y.shape[0])
    return out.reshape(*lead, *geom.im_shape)


# ------------------------------------------------- scipy.sparse restatement ---
def joseph_matrix(geom: OracleGeometry):
    """Joseph system matrix J, shape (n_angles*n_det, n0*n1), CSR float64.
    Row (i,j) = ray; <= 2 taps per marched row/column."""
    import scipy.sparse as sp
    n0, n1, nd = geom.n0, geom.n1, geom.n_det
    rows, cols, vals = [], [], []
    sj = geom.s_min + (np.arange(nd) + 0.5) * geom.ds
    for i, phi in enumerate(geom.angles):
        c, s = np.cos(phi), np.sin(phi)
        if abs(s) > abs(c):
            k0 = np.arange(n0)
            x = geom.x_min + (k0 + 0.5) * geom.dx
            u = ((sj[:, None] - x[None, :] * c) / s - geom.y_min) / geom.dx - 0.5      # [nd, n0]
            fl = np.floor(u); w = u - fl; k = fl.astype(np.int64)
            wt = geom.dx / abs(s)
            for kk, ww in ((k, 1.0 - w), (k + 1, w)):
                ok = (kk >= 0) & (kk < n1)
                jj, k0i = np.nonzero(ok)
                rows.append(i * nd + jj); cols.append(k0i * n1 + kk[ok]); vals.append(ww[ok] * wt)
        else:
            k1 = np.arange(n1)
            y = geom.y_min + (k1 + 0.5) * geom.dx
            u = ((sj[:, None] - y[None, :] * s) / c - geom.x_min) / geom.dx - 0.5      # [nd, n1]
            fl = np.floor(u); w = u - fl; k = fl.astype(np.int64)
            wt = geom.dx / abs(c)
            for kk, ww in ((k, 1.0 - w), (k + 1, w)):
                ok = (kk >= 0) & (kk < n0)
                jj, k1i = np.nonzero(ok)
                rows.append(i * nd + jj); cols.append(kk[ok] * n1 + k1i); vals.append(ww[ok] * wt)
    m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(geom.n_angles * nd, n0 * n1))
    return m.tocsr()


def bp_matrix(geom: OracleGeometry):
    """Pixel-driven backprojection matrix B (includes adj_scale), shape (n0*n1, n_angles*n_det)."""
    import scipy.sparse as sp
    n0, n1, nd = geom.n0, geom.n1, geom.n_det
    x = geom.x_min + (np.arange(n0) + 0.5) * geom.dx
    y = geom.y_min + (np.arange(n1) + 0.5) * geom.dx
    pix = (np.arange(n0)[:, None] * n1 + np.arange(n1)[None, :]).ravel()
    rows, cols, vals = [], [], []
    for i, phi in enumerate(geom.angles):
        t = (x[:, None] * np.cos(phi) + y[None, :] * np.sin(phi)).ravel()
        v = (t - geom.s_min) / geom.ds - 0.5
        fl = np.floor(v); w = v - fl; j = fl.astype(np.int64)
        for jj, ww in ((j, 1.0 - w), (j + 1, w)):
            ok = (jj >= 0) & (jj < nd)
            rows.append(pix[ok]); cols.append(i * nd + jj[ok]); vals.append(ww[ok] * geom.adj_scale)
    m = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(n0 * n1, geom.n_angles * nd))
    return m.tocsr()


def _to_torch_coo(m):
    import torch
    m = m.astype('float32').tocoo()
    idx = torch.stack([torch.from_numpy(m.row.astype(np.int64)), torch.from_numpy(m.col.astype(np.int64))])
    return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data), m.shape).coalesce()


class OracleRayTrafo:
    """CPU ray trafo on torch.sparse matrices: A = Joseph matrix, A* = pixel-driven
    backprojection matrix.  Port of the mechanics of the reference's MatmulRayTrafo
    (src/physics/matmul_ray_trafo.py:107-126: torch.sparse.mm on COO float32), with
    the 4-D <-> flat adapters of BaseRayTrafo (base_ray_trafo.py:75-81,138-146).
    Batches are handled as matrix columns."""

    def __init__(self, geom: OracleGeometry, matched_adjoint=False, odl_autograd=False):
        import torch  # noqa: F401
        self.geom = geom
        # odl_autograd: gradients follow ODL's OperatorFunction [3P] instead
        # of the exact matrix transposes torch.sparse.mm would give: d<g, A x>/dx = A*(g)/c_w and
        # d<h, A* y>/dy = c_w A(h) with c_w = dphi*ds/dx^2 -- what the reference's SimpleTrafo does
        # when its LoRA adaptation differentiates through trafo / trafo_adjoint.
        self.odl_autograd = bool(odl_autograd)
        self.c_w = geom.dphi * geom.ds / geom.dx ** 2
        self.im_shape = geom.im_shape
        self.obs_shape = geom.obs_shape
        J = joseph_matrix(geom)
        self.matrix = _to_torch_coo(J)
        if matched_adjoint:
            c_w = geom.dphi * geom.ds / geom.dx ** 2
            self.matrix_adj = _to_torch_coo((J.T * c_w).tocsr())
        else:
            self.matrix_adj = _to_torch_coo(bp_matrix(geom))
        self.angles = geom.angles

    def _mm(self, matrix, v, out_shape):
        import torch
        nb, nc = v.shape[:2]
        flat = v.reshape(nb * nc, -1).T.contiguous()
        return torch.sparse.mm(matrix, flat).T.reshape(nb, nc, *out_shape)

    def trafo(self, x):
        if self.odl_autograd:
            return _paired_functions()[0].apply(x, self)
        return self._mm(self.matrix, x, self.obs_shape)

    def trafo_adjoint(self, y):
        if self.odl_autograd:
            return _paired_functions()[1].apply(y, self)
        return self._mm(self.matrix_adj, y, self.im_shape)

    __call__ = trafo


_PAIRED = None


def _paired_functions():
    """(trafo, trafo_adjoint) autograd Functions with ODL's gradient pairing (built lazily: torch)."""
    global _PAIRED
    if _PAIRED is None:
        import torch

        class _Trafo(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, rt):
                ctx.rt = rt
                return rt._mm(rt.matrix, x.detach(), rt.obs_shape)

            @staticmethod
            def backward(ctx, g):
                return _Adjoint.apply(g, ctx.rt) / ctx.rt.c_w, None

        class _Adjoint(torch.autograd.Function):
            @staticmethod
            def forward(ctx, y, rt):
                ctx.rt = rt
                return rt._mm(rt.matrix_adj, y.detach(), rt.im_shape)

            @staticmethod
            def backward(ctx, g):
                return _Trafo.apply(g, ctx.rt) * ctx.rt.c_w, None

        _PAIRED = (_Trafo, _Adjoint)
    return _PAIRED


# ----------------------------------------- ports of the reference's torch code ---
def ref_port_alpha_bar(beta_min=1e-4, beta_max=0.02, num_steps=1000):
    """DDPM._compute_alpha_cumprod table (reference src/utils/sde.py:159-174)."""
    import torch
    betas = torch.from_numpy(np.linspace(beta_min, beta_max, num_steps, dtype=np.float64))
    betas = torch.cat([torch.zeros(1), betas], dim=0)
    return (1 - betas).cumprod(dim=0).to(torch.float32)


def ref_port_cg(op, x, rhs, n_iter=5):
    """reference src/utils/cg.py:11-39, statement by statement."""
    import torch
    r = rhs - op(x)
    p = r
    sq_old = torch.linalg.norm(r.reshape(r.shape[0], -1), dim=1) ** 2
    for _ in range(n_iter):
        d = op(p)
        inner = (p * d).sum(dim=[1, 2, 3])
        alpha = sq_old / inner
        x = x + alpha[:, None, None, None] * p
        r = r - alpha[:, None, None, None] * d
        sq_new = torch.linalg.norm(r.reshape(r.shape[0], -1), dim=1) ** 2
        beta = sq_new / sq_old
        sq_old = sq_new
        p = r + beta[:, None, None, None] * p
    return x


def ref_port_tweedie(s, x, abar, t):
    """apTweedy for DDPM (reference src/samplers/utils.py:370-378)."""
    ab = abar.index_select(0, t.long() + 1)
    div = ab.pow(.5)[:, None, None, None].pow(-1)
    std = (1. - ab).pow(.5)[:, None, None, None]
    return (x - s * std) * div


def ref_port_ddim(s, xhat, abar, t, tm1, eta, noise):
    """ddim DDPM branch (reference src/samplers/utils.py:356-368) with the noise passed in."""
    import torch
    m1 = abar.index_select(0, tm1.long() + 1).pow(.5)[:, None, None, None]
    m = abar.index_select(0, t.long() + 1).pow(.5)[:, None, None, None]
    tbeta = ((1 - m1.pow(2)) / (1 - m.pow(2))).pow(.5) * (1 - m.pow(2) * m1.pow(-2)).pow(.5)
    if any(tbeta.isnan()):
        tbeta = torch.zeros(*tbeta.shape)
    xhat = xhat * m1
    det = torch.sqrt(1 - m1.pow(2) - tbeta.pow(2) * eta ** 2) * s
    sto = eta * tbeta * noise
    return xhat + det + sto


def ref_port_dds_step(score, x, atb, abar, t, tm1, gamma, eta, n_iter, ray_trafo, noise=None):
    """DDS predictor body (reference src/samplers/utils.py:159-218)."""
    import torch
    with torch.no_grad():
        s = score(x, t)
        xhat0 = ref_port_tweedie(s, x, abar, t)
        op = lambda v: v + gamma * ray_trafo.trafo_adjoint(ray_trafo(v))
        xhat = ref_port_cg(op, xhat0, xhat0 + gamma * atb, n_iter)
        if noise is None:
            noise = torch.randn_like(xhat)
        return ref_port_ddim(s, xhat, abar, t, tm1, eta, noise), xhat0


# --------------------------------------------------------------- fbp ---
def ramp_fourier_filter(size):
    """The "ramp" Fourier filter torch-radon's ``FourierFilters.get(size, 'ramp')`` returns [3P; it is
    scikit-image's ``_get_fourier_filter``]: ``2 * Re(FFT(f))`` with ``f`` the band-limited ramp of
    Kak & Slaney eq. 61 on a circular grid of ``size`` samples -- the filter the reference multiplies with
    at src/physics/utils.py:25-27."""
    n = np.concatenate([np.arange(1, size / 2 + 1, 2, dtype=int), np.arange(size / 2 - 1, 0, -2, dtype=int)])
    f = np.zeros(size)
    f[0] = 0.25
    f[1::2] = -1.0 / (np.pi * n) ** 2
    return 2.0 * np.real(np.fft.fft(f))


def filter_sinogram(sino):
    """Restatement of the reference's ``filter_sinogram`` (src/physics/utils.py:11-33), statement by
    statement, in float64: pad the detector axis to ``max(64, 2^ceil(log2(2 N_s)))`` (:18-20), FFT (:22),
    multiply with the ramp filter (:24-26), inverse FFT (:28), drop the padding and scale by
    ``pi / (2 N_theta)`` (:30)."""
    sino = np.asarray(sino, dtype=np.float64)
    size, n_angles = sino.shape[-1], sino.shape[-2]
    padded_size = max(64, int(2 ** np.ceil(np.log2(2 * size))))
    pad = padded_size - size
    padded = np.concatenate([sino, np.zeros(sino.shape[:-1] + (pad,))], axis=-1)
    spec = np.fft.fft(padded, axis=-1) * ramp_fourier_filter(padded_size)
    filtered = np.fft.ifft(spec, axis=-1)[..., :-pad] * (np.pi / (2 * n_angles))
    return filtered.real


def fbp(geom: OracleGeometry, sino):
    """Filtered back-projection as the reference's iradon branch composes it (src/physics/trafo.py:42:
    ``backprojection(filter_sinogram(x))``): the recipe above followed by the plain pixel-driven sum of
    interpolated detector values.  The recipe is written for a unit detector cell (torch-radon's
    ``det_spacing = 1``); on the ODL detector of this geometry (cell ``ds``) the discrete ramp carries
    ``1/ds^2`` and the Riemann sum ``ds``, so the result is divided by ``ds`` -- the only factor added to the
    reference recipe, needed for ``fbp(A x) ~ x`` [the ODL branch, ``odl.tomo.fbp_op``, is 3P and unpinned]."""
    q = filter_sinogram(sino) / geom.ds
    plain = OracleGeometry(geom.im_shape, geom.n_angles, adj_scale=1.0)
    return bp(plain, q.astype(np.float32))
