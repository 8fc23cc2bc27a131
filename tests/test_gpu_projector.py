"""GPU parity of A (fp_packq + fp_march) and A* (sino_pack + bp_tile) against the CPU oracle, through the C ABI.

Tolerance: 1e-4 relative L2 (BASELINE.json north_star); observed values are ~1e-6.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rt(im_shape, num_angles, **kw):
    from diffusion_models_dev_project_b200 import B200RayTrafo
    return B200RayTrafo(im_shape, num_angles, **kw)


def _phantoms(golden, n):
    d = golden('dds_256.npz')
    imgs = [d['gt_0'][0, 0], d['gt_1'][0, 0]]
    rng = np.random.default_rng(0)
    while len(imgs) < n:
        imgs.append(rng.random((256, 256), dtype=np.float32))
    return np.stack(imgs[:n])[:, None]


@pytest.mark.parametrize('batch', [1, 3, 16, 65])
def test_fp_256_matches_oracle(golden, batch):
    geom = O.OracleGeometry((256, 256), 60)
    rt = _rt((256, 256), 60)
    x = _phantoms(golden, min(batch, 4))
    x = np.concatenate([x] * ((batch + len(x) - 1) // len(x)))[:batch]
    x = x * np.linspace(0.5, 1.5, batch, dtype=np.float32)[:, None, None, None]
    y = rt(torch.from_numpy(x).cuda()).cpu().numpy()
    assert y.shape == (batch, 1, 60, 365)
    y_ref = O.fp(geom, x[:4])
    assert rel_l2(y[:4], y_ref) < TOL
    # remaining samples are scaled copies: linearity gives the check for free
    for b in range(4, batch):
        assert rel_l2(y[b] / np.float32(np.linspace(0.5, 1.5, batch)[b]),
                      y[b % 4] / np.float32(np.linspace(0.5, 1.5, batch)[b % 4])) < 1e-5


@pytest.mark.parametrize('batch', [1, 3, 16, 65])
def test_bp_256_matches_oracle(batch):
    geom = O.OracleGeometry((256, 256), 60)
    rt = _rt((256, 256), 60)
    rng = np.random.default_rng(1)
    y = rng.standard_normal((batch, 1, 60, 365)).astype(np.float32)
    x = rt.trafo_adjoint(torch.from_numpy(y).cuda()).cpu().numpy()
    assert x.shape == (batch, 1, 256, 256)
    n = min(batch, 3)
    assert rel_l2(x[:n], O.bp(geom, y[:n])) < TOL
    if batch > 3:
        assert rel_l2(x[-1], O.bp(geom, y[-1])) < TOL


@pytest.mark.parametrize('im_shape,num_angles', [((501, 501), 24), ((48, 80), 7), ((33, 17), 5), ((128, 128), 90)])
def test_fp_bp_other_shapes(im_shape, num_angles):
    geom = O.OracleGeometry(im_shape, num_angles)
    rt = _rt(im_shape, num_angles)
    rng = np.random.default_rng(2)
    x = rng.random((2, 1, *im_shape), dtype=np.float32)
    y = rt(torch.from_numpy(x).cuda()).cpu().numpy()
    assert y.shape[-2:] == geom.obs_shape
    assert rel_l2(y, O.fp(geom, x)) < TOL
    g = rng.standard_normal((2, 1, *geom.obs_shape)).astype(np.float32)
    z = rt.trafo_adjoint(torch.from_numpy(g).cuda()).cpu().numpy()
    assert rel_l2(z, O.bp(geom, g)) < TOL


@pytest.mark.parametrize('tune', [dict(fp_samples=1, fp_angles=1), dict(fp_samples=2, fp_angles=3, fp_cluster=2),
                                  dict(fp_samples=4, fp_angles=4, fp_rows=8, fp_cluster=1), dict(fp_samples=4, fp_angles=2, fp_rows=4, fp_cluster=4),
                                  dict(fp_samples=8, fp_angles=4, fp_rows=8, fp_cluster=8), dict(fp_samples=8, fp_angles=1, fp_rows=4, fp_cluster=1, fp_nbuf=2),
                                  dict(fp_samples=16, fp_angles=4, fp_rows=4, fp_cluster=1), dict(fp_samples=16, fp_angles=3, fp_rows=2, fp_cluster=2),
                                  dict(fp_samples=16, fp_angles=4, fp_rows=4, fp_threads=1024), dict(fp_samples=8, fp_angles=2, fp_rows=8, fp_threads=1024, fp_plan=1),
                                  dict(fp_samples=1, bp_tile=32), dict(fp_samples=2, bp_tile=16), dict(fp_samples=4, bp_tile=32),
                                  dict(fp_samples=8, bp_tile=32), dict(fp_samples=8, bp_tile=16), dict(fp_samples=8, bp_tile=8), dict(fp_samples=16, bp_tile=8), dict(fp_samples=16),
                                  dict(fp_samples=16, fp_angles=4, fp_rows=4, fp_threads=768), dict(fp_samples=8, fp_angles=3, fp_rows=8, fp_threads=768, fp_cluster=2), dict(fp_samples=4, fp_threads=768),
                                  dict(fp_samples=4, bp_tile=16), dict(fp_samples=2, bp_tile=32),
                                  dict(fp_samples=4, fp_source=1), dict(fp_samples=8, fp_source=1, fp_cluster=2), dict(fp_samples=16, fp_source=1, fp_rows=2),
                                  dict(fp_samples=1, fp_source=2), dict(fp_samples=1, fp_source=1, fp_angles=2), dict(fp_samples=4, fp_cls0=1), dict(fp_samples=8, fp_cls0=2), dict(fp_samples=16, fp_cls0=1, fp_rows=2), dict(fp_samples=16, fp_cls0=1, fp_cluster=2)])
def test_tuning_variants_agree(tune):
    """Every template instantiation (samples per thread, rays per thread, tile shape) computes the same thing."""
    geom = O.OracleGeometry((96, 96), 20)
    rt = _rt((96, 96), 20)
    rt.set_tuning('cuda', **tune)
    rng = np.random.default_rng(3)
    x = rng.random((5, 1, 96, 96), dtype=np.float32)
    g = rng.standard_normal((5, 1, *geom.obs_shape)).astype(np.float32)
    assert rel_l2(rt(torch.from_numpy(x).cuda()).cpu().numpy(), O.fp(geom, x)) < TOL
    assert rel_l2(rt.trafo_adjoint(torch.from_numpy(g).cuda()).cpu().numpy(), O.bp(geom, g)) < TOL


@pytest.mark.parametrize('batch', [0, 5, 17, 33])
def test_ragged_groups_and_empty_batch(batch):
    """Batches that do not fill the last sample group (and the empty batch): every sample correct."""
    geom = O.OracleGeometry((80, 64), 11)
    rt = _rt((80, 64), 11)
    rng = np.random.default_rng(6)
    x = rng.random((batch, 1, 80, 64), dtype=np.float32)
    y = rt(torch.from_numpy(x).cuda())
    assert y.shape == (batch, 1, *geom.obs_shape)
    g = rng.standard_normal((batch, 1, *geom.obs_shape)).astype(np.float32)
    z = rt.trafo_adjoint(torch.from_numpy(g).cuda())
    assert z.shape == (batch, 1, 80, 64)
    if batch:
        idx = sorted({0, batch // 2, batch - 1})
        assert rel_l2(y.cpu().numpy()[idx], O.fp(geom, x[idx])) < TOL
        assert rel_l2(z.cpu().numpy()[idx], O.bp(geom, g[idx])) < TOL
        v = torch.from_numpy(x).cuda()
        ref = v + 0.03 * rt.trafo_adjoint(rt(v))
        assert float((rt.normal_apply(v, 0.03) - ref).norm() / ref.norm()) < 1e-6


@pytest.mark.parametrize('im_shape,num_angles', [((1024, 1024), 3), ((16, 16), 1), ((700, 40), 4)])
def test_extreme_shapes(im_shape, num_angles):
    """Largest supported strips (1024^2: fewer samples per group), a single angle, a very elongated image."""
    geom = O.OracleGeometry(im_shape, num_angles)
    rt = _rt(im_shape, num_angles)
    rng = np.random.default_rng(7)
    x = rng.random((3, 1, *im_shape), dtype=np.float32)
    y = rt(torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel_l2(y, O.fp(geom, x)) < TOL
    g = rng.standard_normal((3, 1, *geom.obs_shape)).astype(np.float32)
    z = rt.trafo_adjoint(torch.from_numpy(g).cuda()).cpu().numpy()
    assert rel_l2(z, O.bp(geom, g)) < TOL


def test_custom_geometry_through_the_abi():
    """Anything scd_geom_create accepts: unsorted angles beyond pi, non-unit pixels, off-centre domain, a
    detector that neither covers the image nor is centred on it (rays that miss, pixels that see no bin)."""
    from diffusion_models_dev_project_b200.physics.geometry import ParallelBeamGeometry2D
    rng = np.random.default_rng(11)
    angles = np.array([0.3, 5.9, 2.2, 3.9, 1.1, 4.6, 2.9])
    g = ParallelBeamGeometry2D(n0=50, n1=70, x_min=-10.3, y_min=5.1, dx=0.7, angles=angles, n_det=90,
                               s_min=-31.0, ds=0.9)
    rt = _rt((50, 70), 7, geometry=g)
    og = O.OracleGeometry((50, 70), 7)
    og.x_min, og.y_min, og.dx = g.x_min, g.y_min, g.dx
    og.n_det, og.s_min, og.ds, og.angles = g.n_det, g.s_min, g.ds, angles
    og.obs_shape = (7, 90)
    og.adj_scale = rt.adj_scale
    for batch in (1, 6, 19):
        x = rng.random((batch, 1, 50, 70), dtype=np.float32)
        y = rt(torch.from_numpy(x).cuda()).cpu().numpy()
        assert rel_l2(y, O.fp(og, x)) < TOL
        s = rng.standard_normal((batch, 1, 7, 90)).astype(np.float32)
        z = rt.trafo_adjoint(torch.from_numpy(s).cuda()).cpu().numpy()
        assert rel_l2(z, O.bp(og, s)) < TOL
        v = torch.from_numpy(x).cuda()
        ref = v + 0.02 * rt.trafo_adjoint(rt(v))
        assert float((rt.normal_apply(v, 0.02) - ref).norm() / ref.norm()) < 1e-6


def test_known_answers_disc_and_ones():
    """Analytic line integrals: centred disc -> 2*sqrt(R^2-s^2); A*(1) = pi inside the FOV."""
    rt = _rt((256, 256), 60)
    geom = rt.geometry
    k = np.arange(256) - 128 + 0.5
    R = 80.0
    disc = ((k[:, None] ** 2 + k[None, :] ** 2) <= R * R).astype(np.float32)
    y = rt(torch.from_numpy(disc)[None, None].cuda()).cpu().numpy()[0, 0]
    s = geom.s_min + (np.arange(geom.n_det) + 0.5) * geom.ds
    exact = 2 * np.sqrt(np.clip(R * R - s * s, 0, None))
    inner = np.abs(s) < R - 3
    assert np.abs(y[:, inner] - exact[inner]).max() < 1.5          # pixelised disc edge
    assert abs(y[:, inner].mean() / exact[inner].mean() - 1) < 2e-3
    ones = torch.ones(1, 1, 60, 365).cuda()
    bp1 = rt.trafo_adjoint(ones).cpu().numpy()
    assert np.allclose(bp1, np.pi, rtol=1e-5)


def test_angle_ranges_partition():
    """Angle-sharded A / A*: the per-range results add up to the full operator (config 4)."""
    geom = O.OracleGeometry((64, 64), 24)
    rt = _rt((64, 64), 24)
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.random((3, 1, 64, 64), dtype=np.float32)).cuda()
    g = torch.from_numpy(rng.standard_normal((3, 1, *geom.obs_shape)).astype(np.float32)).cuda()
    full_y, full_x = rt(x), rt.trafo_adjoint(g)
    parts = [(0, 5), (5, 6), (6, 19), (19, 24)]
    acc_y = torch.zeros_like(full_y)
    acc_x = torch.zeros_like(full_x)
    for lo, hi in parts:
        yp = rt._fp(x, angle_range=(lo, hi))
        assert torch.equal(yp[..., lo:hi, :], full_y[..., lo:hi, :])
        assert float(yp[..., :lo, :].abs().max() if lo else 0) == 0
        acc_y += yp
        xp = rt._bp(g, rt.adj_scale, angle_range=(lo, hi))
        assert rel_l2(xp.cpu().numpy(), O.bp(geom, g.cpu().numpy(), angle_range=(lo, hi))) < TOL
        acc_x += xp
    assert torch.equal(acc_y, full_y)
    assert rel_l2(acc_x.cpu().numpy(), full_x.cpu().numpy()) < 1e-6


def test_normal_apply_composes_without_relayout():
    """v + gamma A*(A v) through the interleaved sinogram == the two public operators in sequence."""
    rt = _rt((96, 80), 18)
    gen = torch.Generator(device='cuda').manual_seed(1)
    for batch in (1, 3, 8, 21):
        v = torch.rand(batch, 1, 96, 80, device='cuda', generator=gen)
        ref = v + 0.05 * rt.trafo_adjoint(rt(v))
        out = rt.normal_apply(v, 0.05)
        assert float((out - ref).norm() / ref.norm()) < 1e-6
        part = rt.normal_apply(v, 0.05, angle_range=(4, 11), add_identity=False)
        refp = 0.05 * rt._bp(rt._fp(v, angle_range=(4, 11)), rt.adj_scale, angle_range=(4, 11))
        assert float((part - refp).norm() / refp.norm()) < 1e-6


def test_linearity_and_dot_product_full_size():
    """Size-independent properties at the bench size (batch 8): linearity, and <Ax,y> vs <x,A*y>."""
    rt = _rt((256, 256), 60)
    gen = torch.Generator(device='cuda').manual_seed(0)
    x1 = torch.rand(8, 1, 256, 256, device='cuda', generator=gen)
    x2 = torch.rand(8, 1, 256, 256, device='cuda', generator=gen)
    y = torch.randn(8, 1, 60, 365, device='cuda', generator=gen)
    lin = rt(2 * x1 - 3 * x2) - (2 * rt(x1) - 3 * rt(x2))
    assert float(lin.norm() / rt(x1).norm()) < 1e-5
    # A* is the pixel-driven backprojector, not the exact transpose: the weighted inner
    # products agree to the matched/unmatched gap (~1e-2), not to round-off
    c_w = rt.geometry.range_weight
    lhs = (rt(x1).double() * y.double()).sum().item() * c_w
    rhs = (x1.double() * rt.trafo_adjoint(y).double()).sum().item()
    assert abs(lhs - rhs) / abs(rhs) < 0.05
    # with a smooth sinogram the two adjoints agree closely
    ys = rt(x2)
    lhs = (rt(x1).double() * ys.double()).sum().item() * c_w
    rhs = (x1.double() * rt.trafo_adjoint(ys).double()).sum().item()
    assert abs(lhs - rhs) / abs(rhs) < 2e-3


def test_full_size_batch_independence_and_angle_shards():
    """BASELINE config 3 / 4 sizes through size-independent properties: a sample's result does not depend on
    the batch it travels in (256 samples vs alone), and the angle shards of the 501^2 / 1200-angle operator
    add up to the whole."""
    rt = _rt((256, 256), 60)
    gen = torch.Generator(device='cuda').manual_seed(9)
    x = torch.rand(256, 1, 256, 256, device='cuda', generator=gen)
    y = rt(x)
    z = rt.trafo_adjoint(y)
    for i in (0, 100, 255):
        yi = rt(x[i:i + 1])
        assert float((yi - y[i:i + 1]).norm() / yi.norm()) < 1e-6
        zi = rt.trafo_adjoint(y[i:i + 1])
        assert float((zi - z[i:i + 1]).norm() / zi.norm()) < 1e-6
    nz = rt.normal_apply(x, 0.01)
    assert float((nz - (x + 0.01 * z)).norm() / nz.norm()) < 1e-6
    del x, y, z, nz
    big = _rt((501, 501), 1200)
    xs = torch.rand(2, 1, 501, 501, device='cuda', generator=gen)
    ys = big(xs)
    full = big.trafo_adjoint(ys)
    acc = torch.zeros_like(full)
    for r in range(8):                                    # the 8-GPU sharding of config 4, on one device
        lo, hi = r * 150, (r + 1) * 150
        part = big._fp(xs, angle_range=(lo, hi))
        assert float((part[..., lo:hi, :] - ys[..., lo:hi, :]).norm() / ys[..., lo:hi, :].norm()) < 1e-6
        acc += big._bp(ys, big.adj_scale, angle_range=(lo, hi))
    assert float((acc - full).norm() / full.norm()) < 1e-6
    # A*(1) = pi inside the field of view also at this size
    ones = torch.ones(1, 1, 1200, big.obs_shape[1], device='cuda')
    assert torch.allclose(big.trafo_adjoint(ones), torch.full((1, 1, 501, 501), float(np.pi), device='cuda'), rtol=2e-5)


def test_result_independent_of_batch_grouping():
    """The group size (1..16 samples) and tile height depend on the batch; results agree to fp32 position
    rounding (1e-5 on a white-noise sinogram, the worst case for the interpolation)."""
    rt = _rt((96, 96), 30)
    gen = torch.Generator(device='cuda').manual_seed(12)
    x = torch.rand(21, 1, 96, 96, device='cuda', generator=gen)
    y = torch.randn(21, 1, *rt.obs_shape, device='cuda', generator=gen)
    yf, zf = rt(x), rt.trafo_adjoint(y)
    for b in (1, 2, 3, 5, 9):
        assert float((rt(x[:b]) - yf[:b]).norm() / yf[:b].norm()) < 1e-5
        assert float((rt.trafo_adjoint(y[:b]) - zf[:b]).norm() / zf[:b].norm()) < 1e-5


def test_flat_interface_and_shapes():
    rt = _rt((64, 64), 10)
    x = torch.rand(2, 3, 64, 64, device='cuda')
    y = rt(x)
    assert y.shape == (2, 3, 10, rt.obs_shape[1])
    yf = rt.trafo_flat(x.reshape(6, -1).T)
    assert yf.shape == (10 * rt.obs_shape[1], 6)
    assert torch.equal(yf.T.reshape(2, 3, 10, -1), y)
    xf = rt.trafo_adjoint_flat(yf)
    assert torch.equal(xf.T.reshape(2, 3, 64, 64), rt.trafo_adjoint(y))
    assert not hasattr(rt, 'resize')
    assert rt.angles.shape == (10,)


def test_errors_are_loud():
    rt = _rt((64, 64), 10)
    with pytest.raises(RuntimeError):
        rt(torch.rand(1, 1, 64, 64))                      # CPU tensor: no fallback
    with pytest.raises(TypeError):
        rt(torch.rand(1, 1, 64, 64, device='cuda', dtype=torch.float64))
    with pytest.raises(ValueError):
        rt(torch.rand(1, 1, 32, 64, device='cuda'))
    with pytest.raises(ValueError):
        rt.trafo_adjoint(torch.rand(1, 1, 10, 7, device='cuda'))


def test_host_buffer_entry_points():
    import ctypes as C
    from diffusion_models_dev_project_b200 import _lib
    geom = O.OracleGeometry((64, 64), 10)
    rt = _rt((64, 64), 10)
    h = rt._handle(torch.device('cuda'))
    x = np.random.default_rng(5).random((2, 64, 64), dtype=np.float32)
    y = np.empty((2, *geom.obs_shape), dtype=np.float32)
    _lib.check(h._lib.scd_fp_host(h.ptr, x.ctypes.data, y.ctypes.data, 2), 'scd_fp_host')
    assert rel_l2(y, O.fp(geom, x)) < TOL
    z = np.empty_like(x)
    _lib.check(h._lib.scd_bp_host(h.ptr, y.ctypes.data, z.ctypes.data, 2), 'scd_bp_host')
    assert rel_l2(z, O.bp(geom, y)) < TOL
    assert h._lib.scd_fp(h.ptr, None, None, 1, 0, 10, None, 0, None) == _lib.SCD_E_INVALID
    assert 'null' in _lib.last_error()


@pytest.mark.parametrize('im_shape,num_angles,batch', [((256, 256), 60, 8), ((256, 256), 60, 40), ((96, 70), 13, 3),
                                                        ((501, 501), 24, 5), ((64, 64), 9, 16), ((33, 17), 5, 4)])
def test_interleaved_image_path(im_shape, num_angles, batch):
    """Sample-interleaved images (batches >= 3): pack / unpack round trip, A through tensor copies from the image
    (no packed copy), A* writing the interleaved layout -- against the oracle and against the packed-copy path."""
    geom = O.OracleGeometry(im_shape, num_angles)
    rt = _rt(im_shape, num_angles)
    assert rt.il_supported(batch, 'cuda') and not rt.il_supported(2, 'cuda')
    assert rt.il_supported(1, 'cuda') == (im_shape[1] % 4 == 0)        # one sample per group: il image = reference layout
    rng = np.random.default_rng(11)
    x = torch.from_numpy(rng.random((batch, 1, *im_shape), dtype=np.float32)).cuda()
    x_il = rt._img_il(x)
    assert torch.equal(rt._img_from_il(x_il, x.shape[:-2]), x)
    n = min(batch, 3)
    y_ref = O.fp(geom, x[:n].cpu().numpy())
    y = rt(x)
    assert rel_l2(y[:n].cpu().numpy(), y_ref) < TOL
    assert rel_l2(y[-1].cpu().numpy(), O.fp(geom, x[-1].cpu().numpy())) < TOL
    rt.set_tuning('cuda', fp_source=1)
    y_packed = rt(x)
    rt.set_tuning('cuda', fp_source=0)
    assert rel_l2(y.cpu().numpy(), y_packed.cpu().numpy()) < 1e-6
    q = rt._fp_ilimg(x_il, batch)
    z = rt._img_from_il(rt._bp_ilimg(q, batch, rt.adj_scale), x.shape[:-2])
    assert rel_l2(z[:n].cpu().numpy(), O.bp(geom, y_ref)) < TOL
    assert rel_l2(z.cpu().numpy(), rt.trafo_adjoint(y).cpu().numpy()) < 1e-6
    z2 = rt._img_from_il(rt._bp_ilimg(q, batch, 0.5 * rt.adj_scale, addend_il=x_il, addend_scale=2.0), x.shape[:-2])
    assert rel_l2(z2.cpu().numpy(), (0.5 * z + 2.0 * x).cpu().numpy()) < 1e-6
    with pytest.raises(ValueError):
        rt._img_il(x[:2])


@pytest.mark.parametrize('im_shape,num_angles,batch', [((256, 256), 60, 16), ((256, 256), 60, 256), ((501, 501), 40, 20), ((96, 70), 13, 6)])
def test_pair_march_is_bit_identical_to_the_plain_march(im_shape, num_angles, batch):
    """bp_tile's pair march (3 loads for the 4 taps of a column pair; the lower pixel of the pair -- known from the
    sign of cos(phi), uniform per angle -- takes two taps, the upper one the three-bin form with an exact zero weight)
    against the plain two-taps-per-pixel march (tuning bp_share = 1): same bits."""
    rt = _rt(im_shape, num_angles)
    gen = torch.Generator(device='cuda').manual_seed(5)
    y = torch.randn(batch, 1, *rt.obs_shape, device='cuda', generator=gen)
    z = rt.trafo_adjoint(y)
    rt.set_tuning('cuda', bp_share=1)
    z_plain = rt.trafo_adjoint(y)
    rt.set_tuning('cuda', bp_share=0)
    assert torch.equal(z, z_plain)


@pytest.mark.parametrize('im_shape,num_angles', [((256, 256), 60), ((96, 72), 13), ((64, 200), 7), ((501, 500), 24)])
def test_single_sample_on_the_tensor_copy_path(im_shape, num_angles):
    """Batch 1 (BASELINE configs 1 and 5): with one sample per group the interleaved image is the reference layout,
    so A reads the caller's tensor directly (bulk rows behind 4 lead pixels, class 1 gathered in 32-byte pieces) --
    no packed copy, and CG runs 3 launches per iteration.  Against the oracle, the packed-copy path and the reference
    cg recurrences."""
    import diffusion_models_dev_project_b200 as pkg
    geom = O.OracleGeometry(im_shape, num_angles)
    rt = _rt(im_shape, num_angles)
    assert rt.il_supported(1, 'cuda')
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.random((1, 1, *im_shape), dtype=np.float32)).cuda()
    y = rt(x)
    assert rel_l2(y.cpu().numpy(), O.fp(geom, x.cpu().numpy())) < TOL
    rt.set_tuning('cuda', fp_source=2)                    # packed copy for single-sample groups
    y_packed = rt(x)
    rhs = x + 0.5
    sol_packed = pkg.cg(op=rt.normal_op(0.05), x=x, rhs=rhs, n_iter=3)
    rt.set_tuning('cuda', fp_source=0)
    assert rel_l2(y.cpu().numpy(), y_packed.cpu().numpy()) < 1e-6
    from diffusion_models_dev_project_b200 import fused
    fused.launch_count(reset=True)
    sol = pkg.cg(op=rt.normal_op(0.05), x=x, rhs=rhs, n_iter=3)
    assert fused.launch_count() == 2 + 3 * 3              # init (A, A*) + 3 x (A, A*, update): no pack, no copies
    assert rel_l2(sol.cpu().numpy(), sol_packed.cpu().numpy()) < 1e-5
    # unaligned view of the caller: falls back to a copy, same result
    xx = torch.zeros(im_shape[0] * im_shape[1] + 1, device='cuda')[1:].view(1, 1, *im_shape).copy_(x[0, 0])
    assert rel_l2(rt(xx).cpu().numpy(), y.cpu().numpy()) < 1e-6
