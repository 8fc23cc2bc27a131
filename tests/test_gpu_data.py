"""GPU parity of the rows either side of the hot path and of the projector at the
full sizes of BASELINE configs 1 and 4: fbp, simulate, get_data_from_ground_truth, the 10 vendored phantoms,
501^2 with 1200 angles."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pkg():
    import diffusion_models_dev_project_b200 as pkg
    return pkg


def _vendored_phantoms(golden):
    """The 10 images of the reference's dataset/disk_ellipses_val_256.pt, baked into the config-1 fixture."""
    d = golden('config1_256.npz')
    return np.stack([d['gt_%d' % i][0] for i in range(int(d['n_images']))])       # [10,1,256,256]


def test_fp_bp_on_all_vendored_phantoms(golden):
    x = _vendored_phantoms(golden)
    assert x.shape == (10, 1, 256, 256)
    geom = O.OracleGeometry((256, 256), 60)
    rt = _pkg().B200RayTrafo((256, 256), 60)
    y = rt(torch.from_numpy(x).cuda())
    y_ref = O.fp(geom, x)
    for i in range(10):
        assert rel_l2(y[i].cpu().numpy(), y_ref[i]) < TOL, i
    z = rt.trafo_adjoint(y).cpu().numpy()
    z_ref = O.bp(geom, y.cpu().numpy())
    for i in range(10):
        assert rel_l2(z[i], z_ref[i]) < TOL, i


def test_fp_bp_501_dense_view_matches_oracle():
    """BASELINE config 4 geometry: 501^2 (off-centre domain [-251, 250]^2), 711 bins, 1200 angles -- the
    operator itself against the oracle on two slices, and one 150-angle shard of it (the 8-GPU split)."""
    geom = O.OracleGeometry((501, 501), 1200)
    rt = _pkg().B200RayTrafo((501, 501), 1200)
    assert rt.obs_shape == (1200, 711)
    rng = np.random.default_rng(3)
    xx, yy = np.meshgrid(np.arange(501) - 250.0, np.arange(501) - 250.0, indexing='ij')
    x = np.stack([((xx / 180) ** 2 + (yy / 120) ** 2 < 1).astype(np.float32) + 0.1 * rng.random((501, 501), dtype=np.float32),
                  rng.random((501, 501), dtype=np.float32)])[:, None]
    y = rt(torch.from_numpy(x).cuda())
    y_ref = O.fp(geom, x)
    assert rel_l2(y.cpu().numpy(), y_ref) < TOL
    ys = rng.standard_normal((1, 1, 1200, 711)).astype(np.float32)
    z = rt.trafo_adjoint(torch.from_numpy(ys).cuda()).cpu().numpy()
    assert rel_l2(z, O.bp(geom, ys)) < TOL
    zs = rt._bp(torch.from_numpy(ys).cuda(), rt.adj_scale, angle_range=(450, 600)).cpu().numpy()
    assert rel_l2(zs, O.bp(geom, ys, angle_range=(450, 600))) < TOL
    yl = rt._fp(torch.from_numpy(x).cuda(), angle_range=(450, 600)).cpu().numpy()
    assert rel_l2(yl[..., 450:600, :], y_ref[..., 450:600, :]) < TOL
    assert float(np.abs(yl[..., :450, :]).sum() + np.abs(yl[..., 600:, :]).sum()) == 0.0


@pytest.mark.parametrize('im_shape,num_angles,batch', [((256, 256), 60, 3), ((96, 70), 25, 2), ((501, 501), 200, 1)])
def test_fbp_matches_oracle_recipe(golden, im_shape, num_angles, batch):
    """scd_ramp_filter + the backprojector against oracle.fbp, the restatement of the reference's recipe
    (src/physics/utils.py:11-33 + trafo.py:42)."""
    pkg = _pkg()
    geom = O.OracleGeometry(im_shape, num_angles)
    rt = pkg.B200RayTrafo(im_shape, num_angles)
    rng = np.random.default_rng(7)
    if im_shape == (256, 256):
        x = _vendored_phantoms(golden)[:batch]
    else:
        x = rng.random((batch, 1, *im_shape), dtype=np.float32)
    y = O.fp(geom, x) + 0.05 * rng.standard_normal((batch, 1, *geom.obs_shape)).astype(np.float32)
    yt = torch.from_numpy(y).cuda()
    q = rt.ramp_filter(yt).cpu().numpy()
    q_ref = O.filter_sinogram(y) / geom.ds / geom.dphi          # the kernel leaves dphi to the backprojector
    assert rel_l2(q, q_ref) < 1e-5
    rec = rt.fbp(yt).cpu().numpy()
    assert rec.shape == (batch, 1, *im_shape)
    assert rel_l2(rec, O.fbp(geom, y)) < TOL
    assert torch.equal(rt.fbp(yt[0, 0]), rt.fbp(yt)[0, 0])       # any leading shape


def test_fbp_reconstructs_the_phantoms(golden):
    pkg = _pkg()
    x = torch.from_numpy(_vendored_phantoms(golden)).cuda()
    rt = pkg.B200RayTrafo((256, 256), 720)
    rec = rt.fbp(rt(x))
    for i in range(10):
        assert pkg.PSNR(rec[i, 0].cpu().numpy(), x[i, 0].cpu().numpy()) > 30.0


def test_simulate_on_the_gpu_reproduces_the_reference_measurements(golden):
    """`simulate` with the CUDA operator against the y_i made by the REFERENCE's simulate on the oracle operator
    (same seeded numpy stream; the operators agree to ~1e-6, so do the noise levels)."""
    pkg = _pkg()
    rt = pkg.B200RayTrafo((256, 256), 60)
    d = golden('config1_256.npz')
    for i in range(int(d['n_images'])):
        gt = torch.from_numpy(d['gt_%d' % i]).cuda()
        y = pkg.simulate(gt, rt, 0.01, rng=np.random.default_rng(1 + i))
        assert y.is_cuda and tuple(y.shape) == (1, 1, 60, 365)
        assert rel_l2(y.cpu().numpy(), d['y_%d' % i]) < 1e-5, i
    y2, level = pkg.simulate(gt, rt, 0.01, rng=np.random.default_rng(10), return_noise_level=True)
    assert torch.equal(y2, y) and abs(level - 0.01 * float(rt(gt).abs().mean())) < 1e-9


def test_get_data_from_ground_truth_and_dataset_on_the_gpu(golden):
    pkg = _pkg()
    rt = pkg.B200RayTrafo((256, 256), 60)
    x = torch.from_numpy(_vendored_phantoms(golden)[:3])
    gt, obs, fbp = pkg.get_data_from_ground_truth(x[0].cuda(), rt, 0.0)
    assert tuple(gt.shape) == (1, 1, 256, 256) and tuple(obs.shape) == (1, 1, 60, 365) and tuple(fbp.shape) == (1, 1, 256, 256)
    assert torch.equal(obs, rt(gt)) and torch.equal(fbp, rt.fbp(obs))
    ds = pkg.SimulatedDataset([x[i] for i in range(3)], rt, 0.01, device='cuda')
    d = golden('config1_256.npz')
    for i, (y, xi, f) in enumerate(ds):
        assert rel_l2(y.cpu().numpy(), d['y_%d' % i][0]) < 1e-5          # seeds 1 + idx, as run_conditional_sampling
        assert torch.equal(xi.cpu(), x[i]) and tuple(f.shape) == (1, 256, 256)
        assert pkg.PSNR(f[0].cpu().numpy(), x[i, 0].numpy()) > 15.0


def test_config1_psnr_gate_small_unet_all_phantoms(golden, monkeypatch):
    """BASELINE config 1 as SURVEY.md section 8(d) specifies it: the 10 vendored phantoms, batch 1, 100 DDIM steps,
    CG(5), gamma 0.01, eta 0.15, seeds 1 + i, the small ADM UNet (random init, same construction seed) as score
    model -- against the chain of the reference's own BaseSampler + DDS predictor on CPU
    (tests/golden/make_golden_config1.py).  Gate: PSNR within 0.1 dB per image."""
    import functools
    pkg = _pkg()
    from bench_support.adm_unet import small_unet
    monkeypatch.setattr(torch, 'randn_like', lambda t, **kw: torch.randn(t.shape, dtype=t.dtype).to(t.device))
    monkeypatch.setattr(torch.backends.cudnn, 'allow_tf32', False)
    monkeypatch.setattr(torch.backends.cuda.matmul, 'allow_tf32', False)
    d = golden('config1_256.npz')
    rt = pkg.B200RayTrafo((256, 256), 60)
    sde = pkg.DDPM()
    torch.manual_seed(0)
    score = small_unet().eval().cuda()
    worst_db, worst_l2 = 0.0, 0.0
    for i in range(int(d['n_images'])):
        gt = d['gt_%d' % i]
        y = torch.from_numpy(d['y_%d' % i]).cuda()
        kw = {'num_steps': int(d['num_steps']), 'batch_size': 1, 'start_time_step': 0, 'im_shape': [1, 256, 256],
              'eps': 1e-3, 'travel_length': 1, 'travel_repeat': 1,
              'predictor': {'eta': float(d['eta']), 'gamma': float(d['gamma']), 'use_simplified_eqn': True,
                            'ray_trafo': rt}}
        predictor = functools.partial(pkg.decomposed_diffusion_sampling_sde_predictor, score=score, sde=sde,
                                      rhs=rt.trafo_adjoint(y), cg_kwargs={'max_iter': int(d['cg_iter'])})
        sampler = pkg.BaseSampler(score=score, sde=sde, predictor=predictor, sample_kwargs=kw, device='cuda')
        torch.manual_seed(1 + i)
        recon = sampler.sample(logging=False).cpu().numpy()
        psnr = pkg.PSNR(recon[0, 0], gt[0, 0])
        worst_db = max(worst_db, abs(psnr - float(d['psnr_%d' % i])))
        blk = recon.reshape(1, 1, 64, 4, 64, 4).mean(axis=(3, 5))
        worst_l2 = max(worst_l2, rel_l2(blk, d['recon_blk_%d' % i]))
        if 'recon_%d' % i in d.files:
            worst_l2 = max(worst_l2, rel_l2(recon, d['recon_%d' % i]))
    print('config 1: worst |dPSNR| %.4f dB, worst rel L2 %.2e' % (worst_db, worst_l2))
    assert worst_db < 0.1, worst_db
    assert worst_l2 < 1e-2, worst_l2
