"""Two-GPU tests (NCCL): angle-sharded A / A* with all-reduce, and sample-sharded DDS steps.

Skipped on a box with fewer than two GPUs; the same logic is covered on CPU over gloo by
tests/test_sharding_cpu.py.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import diffusion_models_dev_project_b200 as pkg
        from diffusion_models_dev_project_b200.sharding import AngleShardedRayTrafo, shard_range
        torch.set_grad_enabled(False)
        rt = pkg.B200RayTrafo((96, 96), 30)
        sh = AngleShardedRayTrafo(rt, chunk=3)
        lo, hi = sh.angle_range
        assert (lo, hi) == shard_range(30, rank, world)
        gen = torch.Generator(device=dev).manual_seed(0)          # replicated vectors
        x = torch.rand(7, 1, 96, 96, device=dev, generator=gen)
        y = torch.randn(7, 1, *rt.obs_shape, device=dev, generator=gen)
        full = rt(x)
        yl = sh(x)
        assert torch.equal(yl[..., lo:hi, :], full[..., lo:hi, :])
        assert float(yl[..., :lo, :].abs().sum() + yl[..., hi:, :].abs().sum()) == 0.0
        assert float((sh.gather_sinogram(yl) - full).norm() / full.norm()) < 1e-6
        # the sharded call runs in slice chunks of 3 (4 samples per group, 16-row tiles), the reference in one
        # batch of 7 (8 per group, 8-row tiles): the fp32 tap positions are rounded relative to different tile
        # origins, which shows at the 1e-6 level on a white-noise sinogram
        z = sh.trafo_adjoint(y)
        zf = rt.trafo_adjoint(y)
        assert float((z - zf).norm() / zf.norm()) < 1e-5
        gamma = 0.05
        sol = pkg.cg(op=sh.normal_op(gamma), x=x, rhs=x + 1.0, n_iter=3)
        ref = pkg.cg(op=rt.normal_op(gamma), x=x, rhs=x + 1.0, n_iter=3)
        assert float((sol - ref).norm() / ref.norm()) < 1e-4
        gathered = [torch.empty_like(sol) for _ in range(world)]
        dist.all_gather(gathered, sol)
        assert all(torch.equal(gathered[0], t) for t in gathered)     # replicas stay identical
        # the same with the sum riding on the backprojector (peer-staged bands instead of the all-reduce)
        shp = AngleShardedRayTrafo(rt, chunk=3, reduce='peer')
        zp = shp.trafo_adjoint(y)
        assert float((zp - zf).norm() / zf.norm()) < 1e-5
        # the banded launch keeps whole 16-row tiles (bands are tile-aligned) while the plain launch levels the
        # SMs with fewer rows per tile: the tap positions are rounded relative to different tile origins
        assert float((zp - z).norm() / z.norm()) < 1e-5, float((zp - z).norm() / z.norm())
        zp2 = shp.trafo_adjoint(y)
        assert torch.equal(zp, zp2)                                # deterministic
        na = shp.normal_apply(x, gamma)
        nb = sh.normal_apply(x, gamma)
        assert float((na - nb).norm() / nb.norm()) < 1e-5, float((na - nb).norm() / nb.norm())
        solp = pkg.cg(op=shp.normal_op(gamma), x=x, rhs=x + 1.0, n_iter=3)
        assert float((solp - ref).norm() / ref.norm()) < 1e-4
        gathered = [torch.empty_like(solp) for _ in range(world)]
        dist.all_gather(gathered, solp.contiguous())
        assert all(torch.equal(gathered[0], t) for t in gathered)
        # a stack that is not a multiple of the chunk, more chunks than staging buffers, odd image height
        rt2 = pkg.B200RayTrafo((70, 50), 9)
        sh2n, sh2p = AngleShardedRayTrafo(rt2, chunk=2), AngleShardedRayTrafo(rt2, chunk=2, reduce='peer')
        y2 = torch.randn(9, 1, *rt2.obs_shape, device=dev, generator=gen)
        a2, b2 = sh2n.trafo_adjoint(y2), sh2p.trafo_adjoint(y2)
        assert float((a2 - b2).norm() / a2.norm()) < 1e-5, float((a2 - b2).norm() / a2.norm())
        # sample sharding: each rank steps its own shard, no collective; shards equal the 1-GPU result
        sde = pkg.DDPM()
        abar = sde.alpha_bar_table(dev)
        s = torch.randn(7, 1, 96, 96, device=dev, generator=gen)
        eps = torch.randn(7, 1, 96, 96, device=dev, generator=gen)
        t = torch.ones(7, device=dev) * 500.
        tp = torch.ones(7, device=dev) * 490.
        atb = rt.trafo_adjoint(y)
        xa, _ = rt.dds_step(x, s, atb, eps, t, tp, abar, gamma=0.05, eta=0.15, n_iter=3)
        blo, bhi = shard_range(7, rank, world)
        xs, _ = rt.dds_step(x[blo:bhi], s[blo:bhi], atb[blo:bhi], eps[blo:bhi], t[blo:bhi], tp[blo:bhi], abar,
                            gamma=0.05, eta=0.15, n_iter=3)
        assert float((xs - xa[blo:bhi]).norm() / xa[blo:bhi].norm()) < 1e-4
        torch.cuda.synchronize()
        open(os.path.join(tmp, 'ok%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_angle_and_sample_sharding_nccl_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
