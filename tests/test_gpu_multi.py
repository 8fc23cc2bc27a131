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
        from diffusion_models_dev_project_b200.sharding import angle_cost_ranges
        sh = AngleShardedRayTrafo(rt, chunk=3)
        lo, hi = sh.angle_range
        assert (lo, hi) == angle_cost_ranges(rt.angles, world)[rank]
        assert AngleShardedRayTrafo(rt, chunk=3, balance='count').angle_range == shard_range(30, rank, world)
        gen = torch.Generator(device=dev).manual_seed(0)          # replicated vectors
        x = torch.rand(7, 1, 96, 96, device=dev, generator=gen)
        y = torch.randn(7, 1, *rt.obs_shape, device=dev, generator=gen)
        full = rt(x)
        yl = sh(x)
        assert torch.equal(yl[..., lo:hi, :], full[..., lo:hi, :])
        assert float(yl[..., :lo, :].abs().sum() + yl[..., hi:, :].abs().sum()) == 0.0
        assert float((sh.gather_sinogram(yl) - full).norm() / full.norm()) < 1e-6
        # every sharded result is checked against the C oracle (kernel-vs-oracle is ~1e-7; gate 1e-4)
        from oracle import oracle as O
        geom = O.OracleGeometry((96, 96), 30)
        z_or = torch.from_numpy(O.bp(geom, y.cpu().numpy())).to(dev)
        zf = rt.trafo_adjoint(y)
        assert float((zf - z_or).norm() / z_or.norm()) < 1e-5
        z = sh.trafo_adjoint(y)
        assert float((z - z_or).norm() / z_or.norm()) < 1e-5
        gamma = 0.05
        n_or = x + gamma * torch.from_numpy(O.bp(geom, O.fp(geom, x.cpu().numpy()))).to(dev)
        assert float((sh.normal_apply(x, gamma) - n_or).norm() / n_or.norm()) < 1e-5
        sol = pkg.cg(op=sh.normal_op(gamma), x=x, rhs=x + 1.0, n_iter=3)
        ref = pkg.cg(op=rt.normal_op(gamma), x=x, rhs=x + 1.0, n_iter=3)
        assert float((sol - ref).norm() / ref.norm()) < 1e-4
        gathered = [torch.empty_like(sol) for _ in range(world)]
        dist.all_gather(gathered, sol)
        assert all(torch.equal(gathered[0], t) for t in gathered)     # replicas stay identical
        with torch.enable_grad():                                     # inference-only view: loud, not silent
            try:
                sh.normal_apply(x.clone().requires_grad_(True), gamma)
                raise AssertionError('autograd through the sharded view must be refused')
            except RuntimeError as e:
                assert 'does not support autograd' in str(e)
        # the same with the sum riding on the backprojector (peer-staged bands instead of the all-reduce)
        shp = AngleShardedRayTrafo(rt, chunk=3, reduce='peer')
        zp = shp.trafo_adjoint(y)
        assert float((zp - z_or).norm() / z_or.norm()) < 1e-5
        zp2 = shp.trafo_adjoint(y)
        assert torch.equal(zp, zp2)                                # deterministic
        na = shp.normal_apply(x, gamma)
        assert float((na - n_or).norm() / n_or.norm()) < 1e-5
        # TIGHT check of the band reduction against the all-reduce: force the same tiling in both launches (whole
        # tiles, bp_rows = 1; same chunking) -- then the per-rank partials are bit-identical and, with two ranks,
        # so is their sum (a + b is commutative; the peer path adds the staged copies in rank order)
        rt.set_tuning(dev, bp_rows=1)
        z_t, zp_t = sh.trafo_adjoint(y), shp.trafo_adjoint(y)
        if world == 2:
            assert torch.equal(z_t, zp_t), float((z_t - zp_t).abs().max())
        assert float((z_t - zp_t).norm() / z_t.norm()) < 1e-6
        # op: the all-reduce path carries the identity term on rank 0's partial, the peer path adds it in the
        # owner's reduction -- same terms, different association
        nb_t, na_t = sh.normal_apply(x, gamma), shp.normal_apply(x, gamma).clone()
        assert float((na_t - nb_t).norm() / nb_t.norm()) < 1e-6, float((na_t - nb_t).norm() / nb_t.norm())
        rt.set_tuning(dev, bp_rows=0)
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
        o2 = torch.from_numpy(O.bp(O.OracleGeometry((70, 50), 9), y2.cpu().numpy())).to(dev)
        assert float((a2 - o2).norm() / o2.norm()) < 1e-5 and float((b2 - o2).norm() / o2.norm()) < 1e-5
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
