"""World-size-2 tests of the multi-GPU partitioning logic on CPU (gloo backend)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_l2
from oracle import oracle as O
from diffusion_models_dev_project_b200.sharding import AngleShardedRayTrafo, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 8, 60, 1200, 501):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


class OracleBase:
    """CPU stand-in with the `_fp/_bp(angle_range=)` surface of B200RayTrafo, backed by the oracle."""

    def __init__(self, geom):
        self.geom = geom
        self.im_shape, self.obs_shape = geom.im_shape, geom.obs_shape
        self.adj_scale = geom.adj_scale
        self.angles = geom.angles

    def _fp(self, x, angle_range=None):
        y = torch.from_numpy(O.fp(self.geom, x.numpy()))
        if angle_range is not None:
            lo, hi = angle_range
            mask = torch.zeros(self.obs_shape[0], 1)
            mask[lo:hi] = 1
            y = y * mask
        return y

    def _bp(self, y, scale, angle_range=None):
        out = O.bp(self.geom, y.numpy(), angle_range=angle_range)
        return torch.from_numpy(out) * (scale / self.geom.adj_scale)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import diffusion_models_dev_project_b200 as pkg
        geom = O.OracleGeometry((24, 24), 10)
        base = OracleBase(geom)
        sh = AngleShardedRayTrafo(base, chunk=2)
        from diffusion_models_dev_project_b200.sharding import angle_cost_ranges
        assert sh.angle_range == angle_cost_ranges(geom.angles, world)[rank]
        g = torch.Generator().manual_seed(0)            # same data on every rank (replicated vectors)
        x = torch.rand(5, 1, 24, 24, generator=g)
        y = torch.randn(5, 1, *geom.obs_shape, generator=g)
        # A stays sharded: own rows only; gathered = full operator
        yl = sh(x)
        lo, hi = sh.angle_range
        full = base._fp(x)
        assert torch.equal(yl[..., lo:hi, :], full[..., lo:hi, :])
        assert float(yl[..., :lo, :].abs().sum() + yl[..., hi:, :].abs().sum()) == 0.0
        assert rel_l2(sh.gather_sinogram(yl).numpy(), full.numpy()) < 1e-6
        # A*: partial images summed by all-reduce == single-process result
        assert rel_l2(sh.trafo_adjoint(y).numpy(), base._bp(y, geom.adj_scale).numpy()) < 1e-6
        # normal operator and CG on replicated vectors: identical on all ranks, equal to 1 process
        gamma = 0.05
        ref_op = lambda v: v + gamma * base._bp(base._fp(v), geom.adj_scale)      # noqa: E731
        assert rel_l2(sh.normal_apply(x, gamma).numpy(), ref_op(x).numpy()) < 1e-6
        sol = pkg.cg(op=sh.normal_op(gamma), x=x, rhs=x + 1.0, n_iter=3)
        ref = pkg.cg(op=ref_op, x=x, rhs=x + 1.0, n_iter=3)
        assert rel_l2(sol.numpy(), ref.numpy()) < 1e-5
        gathered = [torch.empty_like(sol) for _ in range(world)]
        dist.all_gather(gathered, sol)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        # sample sharding: disjoint contiguous shards, no collective on the data path
        blo, bhi = shard_range(5, rank, world)
        mine = ref_op(x[blo:bhi])
        assert torch.equal(mine, ref_op(x)[blo:bhi])
        open(os.path.join(tmp, 'ok%d' % rank), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_angle_and_sample_sharding_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']


def test_band_layout_covers_the_image_with_tile_aligned_bands():
    """Row bands of the peer-staged reduction: one band per rank, a multiple of 32 rows (the tallest
    backprojector tile), together covering the image; trailing bands may be empty."""
    from diffusion_models_dev_project_b200.sharding import band_layout
    for n_rows in (16, 33, 64, 256, 501, 1024):
        for world in (1, 2, 3, 4, 8):
            rows = band_layout(n_rows, world)
            assert rows % 32 == 0 and rows * world >= n_rows
            assert rows - 32 < -(-n_rows // world) <= rows          # no more padding than one tile row
            covered = sum(max(0, min(n_rows, (r + 1) * rows) - min(n_rows, r * rows)) for r in range(world))
            assert covered == n_rows


def test_sharded_view_refuses_autograd():
    """The sharded operators bypass the autograd Functions of the wrapped operator: a gradient through A*A would
    be dropped silently, so inputs that require grad are rejected (single process, world size 1)."""
    geom = O.OracleGeometry((16, 16), 6)
    sh = AngleShardedRayTrafo(OracleBase(geom), chunk=2)
    x = torch.rand(3, 1, 16, 16, requires_grad=True)
    for call in (lambda: sh(x), lambda: sh.normal_apply(x, 0.1),
                 lambda: sh.trafo_adjoint(torch.rand(3, 1, *geom.obs_shape, requires_grad=True))):
        with pytest.raises(RuntimeError, match='does not support autograd'):
            call()
    with torch.no_grad():
        out = sh.normal_apply(x, 0.1)
    ref = x.detach() + 0.1 * OracleBase(geom)._bp(OracleBase(geom)._fp(x.detach()), geom.adj_scale)
    assert rel_l2(out.numpy(), ref.numpy()) < 1e-6


def test_angle_cost_ranges_partition_and_balance():
    """Cost-balanced contiguous angle ranges: a partition of all angles, every rank non-empty, per-rank cost within
    one angle's cost of the mean; the dear angles (near the axes) end up in shorter ranges."""
    from diffusion_models_dev_project_b200.sharding import angle_cost_ranges
    for n in (8, 30, 60, 1200):
        ang = (np.arange(n) + 0.5) * np.pi / n
        cost = 1.75 * np.maximum(np.abs(np.cos(ang)), np.abs(np.sin(ang))) / 0.9 + 1.0
        for world in (1, 2, 3, 4, 8):
            parts = angle_cost_ranges(ang, world)
            assert parts[0][0] == 0 and parts[-1][1] == n and len(parts) == world
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert all(hi > lo for lo, hi in parts)
            per = [cost[lo:hi].sum() for lo, hi in parts]
            assert max(per) - cost.sum() / world <= cost.max() + 1e-9
    parts = angle_cost_ranges((np.arange(1200) + 0.5) * np.pi / 1200, 8)
    sizes = [hi - lo for lo, hi in parts]
    assert sizes[0] < sizes[1] and sizes[3] < sizes[2] and sizes == sizes[::-1]        # 142 / 158 instead of 150 / 150
    assert angle_cost_ranges(np.zeros(3), 8)[7] == shard_range(3, 7, 8)                  # fewer angles than ranks


def test_chunk_bounds_are_levelled_and_cover_the_stack():
    geom = O.OracleGeometry((16, 16), 6)
    for n, chunk, want in ((501, 256, [251, 250]), (501, 128, [126, 125, 125, 125]), (7, 3, [3, 2, 2]), (5, 8, [5]), (0, 4, [])):
        sh = AngleShardedRayTrafo(OracleBase(geom), chunk=chunk)
        b = sh._chunk_bounds(n)
        assert [hi - lo for lo, hi in b] == want
        assert all(hi - lo <= chunk for lo, hi in b)
        assert (b[0][0], b[-1][1]) == (0, n) if n else b == []
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
