"""Config 4 of BASELINE.json: one large slice stack, angle-sharded over the GPUs of a box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/stack_bench.py [--im 501] [--angles 1200] [--slices 501] [--chunk 64]

Every rank holds the whole (replicated) stack `[S, 1, n, n]` and owns the angles
`[r*N_theta/P, (r+1)*N_theta/P)`.  A stays sharded (no communication); A* backprojects the own
angles into partial images that are summed with an NCCL all-reduce in slice chunks on a side
stream, chunk c's reduction overlapping chunk c+1's backprojection
(`sharding.AngleShardedRayTrafo`).  Prints one JSON line (rank 0): device times (CUDA events,
max over ranks) of A, the local part of A*, A* including the all-reduce, the all-reduce alone,
and the algorithmic GB/s of SURVEY.md section 8(d) for the per-rank share.
With N = 1 it measures the un-sharded operator.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_models_dev_project_b200 as pkg  # noqa: E402
from diffusion_models_dev_project_b200.sharding import AngleShardedRayTrafo  # noqa: E402


def timed(fn, iters, dev, world):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--im', type=int, default=501)
    ap.add_argument('--angles', type=int, default=1200)
    ap.add_argument('--slices', type=int, default=501)
    ap.add_argument('--chunk', type=int, default=64)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--check', type=int, default=2, help='slices compared against the un-sharded operator')
    ap.add_argument('--reduce', default='nccl', choices=['nccl', 'peer'],
                    help="how the A* partials are summed: chunked NCCL all-reduce, or peer-staged bands written by the "
                         "backprojector's epilogue (sharding.py)")
    a = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.set_grad_enabled(False)

    rt = pkg.B200RayTrafo((a.im, a.im), a.angles)
    sh = AngleShardedRayTrafo(rt, chunk=a.chunk, reduce=a.reduce)
    lo, hi = sh.angle_range
    gen = torch.Generator(device=dev).manual_seed(0)          # replicated stack: same seed on every rank
    x = torch.rand(a.slices, 1, a.im, a.im, device=dev, generator=gen)
    n_det = rt.obs_shape[1]

    # parity of the sharded operators against the un-sharded ones on a few slices
    errs = {}
    if a.check:
        xs = x[:a.check]
        y_full = rt(xs)
        y_loc = sh(xs)
        # not bit-equal in general: the row split (cluster size) is chosen per launch, and it fixes the
        # order in which a line integral's partial sums are added
        errs['A_own_rows_rel_l2'] = float((y_loc[..., lo:hi, :] - y_full[..., lo:hi, :]).norm()
                                          / y_full[..., lo:hi, :].norm())
        z_full = rt.trafo_adjoint(y_full)
        z_sh = sh.trafo_adjoint(y_full)
        errs['Aadj_rel_l2'] = float((z_sh - z_full).norm() / z_full.norm())

    y = sh(x)                                                  # own rows, others zero
    t_fp = timed(lambda: sh(x), a.iters, dev, world)
    t_bp_local = timed(lambda: rt._bp(y[:a.chunk], rt.adj_scale, angle_range=(lo, hi)), a.iters, dev, world) \
        * (a.slices / min(a.chunk, a.slices))
    t_bp = timed(lambda: sh.trafo_adjoint(y), a.iters, dev, world)
    t_normal = timed(lambda: sh.normal_apply(x, 0.01), a.iters, dev, world)
    part = torch.empty(min(a.chunk, a.slices), 1, a.im, a.im, device=dev)
    t_ar = 0.0
    if world > 1:
        t_ar = timed(lambda: dist.all_reduce(part), a.iters, dev, world) * (a.slices / part.shape[0])

    if rank == 0:
        nang = hi - lo
        bytes_rank = 4 * (a.im * a.im + nang * n_det) * a.slices      # per-rank algorithmic bytes of A (and of A*)
        line = {
            'workload': 'slice stack %dx%d x %d slices, %d angles x %d bins, angle-sharded over %d GPU(s)'
                        % (a.im, a.im, a.slices, a.angles, n_det, world),
            'n_gpus': world, 'angles_per_rank': nang, 'chunk_slices': a.chunk,
            'reduce': a.reduce, 'normal_apply_ms': t_normal,
            'A_ms': t_fp, 'Aadj_local_ms': t_bp_local, 'Aadj_with_allreduce_ms': t_bp, 'allreduce_alone_ms': t_ar,
            'overlap_hidden_ms': max(0.0, t_bp_local + t_ar - t_bp),
            'allreduce_bytes': 4 * a.slices * a.im * a.im if world > 1 else 0,
            'A_GBps_per_rank': bytes_rank / (t_fp * 1e-3) / 1e9,
            'Aadj_GBps_per_rank': bytes_rank / (t_bp * 1e-3) / 1e9,
            'parity': errs,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
