"""The BASELINE.json configurations other than the headline one, measured inside bench.py so that they land in
the driver-run record (every block is called by ALL ranks of a bench run; rank 0 keeps the result):

* ``stack_block``   config 4 -- one 501^2 x 501-slice stack, 1200 angles, angle-sharded over the ranks
                    (`sharding.AngleShardedRayTrafo`, NCCL all-reduce and the peer-staged reduction), with a
                    parity figure against the un-sharded operator; at one rank it is the un-sharded operator
                    itself, so the per-N lines give the strong-scaling curve of `op(v) = v + gamma A*(A v)`.
* ``sweep_block``   config 3 -- data-consistency step alone, global batch 8 -> 256 split evenly over the ranks
                    (no collective).
* ``adapted_block`` config 5 -- SCD adapted sampling at batch 1 per rank (run_adapted_sampling.py defaults):
                    ms per adapted reverse step with the full-size UNet caller and for this repository's share
                    alone (weight-free score of tests/scorenet.py).

Sizes of config 4: reference src/dataset/walnut_utils.py:39-40 (MAX_NUM_ANGLES = 1200, VOL_SZ = 501).
"""
import copy
import os
import sys
from types import SimpleNamespace as NS

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _max_over_ranks(ms, dev, world):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed(fn, iters, dev, world, warm=2):
    for _ in range(warm):
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
    return _max_over_ranks(e0.elapsed_time(e1) / iters, dev, world)


# --------------------------------------------------------------------------- config 4 ---
def stack_block(dev, world, rank, im=501, angles=1200, slices=501, chunk=256, iters=3, check=2, gamma=0.01):
    import diffusion_models_dev_project_b200 as pkg
    from diffusion_models_dev_project_b200.sharding import AngleShardedRayTrafo
    rt = pkg.B200RayTrafo((im, im), angles)
    n_det = rt.obs_shape[1]
    sh = AngleShardedRayTrafo(rt, chunk=chunk, reduce='nccl')
    lo, hi = sh.angle_range
    gen = torch.Generator(device=dev).manual_seed(0)              # replicated stack: same seed on every rank
    x = torch.rand(slices, 1, im, im, device=dev, generator=gen)

    out = {'workload': 'slice stack %dx%d x %d slices, %d angles x %d bins, angle-sharded over %d GPU(s), '
                       'slice chunks of %d' % (im, im, slices, angles, n_det, world, chunk),
           'n_gpus': world, 'angles_per_rank': hi - lo, 'chunk_slices': chunk}

    # ---- parity of the sharded operators against the un-sharded ones on a slice subset ----
    xs = x[:check].contiguous()
    y_full = rt(xs)                                               # all 1200 angles on this rank
    z_full = rt.trafo_adjoint(y_full)
    n_full = rt.normal_apply(xs, gamma)
    parity = {}
    y_loc = sh(xs)
    parity['A_own_rows_rel_l2'] = float((y_loc[..., lo:hi, :] - y_full[..., lo:hi, :]).norm()
                                        / y_full[..., lo:hi, :].norm())
    parity['Aadj_nccl_rel_l2'] = float((sh.trafo_adjoint(y_full) - z_full).norm() / z_full.norm())
    parity['op_nccl_rel_l2'] = float((sh.normal_apply(xs, gamma) - n_full).norm() / n_full.norm())
    shp = None
    if world > 1:
        try:
            shp = AngleShardedRayTrafo(rt, chunk=chunk, reduce='peer')
            parity['Aadj_peer_rel_l2'] = float((shp.trafo_adjoint(y_full) - z_full).norm() / z_full.norm())
            parity['op_peer_rel_l2'] = float((shp.normal_apply(xs, gamma) - n_full).norm() / n_full.norm())
        except Exception as e:                                    # symmetric memory unavailable on this box
            out['peer_unavailable'] = '%s: %s' % (type(e).__name__, str(e)[:200])
            shp = None
        ok = torch.tensor([1.0 if shp is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            shp = None
    # every rank must hold the same replicated result after the sum
    if world > 1:
        z = sh.trafo_adjoint(y_full)
        zs = [torch.empty_like(z) for _ in range(world)]
        dist.all_gather(zs, z)
        parity['replicas_identical'] = bool(all(torch.equal(zs[0], t) for t in zs))
    out['parity'] = parity
    del y_full, z_full, n_full, y_loc

    # ---- timings (CUDA events, max over ranks) ----
    y = sh(x)                                                     # own rows, others zero
    out['A_ms'] = _timed(lambda: sh(x), iters, dev, world)
    nloc = min(chunk, slices)
    out['Aadj_local_ms'] = _timed(lambda: rt._bp(y[:nloc], rt.adj_scale, angle_range=(lo, hi)), iters, dev, world) \
        * (slices / nloc)
    out['Aadj_ms'] = _timed(lambda: sh.trafo_adjoint(y), iters, dev, world)
    if shp is not None:
        out['Aadj_ms_peer'] = _timed(lambda: shp.trafo_adjoint(y), iters, dev, world)
    del y
    out['allreduce_alone_ms'] = 0.0
    out['allreduce_bytes'] = 0
    if world > 1:
        part = torch.empty(nloc, 1, im, im, device=dev)
        out['allreduce_alone_ms'] = _timed(lambda: dist.all_reduce(part), iters, dev, world) * (slices / nloc)
        out['allreduce_bytes'] = 4 * slices * im * im
        out['allreduce_GBps_algbw'] = out['allreduce_bytes'] / (out['allreduce_alone_ms'] * 1e-3) / 1e9
        del part
    out['op_ms_nccl'] = _timed(lambda: sh.normal_apply(x, gamma), iters, dev, world)
    out['op_local_ms'] = _timed(lambda: rt.normal_apply(x[:nloc], gamma, angle_range=(lo, hi), add_identity=False),
                                iters, dev, world) * (slices / nloc)
    if shp is not None:
        out['op_ms_peer'] = _timed(lambda: shp.normal_apply(x, gamma), iters, dev, world)
        # the owner's reduction storing once to the NVSwitch multicast address instead of once per peer
        try:
            shm = AngleShardedRayTrafo(rt, chunk=chunk, reduce='peer', multicast=True)
            shm.normal_apply(xs, gamma)
            has_mc = 1.0 if shm._peer is not None and shm._peer.use_multicast else 0.0
        except Exception as e:
            out['multicast_unavailable'] = '%s: %s' % (type(e).__name__, str(e)[:200])
            has_mc = 0.0
        ok = torch.tensor([has_mc], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) > 0:
            n_full = rt.normal_apply(xs, gamma)
            out['parity']['op_peer_multicast_rel_l2'] = float((shm.normal_apply(xs, gamma) - n_full).norm() / n_full.norm())
            out['op_ms_peer_multicast'] = _timed(lambda: shm.normal_apply(x, gamma), iters, dev, world)
        del shm
    out['op_ms'] = min(v for v in (out['op_ms_nccl'], out.get('op_ms_peer'), out.get('op_ms_peer_multicast')) if v is not None)
    out['exposed_collective_ms'] = max(0.0, out['op_ms_nccl'] - out['op_local_ms']) if world > 1 else 0.0
    out['overlap_hidden_ms'] = max(0.0, out['allreduce_alone_ms'] - out['exposed_collective_ms'])
    bytes_rank = 4 * (im * im + (hi - lo) * n_det) * slices       # per-rank algorithmic bytes of A (and of A*)
    out['A_GBps_per_rank'] = bytes_rank / (out['A_ms'] * 1e-3) / 1e9
    out['Aadj_GBps_per_rank'] = bytes_rank / (out['Aadj_ms'] * 1e-3) / 1e9
    out['limiter'] = ('single GPU: A (fp_march) %.1f ms + A* (bp_tile) %.1f ms, no collective'
                      % (out['A_ms'], out['Aadj_ms'])) if world == 1 else \
        ('op = %.2f ms of which local A + A* %.2f ms; exposed part of the chunked all-reduce %.2f ms '
         '(all-reduce alone %.2f ms)' % (out['op_ms_nccl'], out['op_local_ms'], out['exposed_collective_ms'],
                                         out['allreduce_alone_ms']))
    del x, sh, shp
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- config 3 ---
def sweep_block(dev, world, rank, im=256, angles=60, k=5, batches=(8, 16, 32, 64, 128, 256), iters=20):
    import diffusion_models_dev_project_b200 as pkg
    rt = pkg.B200RayTrafo((im, im), angles)
    abar = pkg.DDPM().alpha_bar_table(dev)
    rows = []
    for gb in batches:
        if gb % world:
            continue
        B = gb // world
        gen = torch.Generator(device=dev).manual_seed(rank)
        x = torch.rand(B, 1, im, im, device=dev, generator=gen)
        s = torch.randn(B, 1, im, im, device=dev, generator=gen)
        eps = torch.randn(B, 1, im, im, device=dev, generator=gen)
        atb = rt.trafo_adjoint(rt(x))
        t = torch.ones(B, device=dev) * 500.
        tp = torch.ones(B, device=dev) * 490.
        ms = _timed(lambda: rt.dds_step(x, s, atb, eps, t, tp, abar, 0.01, 0.15, k), iters, dev, world, warm=3)
        rows.append({'global_batch': gb, 'batch_per_gpu': B, 'ms_per_step': ms,
                     'hot_path_samples_per_s': gb / (100 * ms * 1e-3),
                     'hot_path_samples_per_s_per_gpu': B / (100 * ms * 1e-3)})
    return {'workload': 'data-consistency step alone (Tweedie + CG(%d) + DDIM), %dx%d, %d angles, global batch split '
                        'evenly over %d GPU(s), no collective; samples/s at 100 reverse steps per sample'
                        % (k, im, im, angles, world), 'rows': rows}


# --------------------------------------------------------------------------- config 5 ---
def adapted_block(dev, world, rank, steps=2, warm=1, full_unet=True):
    import diffusion_models_dev_project_b200 as pkg
    from diffusion_models_dev_project_b200 import _lib
    from bench_support.adm_unet import aapm_unet
    from bench_support.lora import inject_trainable_lora
    from bench_support.phantoms import disk_ellipses
    tests_dir = os.path.join(ROOT, 'tests')
    if tests_dir not in sys.path:
        sys.path.insert(0, tests_dir)
    from scorenet import AdaptableScore

    args = NS(method='dds', num_steps=50, adapt_freq=1, eta=0.85, gamma=0.01, adaptation='lora',
              lora_include_blocks=None, lora_rank=4, tv_penalty=1e-6, num_optim_step=10,
              lr=1e-3, add_cg=True, dc_type='cg', cg_iter=1, early_stopping_pct=1.0)
    config = NS(device=dev, sampling=NS(batch_size=1, eps=1e-3, travel_length=1, travel_repeat=1),
                model=NS(in_channels=1))
    rt = pkg.B200RayTrafo((256, 256), 60)
    sde = pkg.DDPM()
    gt = torch.from_numpy(disk_ellipses(1, 256, seed=1 + rank)).to(dev)
    y = pkg.simulate(gt, rt, 0.01, rng=np.random.default_rng(1 + rank))
    lib = _lib.load()

    def time_steps(sampler):
        kw = sampler.sample_kwargs
        skip = sde.num_steps // kw['num_steps']
        torch.manual_seed(1 + rank)
        x = sde.prior_sampling([kw['batch_size'], *kw['im_shape']]).to(dev)
        ones = torch.ones(kw['batch_size'], device=dev)
        pred_kw = dict(kw['predictor'], use_adapt=True)

        def step(i, x):
            t = (kw['num_steps'] - 1 - i) * skip
            o, _ = sampler.predictor(score=sampler.score, sde=sde, x=x,
                                     time_step=(ones * t, ones * max(t - skip, -1)), step_size=1,
                                     datafitscale=1., **pred_kw)
            return o
        for i in range(warm):
            x = step(i, x)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.scd_launch_count_reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            x = step(warm + i, x)
        b.record()
        torch.cuda.synchronize()
        finite = bool(torch.isfinite(x).all())
        return _max_over_ranks(a.elapsed_time(b) / steps, dev, world), int(lib.scd_launch_count()) // steps, finite

    import contextlib
    with torch.enable_grad(), contextlib.redirect_stdout(sys.stderr):        # the factory prints; stdout is the JSON line's
        a2 = copy.copy(args)
        a2.adaptation = 'full'
        path = pkg.get_standard_adapted_sampler(args=a2, config=config, score=AdaptableScore().to(dev), sde=sde,
                                                ray_trafo=rt, observation=y, device=dev)
        ms_path, launches, ok1 = time_steps(path)
        # the same with one Adam step (score model, scd_adapt_fwd / scd_adapt_bwd, optimizer) captured in a CUDA graph
        a3 = copy.copy(a2)
        a3.adapt_cuda_graph = True
        pathg = pkg.get_standard_adapted_sampler(args=a3, config=config, score=AdaptableScore().to(dev), sde=sde,
                                                 ray_trafo=rt, observation=y, device=dev)
        ms_graph, _, ok3 = time_steps(pathg)
        ok1 = ok1 and ok3
        out = {'workload': 'SCD adapted sampling 256x256, 60 angles, batch 1 per GPU, 50 reverse steps per sample, '
                           '10 Adam steps per reverse step, LoRA rank 4, CG(1), tv 1e-6, eta 0.85',
               'ms_per_reverse_step_path_only': ms_path, 'ms_per_reverse_step_path_only_cuda_graph': ms_graph,
               'library_launches_per_reverse_step': launches,
               'finite': ok1}
        if full_unet:
            torch.manual_seed(0)
            score = aapm_unet().to(dev).eval()
            full = pkg.get_standard_adapted_sampler(args=args, config=config, score=score, sde=sde, ray_trafo=rt,
                                                    observation=y, device=dev, lora_inject_fn=inject_trainable_lora)
            ms_full, _, ok2 = time_steps(full)
            out.update({'ms_per_reverse_step_full_unet': ms_full,
                        'samples_per_s_full_unet': world * 1e3 / (ms_full * args.num_steps),
                        'path_share_of_step': ms_path / ms_full, 'finite': ok1 and ok2})
            del score, full
            # the same with the whole Adam step -- UNet forward + backward included -- captured in a CUDA graph: possible
            # because scd_adapt_fwd / scd_adapt_bwd allocate nothing and never synchronise
            try:
                torch.manual_seed(0)
                score = aapm_unet().to(dev).eval()
                a4 = copy.copy(args)
                a4.adapt_cuda_graph = True
                fullg = pkg.get_standard_adapted_sampler(args=a4, config=config, score=score, sde=sde, ray_trafo=rt,
                                                         observation=y, device=dev, lora_inject_fn=inject_trainable_lora)
                ms_fullg, _, ok4 = time_steps(fullg)
                out.update({'ms_per_reverse_step_full_unet_cuda_graph': ms_fullg,
                            'samples_per_s_full_unet_cuda_graph': world * 1e3 / (ms_fullg * args.num_steps),
                            'finite': out['finite'] and ok4})
                del score, fullg
            except Exception as e:                      # capture of a caller's model is best effort
                out['full_unet_cuda_graph_unavailable'] = '%s: %s' % (type(e).__name__, str(e)[:200])
    torch.cuda.empty_cache()
    return out
