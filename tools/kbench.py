"""Kernel micro-benchmark / tuning sweep for fp_march and bp_tile (CUDA events, L2 flushed).

    python tools/kbench.py --kernel fp --batch 8 --sweep
    python tools/kbench.py --kernel fp --batch 256 --set fp_samples=4,fp_angles=2 --iters 5   (for ncu)
"""
import argparse
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_models_dev_project_b200 as pkg  # noqa: E402


def timed(fn, iters, flush):
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in evs]
    return float(np.median(t)), float(np.min(t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--kernel', default='fp', choices=['fp', 'march', 'bp', 'cg'])
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--im', type=int, default=256)
    ap.add_argument('--angles', type=int, default=60)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--set', default='')
    ap.add_argument('--sweep', action='store_true')
    ap.add_argument('--no-flush', action='store_true')
    a = ap.parse_args()
    dev = torch.device('cuda')
    rt = pkg.B200RayTrafo((a.im, a.im), a.angles)
    x = torch.rand(a.batch, 1, a.im, a.im, device=dev)
    y = rt(x)
    p = torch.rand_like(x)
    flush = None if a.no_flush else torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    nbytes = 4 * (a.im * a.im + rt.obs_shape[0] * rt.obs_shape[1]) * a.batch

    def run(tune):
        keys = ['fp_samples', 'fp_angles', 'fp_rows', 'fp_threads', 'fp_nbuf', 'fp_cluster', 'fp_plan', 'fp_plan_cost', 'fp_source', 'fp_cls0', 'bp_tile', 'bp_share', 'bp_rows']
        rt.set_tuning(dev, **{k: tune.get(k, 0) for k in keys})
        if a.kernel == 'fp':
            fn = lambda: rt._fp(x)          # noqa: E731
        elif a.kernel == 'march':           # the projector alone on an interleaved image (as inside CG)
            x_il = rt._img_il(x)
            fn = lambda: rt._fp_ilimg(x_il, a.batch)   # noqa: E731
        elif a.kernel == 'bp':
            fn = lambda: rt._bp(y, 0.01, addend=p, addend_scale=1.0)   # noqa: E731
        else:
            op = rt.normal_op(0.01)
            fn = lambda: pkg.cg(op, x, p, 5)   # noqa: E731
        for _ in range(3):
            fn()
        med, mn = timed(fn, a.iters, flush)
        print('%-60s med %8.1f us  min %8.1f us  %7.1f GB/s  %6.3f us/sample' %
              (tune, med * 1e3, mn * 1e3, nbytes / (med * 1e-3) / 1e9, med * 1e3 / a.batch), flush=True)

    if a.sweep:
        if a.kernel == 'fp':
            run(dict())
            sbs = [sb for sb in (4, 8, 16) if sb <= max(4, a.batch)]
            big = a.batch >= 64
            for SB, NA, TR, CS, TH in itertools.product(sbs, [2, 4], [2, 4, 8], [1, 2, 4], [512, 768, 1024]):
                if (SB >= 16 and TR == 8) or (SB <= 8 and TR == 2) or (SB == 4 and TH == 1024):
                    continue
                if big and (CS > 1 or SB == 4):
                    continue
                if CS > 1 and a.batch // SB * (a.angles // NA) * CS > 1200:
                    continue
                run(dict(fp_samples=SB, fp_angles=NA, fp_rows=TR, fp_cluster=CS, fp_threads=TH))
        elif a.kernel == 'bp':
            run(dict())
            for SB, T in itertools.product([4, 8, 16], [8, 16, 32]):
                if SB > max(4, a.batch):
                    continue
                run(dict(fp_samples=SB, bp_tile=T))
        else:
            run(dict())
            for TH, PL in itertools.product([512, 768, 1024], [0, 1]):
                run(dict(fp_threads=TH, fp_plan=PL))
    else:
        tune = {}
        for kv in a.set.split(','):
            if kv:
                k, v = kv.split('=')
                tune[k] = int(v)
        run(tune)


if __name__ == '__main__':
    main()
