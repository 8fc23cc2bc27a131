"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of `bench.py`:
one reverse step = the launches between two consecutive ddim launches (ddim_il_kernel / ddim_kernel).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_bench_launches_summary.txt
"""
import collections
import csv
import sys

OURS = ('fp_packq_kernel', 'fp_march_kernel', 'bp_tile_kernel', 'sino_pack_kernel', 'cg_update_xr_kernel',
        'cg_update_xr_il_kernel', 'tweedie_rhs_kernel', 'tweedie_il_kernel', 'ddim_kernel', 'ddim_il_kernel',
        'il_pack_kernel', 'il_unpack_kernel', 'residual_sq_kernel', 'tv_fwd_kernel', 'tv_grad_kernel', 'ramp_filter_kernel',
        'adapt_', 'band_reduce')


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, data = rows[0], rows[1:]
    ni, vi = head.index('Kernel Name'), head.index('Metric Value')
    launches = [(r[ni], float(r[vi]) / 1e3) for r in data]           # us
    ends = [i for i, (n, _) in enumerate(launches) if 'ddim_il_kernel' in n or 'ddim_kernel' in n]
    if len(ends) < 2:
        raise SystemExit('need two complete steps in the capture')
    lo, hi = ends[-2] + 1, ends[-1] + 1
    step = launches[lo:hi]
    total = sum(t for _, t in step)
    mine = [(n, t) for n, t in step if any(k in n for k in OURS)]
    print('# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --kernel-batch 0 --no-extra` (B200)')
    print('# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv ...')
    print('# one reverse step = launches %d..%d of the capture: %d launches, sum of kernel time %.1f us' % (lo, hi - 1, len(step), total))
    print('# (cold-cache, serialised per-launch times: compare SHARES, not absolutes)')
    print('# kernels of libscd_b200.so in the step: %d launches, %.1f us = %.2f %% of the step' % (
        len(mine), sum(t for _, t in mine), 100 * sum(t for _, t in mine) / total))
    print('\n## kernels of libscd_b200.so, in launch order')
    for n, t in mine:
        print('%-70s %9.1f us' % (n.replace('void ', '')[:70], t))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in step:
        k = n.replace('void ', '')[:90]
        agg[k][0] += 1
        agg[k][1] += t
    print('\n## aggregated by kernel')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-92s n=%4d %10.1f us %6.2f%%' % (k, c, t, 100 * t / total))


if __name__ == '__main__':
    main(sys.argv[1])
