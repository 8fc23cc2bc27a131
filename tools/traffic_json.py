"""rNN_traffic.json from an ncu `--set full` capture of the hot kernels inside the bench command: per kernel the
mean DRAM bytes per launch, duration, L1 / L2 hit rates and the share of the shared-memory pipe
(l1tex__data_pipe_lsu_wavefronts_mem_shared over elapsed cycles x 148 SMs).  bench.py reads the newest file for
`roofline.traffic`.

    python tools/traffic_json.py gpurun_out/r2/prof_b8.ncu-rep > profiles/r02_traffic.json
"""
import collections
import csv
import json
import subprocess
import sys


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(head)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tscale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
    acc = collections.defaultdict(list)
    for r in data:
        name = r[col['Kernel Name']].replace('void ', '').split('<')[0].split('(')[0]
        dram = sum(float(r[col[k]]) * scale.get(units[col[k]], 1.0) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        us = float(r[col['gpu__time_duration.sum']]) * tscale.get(units[col['gpu__time_duration.sum']], 1.0)
        wf = float(r[col['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']])
        pipe = wf / (float(r[col['sm__cycles_elapsed.max']]) * 148.0)
        acc[name].append((dram, pipe, us, float(r[col['l1tex__t_sector_hit_rate.pct']]), float(r[col['lts__t_sector_hit_rate.pct']])))
    res = {}
    for name, v in acc.items():
        n = len(v)
        res[name] = {'dram_bytes_per_launch': sum(x[0] for x in v) / n, 'shared_pipe_frac': sum(x[1] for x in v) / n,
                     'us_under_ncu': sum(x[2] for x in v) / n, 'l1_hit_pct': sum(x[3] for x in v) / n,
                     'l2_hit_pct': sum(x[4] for x in v) / n, 'launches_sampled': n}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
