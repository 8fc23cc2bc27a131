"""Print the metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    name_i = head.index('Kernel Name')
    for r in data:
        print('# kernel:', r[name_i][:80])
        for w in WANT:
            if w in head:
                i = head.index(w)
                print('%-84s %s %s' % (w, r[i], units[i]))




def table(path):
    """One line per launch: the figures DESIGN.md quotes (duration, HBM GB/s, L1/L2 hit rates, pipes)."""
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, data = rows[0], rows[2:]
    col = {k: head.index(k) for k in head}

    def g(r, k, default='nan'):
        return r[col[k]] if k in col else default
    print('%-34s %9s %7s %5s %8s %8s %7s %7s %7s %7s %7s %7s' % (
        'kernel', 'us', 'grid', 'regs', 'dram MB', 'HBM GB/s', 'L1hit%', 'L2hit%', 'smem%', 'issue%', 'lts%', 'active%'))
    for r in data:
        us = float(g(r, 'gpu__time_duration.sum'))
        scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
        mb = sum(float(g(r, k)) * scale.get(rows[1][col[k]], 1.0) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        wf = float(g(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'))
        smem = 100.0 * wf / (float(g(r, 'sm__cycles_elapsed.max')) * 148.0)
        name = g(r, 'Kernel Name').replace('void ', '')[:34]
        act = 100.0 * float(g(r, 'sm__cycles_active.avg')) / float(g(r, 'sm__cycles_elapsed.max'))
        print('%-34s %9.1f %7s %5s %8.2f %8.1f %7.1f %7.1f %7.1f %7.1f %7.1f %7.1f' % (
            name, us, g(r, 'launch__grid_size'), g(r, 'launch__registers_per_thread'), mb, mb / us * 1e3,
            float(g(r, 'l1tex__t_sector_hit_rate.pct')), float(g(r, 'lts__t_sector_hit_rate.pct')),
            smem,
            float(g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')),
            float(g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')), act))


if __name__ == '__main__' and len(sys.argv) > 2 and sys.argv[2] == '--table':
    table(sys.argv[1])
elif __name__ == '__main__':
    main(sys.argv[1])
