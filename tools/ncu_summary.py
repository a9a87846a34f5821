"""Summarise an .ncu-rep (run here, no GPU needed): python tools/ncu_summary.py <rep> [out.md]
Keeps the metrics DESIGN.md / bench.py cite: duration, DRAM bytes, L2 sectors, hit rates, occupancy,
tensor-pipe activity, registers, top stall reasons."""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex.sum',
    'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index('Kernel Name')
    out.write('# ncu summary of %s\n\n' % rep.split('/')[-1])
    for n, r in enumerate(rows[2:]):
        out.write('## launch %d: `%s`\n\n| metric | value | unit |\n|---|---|---|\n' % (n, r[name_i]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write('| %s | %s | %s |\n' % (k, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                try:
                    stalls.append((float(r[i].replace(',', '')), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:6]:
            out.write('| %s | %.2f | %% |\n' % (h, v))
        out.write('\n')


if __name__ == '__main__':
    main()
