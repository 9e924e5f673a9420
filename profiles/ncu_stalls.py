"""Print duration, DRAM traffic, occupancy and the top warp-stall reasons of every kernel in an .ncu-rep."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print('kernel:', r[hdr.index('Kernel Name')][:80])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:64s} {r[i]:>18s} {units[i]}')
        st = []
        for i, k in enumerate(hdr):
            if 'smsp__average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
                st.append((float(r[i]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        print('  stalls (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)[:7]))


if __name__ == '__main__':
    main(sys.argv[1])
