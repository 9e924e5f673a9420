"""Summarise an .ncu-rep (read with `ncu -i`) into the few numbers the roofline needs."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index('Kernel Name')
    for r in data:
        print('kernel:', r[name_i][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:66s} {r[i]:>16s} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
