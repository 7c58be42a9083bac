"""Print the per-launch metrics bench.py's roofline line relies on from an ncu report:
    python profiles/ncu_extract.py gpurun_out/<name>.ncu-rep
"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    for r in rows[2:]:
        print('--', r[hdr.index('Kernel Name')][:60])
        for i in idx[1:]:
            print('   %-62s %s %s' % (hdr[i], r[i], units[i]))


if __name__ == '__main__':
    main(sys.argv[1])
