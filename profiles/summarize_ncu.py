"""Turn an `ncu --set full` report into the small per-kernel CSV committed under profiles/ (runs on the CPU container: ncu -i).
    python profiles/summarize_ncu.py gpurun_out/r01c_k_p2g.ncu-rep profiles/r01c_ncu_k_p2g.csv"""
import csv
import subprocess
import sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum')


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index('Kernel Name')]
    with open(out, 'w') as f:
        f.write(f'kernel,{kname.replace(",", ";")}\n')
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
                f.write(f'{h},{v.replace(",", "")},{u}\n')
    print(open(out).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
