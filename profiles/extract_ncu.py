"""Turn an `ncu --set full` report into a compact per-launch table (the judged summary under profiles/).

    python profiles/extract_ncu.py gpurun_out/prof_gemm_r1.ncu-rep > profiles/r1_ncu_gemm_persist.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
    ('sm__inst_executed_pipe_tc.sum', 'tc_inst'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('lts__t_sector_hit_rate.pct', 'l2hit%'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_longsb'),
]


def to_base(v, unit):
    v = float(v.replace(',', ''))
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit)
    return v * scale if scale else v


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    print('# %s' % path)
    print('# time in us, dram bytes in MB (read / written by the launch), %% of peak as reported by ncu')
    print('%-4s %-44s %9s %9s %9s %6s %6s %8s %7s %6s %5s %6s %6s %7s' % (
        'id', 'kernel', 'time_us', 'dram_rd', 'dram_wr', 'dram%', 'sm%', 'tensor%', 'issue%', 'occ%', 'regs', 'grid', 'l2hit', 'longsb'))
    for n, r in enumerate(rows[2:]):
        vals = {}
        for m, short in METRICS:
            if m in hdr:
                i = hdr.index(m)
                try:
                    vals[short] = to_base(r[i], units[i])
                except ValueError:
                    vals[short] = float('nan')
            else:
                vals[short] = float('nan')
        name = r[ki].split('(')[0].replace('void ', '')[:44]
        print('%-4d %-44s %9.1f %9.1f %9.1f %6.1f %6.1f %8.1f %7.1f %6.1f %5.0f %6.0f %6.1f %7.2f' % (
            n, name, vals['time'], vals['dram_rd'] / 1e6, vals['dram_wr'] / 1e6, vals['dram%'], vals['sm%'], vals['tensor%'],
            vals['issue%'], vals['occ%'], vals['regs'], vals['grid'], vals['l2hit%'], vals['stall_longsb']))


if __name__ == '__main__':
    main(sys.argv[1])
