"""Print the key metrics (and optionally the hottest SASS lines) of an .ncu-rep file."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f'  {h:75s} {units[i]:10s} {vals[i]}')
        elif 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v > 0.3:
                print(f'  {h:75s} {vals[i]}')


def source(path, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    isrc, ismp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ismp]), r))
        except Exception:
            pass
    tot = sum(s for s, _ in data) or 1
    print(f'  total samples {tot}')
    for s, r in sorted(data, key=lambda x: -x[0])[:top]:
        st = sorted([(int(r[i] or 0), hdr[i]) for i in stall], reverse=True)[:2]
        print(f'  {100 * s / tot:5.1f}% ex={r[iex]:>9s} {r[isrc][:60]:60s} {st}')


if __name__ == '__main__':
    for p in sys.argv[1:]:
        if p.startswith('--'):
            continue
        print(p)
        raw(p)
        if '--src' in sys.argv:
            source(p)
