"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.
    python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/launches_X.txt
    python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep > profiles/prof_X.txt
"""
import collections
import csv
import os
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_wait']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'])
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
        name = re.sub(r'\(.*', '', row['Kernel Name'])[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    print(f'# ncu --metrics gpu__time_duration.sum --clock-control none : one bench step ({path})')
    print(f'# cold-cache, serialised launch times: compare SHARES, not absolutes')
    print(f'total {tot:.1f} us over {n} launches')
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'{t:10.1f} us {100 * t / tot:5.1f}%  n={c:4d}  {k}')


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f'# ncu --set full --clock-control none --import-source on ({path})')
    for r in rows[2:]:
        print('-' * 100)
        print(r[hdr.index('Kernel Name')][:160])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:75s} {r[i]:>18s} {units[i]}')


def traffic(path, out_json):
    """average DRAM bytes (read + write) per launch of the captured kernel -> small JSON bench.py reads"""
    import json
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tot, n = 0.0, 0
    for r in rows[2:]:
        tot += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        n += 1
    json.dump({'kernel_capture': os.path.basename(path), 'launches': n, 'dram_bytes_per_launch': tot / max(n, 1)},
              open(out_json, 'w'))
    print(open(out_json).read())


if __name__ == '__main__':
    if sys.argv[1] == 'traffic':
        traffic(sys.argv[2], sys.argv[3])
    else:
        {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
