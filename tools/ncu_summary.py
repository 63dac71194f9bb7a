"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv profiles/x_launches.txt
    python tools/ncu_summary.py full gpurun_out/x.ncu-rep profiles/x_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv, mu, mn = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('Metric Name')
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':      # the CSV may carry the DRAM byte counters as well
            continue
        v = float(r[mv].replace(',', ''))
        v = v / 1e3 if r[mu] == 'ns' else v * 1e3 if r[mu] == 'ms' else v
        a = agg.setdefault(r[kn].split('(')[0], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
        f.write('# source: {}\n# total {:.1f} us over {} launches\n'.format(src, tot, sum(a[0] for a in agg.values())))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('{:10.1f} us {:5.1f}%  n={:4d}  avg {:8.1f} us  {}\n'.format(t, 100 * t / tot, n, t / n, k))


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('# ncu --set full --clock-control none; source: {}\n'.format(src))
        for r in rows[2:]:
            f.write('\n== {}  grid {} block {}\n'.format(r[h.index('Kernel Name')][:120], r[h.index('Grid Size')], r[h.index('Block Size')]))
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write('  {:75s} {:>18s} {}\n'.format(k, r[i], units[i]))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
