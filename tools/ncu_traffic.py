"""Average DRAM traffic per launch of each kernel, from an ncu CSV of one bench.py run:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c N --csv \
        --log-file gpurun_out/x.csv python bench.py --steps 2 --warmup 3 [--workload W]
    python tools/ncu_traffic.py gpurun_out/x.csv MODEL profiles/ncu_traffic.json [profiles/x_traffic.txt]

bench.py reads profiles/ncu_traffic.json to fill roofline.traffic (bytes per launch of the dominant kernel).
Weight-packing launches (first inference only) are skipped.
"""
import collections
import csv
import json
import os
import sys

KERNELS = ['conv_f16x2_kernel', 'dwconv3x3_tma_kernel', 'pool_max_tma_kernel', 'dwconv3x3_strip_kernel', 'pool_max_strip_kernel', 'lrn_vec4_kernel', 'nchw_to_nhwc4_x4_kernel', 'nchw_to_nhwc_smallc_kernel',
           'conv_tcgen05_kernel', 'conv_ffma_kernel', 'detection_output_kernel', 'copy2d_kernel', 'affine_act_kernel', 'transpose_kernel',
           'pool_kernel', 'softmax_kernel']


def main(src, model, dst, txt=None):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mn, mv, mu, idc = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit'), h.index('ID')
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(',', ''))
        unit = r[mu]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}.get(unit, 1)
        per.setdefault(r[idc], {'name': r[kn]})[r[mn]] = v * scale
    agg = {}
    for d in per.values():
        name = next((k for k in KERNELS if k in d['name']), None)
        if name is None or 'pack_' in d['name']:
            continue
        a = agg.setdefault(name, {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
        a['n'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['rd'] += d.get('dram__bytes_read.sum', 0.0)
        a['wr'] += d.get('dram__bytes_write.sum', 0.0)
    out = json.load(open(dst)) if os.path.isfile(dst) else {}
    out[model] = {k: (a['rd'] + a['wr']) / a['n'] for k, a in agg.items()}
    json.dump(out, open(dst, 'w'), indent=1, sort_keys=True)
    if txt:
        with open(txt, 'w') as f:
            f.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; source {}\n'.format(src))
            f.write('# per kernel: launches, total us (cold-cache, serialised), DRAM read MB, DRAM write MB, GB/s over its own time\n')
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
                f.write('{:32s} n={:4d} {:10.1f} us  rd {:9.1f} MB  wr {:9.1f} MB  {:7.0f} GB/s\n'.format(
                    k, a['n'], a['us'], a['rd'] / 1e6, a['wr'] / 1e6, (a['rd'] + a['wr']) / max(a['us'], 1e-9) / 1e3))


if __name__ == '__main__':
    main(*sys.argv[1:])
