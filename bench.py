#!/usr/bin/env python
"""bench.py -- images/sec of the hot path (IR model inference through pyopenvino_b200) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload googlenet-v1] [--batch B]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A step = one pass of the fused, CUDA-graph-captured network over one batch of synthetic images per
GPU (weak scaling: the per-GPU batch is fixed).  `value` is device-timed with the batch already
resident in HBM; `e2e` is the same metric through the public API with host arrays
(`Executable_Network.infer`: pinned H2D + graph replay + D2H inside the timed region).  Rank 0
prints exactly one JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (model, default per-GPU batch, BASELINE.json config it corresponds to)
    'googlenet-v1': ('googlenet-v1', 256, 'configs[2]: GoogLeNet-v1 1x3x224x224 IR, synthetic weights, batch 256 per GPU'),
    'mnist_bn': ('mnist_bn', 1024, 'configs[1]: MNIST with BatchNorm IR, synthetic weights, batch 1024 per GPU'),
    'mnist': ('mnist', 1, 'configs[0]: MNIST CNN IR (real weights), batch 1'),
    'ssd_mobilenet_v1_coco': ('ssd_mobilenet_v1_coco', 64, 'configs[3]: SSD-MobileNet-v1 300x300 IR, synthetic weights, batch 64 per GPU'),
}
METRIC = 'images/sec (device-timed, max over ranks)'
CACHE = os.environ.get('B200OV_MODEL_CACHE', '/tmp/b200ov_models')


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'], 'bf16_tflops_sustained': p['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
                 'hw_power_brake': 0x80, 'sync_boost': 0x10, 'applications_clocks': 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas'] or [1])
    except Exception:
        return os.cpu_count() or 1


def time_cpu_port(model, budget_s, min_images=1, max_images=64):
    """Times the oracle port (numpy restatement of the reference engine, kernel_type='special', Const
    re-materialised from python tuples every inference like the reference) batch-1 on the host cores."""
    from oracle import ref_engine
    from tools.synth_bin import ensure_model, synth_input
    xml = ensure_model(model, CACHE)
    exe = ref_engine.load(xml, 'special', faithful_const=True)
    name = exe.net.inputs[0]['name']
    x = synth_input(model, batch=max_images, seed=1)
    if model == 'mnist':
        x = x * np.float32(255.0)
    exe.infer({name: x[:1]})            # warm-up call, like integrity_test.py's niter loop after the first run
    n, t0 = 0, time.time()
    while n < max_images and (n < min_images or time.time() - t0 < budget_s):
        exe.infer({name: x[n:n + 1]})
        n += 1
    dt = time.time() - t0
    return n / dt, n, dt


def run_reference(args, model, desc):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port -- the Python
    reference cannot travel to the GPU box), all host threads, same metric / config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import ref_engine
    from tools.synth_bin import ensure_model, synth_input
    xml = ensure_model(model, CACHE)
    exe = ref_engine.load(xml, 'special', faithful_const=True)
    name = exe.net.inputs[0]['name']
    per_step = {'mnist': 32, 'mnist_bn': 2, 'googlenet-v1': 1, 'ssd_mobilenet_v1_coco': 1}[model]
    x = synth_input(model, batch=per_step, seed=1)
    if model == 'mnist':
        x = x * np.float32(255.0)
    for _ in range(args.warmup):
        exe.infer({name: x[:1]})
    t0 = time.time()
    for _ in range(args.steps):
        for i in range(per_step):
            exe.infer({name: x[i:i + 1]})
    dt = time.time() - t0
    ips = args.steps * per_step / dt
    cores = blas_threads()
    sample = '{} steps x {} images, batch-1 infer loop, kernel_type=special, Const rebuilt per inference'.format(args.steps, per_step)
    line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': desc, 'model': model, 'images_per_step': per_step, 'host_cpus': os.cpu_count()},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='googlenet-v1', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=None, help='images per GPU per step')
    ap.add_argument('--cpu-budget', type=float, default=15.0, help='seconds of CPU work for cpu_baseline')
    ap.add_argument('--layers-out', default=None, help='write the per-layer roofline table (JSON) here')
    ap.add_argument('--math', default=None, choices=[None, 'fp32', 'tf32x3', 'tf32', 'f16x2', 'safe'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    model, default_batch, desc = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, model, desc)
        return
    batch = args.batch or default_batch

    import torch
    from pyopenvino_b200 import _cabi, distributed
    from pyopenvino_b200.inference_engine import IECore
    from tools import roofline
    from tools.synth_bin import ensure_model, synth_input

    rank, world, local = distributed.init()
    assert world == args.gpus or world == 1, 'launch with torchrun --nproc-per-node {}'.format(args.gpus)
    torch.cuda.set_device(local)
    peaks = load_peaks()

    if rank == 0:
        xml = ensure_model(model, CACHE)
    distributed.barrier()
    xml = ensure_model(model, CACHE)
    ie = IECore()
    net = ie.read_network(xml, xml[:-4] + '.bin')
    exe = ie.load_network(net, 'B200', batch_size=batch)
    if args.math:
        exe.kernel_type = args.math
    in_name, out_name = net.inputs[0]['name'], net.outputs[0]['name']
    x = synth_input(model, batch=batch, seed=1 + rank)       # every rank owns different images
    if model == 'mnist':
        x = x * np.float32(255.0)
    flat = exe.load_constants()
    distributed.broadcast_weights(flat)                       # rank 0's weights are the replica everyone uses
    out = exe.infer({in_name: x})[out_name]                   # builds the plan, warms up, captures the CUDA graph
    launches_per_step = exe.kernels_per_inference()
    out_dev = exe._static_out[out_name]

    def step_resident():
        exe.replay()
        if world > 1:
            distributed.gather_outputs(out_dev.t[:out_dev.size].view(out_dev.shape[0], -1))

    sampler = ClockSampler(local)
    with torch.cuda.stream(exe.stream):
        exe.stage_inputs({in_name: x})
        for _ in range(args.warmup):
            step_resident()
        exe.stream.synchronize()
        distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            e0.record()
            for _ in range(args.steps):
                step_resident()
            e1.record()
            torch.cuda.synchronize()
        distributed.barrier()
        ms_total = distributed.max_over_ranks(e0.elapsed_time(e1))

        # end to end through the public API with host arrays: every step's batch sits in pinned host memory and
        # pays its own H2D + graph replay + D2H of the result inside the timed region.  Two requests are kept in
        # flight (Executable_Network.start_async / wait), so the H2D of step i+1 overlaps the kernels of step i.
        s0 = exe.start_async({in_name: x})
        exe.wait(s0)
        bufs = [exe.request_buffer(0, in_name), exe.request_buffer(1, in_name)]
        for b in bufs:
            b[...] = x
        for _ in range(args.warmup):
            exe.wait(exe.start_async({in_name: bufs[0]}))
        torch.cuda.synchronize()
        distributed.barrier()
        t0 = time.perf_counter()
        pending = None
        for i in range(args.steps):
            slot = exe.start_async({in_name: bufs[i & 1]})
            if pending is not None:
                res = exe.wait(pending)
                if world > 1:
                    distributed.gather_outputs(torch.from_numpy(res[out_name]).cuda(non_blocking=True))
            pending = slot
        res = exe.wait(pending)
        if world > 1:
            distributed.gather_outputs(torch.from_numpy(res[out_name]).cuda(non_blocking=True))
        torch.cuda.synchronize()
        e2e_s = distributed.max_over_ranks(time.perf_counter() - t0)
        # the same, one synchronous infer() per step (no overlap), reported beside it
        x_pinned = exe.input_buffer(in_name)
        x_pinned[...] = x
        exe.infer({in_name: x_pinned})
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            exe.infer({in_name: x_pinned})
        torch.cuda.synchronize()
        e2e_sync_s = distributed.max_over_ranks(time.perf_counter() - t0)
    distributed.barrier()

    images = batch * world * args.steps
    value = images / (ms_total * 1e-3)
    e2e_value = images / e2e_s

    line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic'}

    if rank == 0:
        # ---- per-layer roofline: eager pass of the same fused plan with CUDA events per layer --------
        work = roofline.layer_work(exe)
        steps = exe.profile_steps({in_name: x}, iters=3)
        tf32x3_peak = peaks['bf16_tflops_sustained']          # tensor denominator = measured dense bf16 (stated)
        fam = {}
        layers = []
        total_ms = sum(s['ms'] for s in steps)
        for s in steps:
            w = work.get(s['id'])
            if w is None:
                continue
            f = fam.setdefault(w['kind'], {'ms': 0.0, 'flops': 0, 'bytes': 0, 'launches': 0})
            f['ms'] += s['ms']
            f['flops'] += w['flops']
            f['bytes'] += w['bytes']
            f['launches'] += 1
            roof = roofline.roofline_ms(w, peaks['hbm_gbs'], tf32x3_peak)
            layers.append({'name': s['name'], 'kind': w['kind'], 'ms': s['ms'], 'gflop': w['flops'] / 1e9, 'mbytes': w['bytes'] / 1e6,
                           'tflops': w['flops'] / (s['ms'] * 1e-3) / 1e12 if s['ms'] > 0 else 0.0,
                           'gbs': w['bytes'] / (s['ms'] * 1e-3) / 1e9 if s['ms'] > 0 else 0.0, 'roofline_ms': roof,
                           'frac_of_roofline': roof / s['ms'] if s['ms'] > 0 else 0.0})
        # families -> the CUDA kernel that runs them; the roofline object describes the dominant KERNEL
        def kernel_of(kind):
            if kind.startswith('conv') or kind == 'matmul':
                return 'conv_f16x2_kernel'
            return {'depthwise': 'dwconv3x3_strip_kernel', 'maxpool': 'pool_max_strip_kernel', 'lrn': 'lrn_vec4_kernel',
                    'input_layout': 'nchw_to_nhwc_smallc_kernel'}.get(kind, kind)
        kern = {}
        for kind, f in fam.items():
            k = kern.setdefault(kernel_of(kind), {'ms': 0.0, 'flops': 0, 'bytes': 0, 'launches': 0, 'families': []})
            for key in ('ms', 'flops', 'bytes', 'launches'):
                k[key] += f[key]
            k['families'].append(kind)
        top_kind = max(kern, key=lambda k: kern[k]['ms'])
        top = kern[top_kind]
        ai = top['flops'] / max(top['bytes'], 1)
        # the f16x2 contraction spends 3 tensor-core MMAs per FP32 product: its ridge point uses peak / 3
        mma_per_product = 3 if top_kind == 'conv_f16x2_kernel' else 1
        tensor_bound = ai > (tf32x3_peak / mma_per_product) * 1e12 / (peaks['hbm_gbs'] * 1e9)
        if tensor_bound:
            achieved = top['flops'] / (top['ms'] * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'achieved': achieved, 'peak': tf32x3_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf32x3_peak,
                    'mma_per_fp32_product': mma_per_product, 'tensor_pipe_frac': achieved * mma_per_product / tf32x3_peak}
        else:
            achieved = top['bytes'] / (top['ms'] * 1e-3) / 1e9
            roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs']}
        traffic = None
        tpath = os.path.join(REPO, 'profiles', 'ncu_traffic.json')      # dram bytes per launch from an ncu capture of this command
        if os.path.isfile(tpath):
            try:
                traffic = json.load(open(tpath)).get(model, {}).get(top_kind)
            except Exception:
                traffic = None
        roof.update({'traffic': traffic, 'kernel': top_kind, 'families': sorted(top['families']), 'launches_per_step': top['launches'],
                     'share_of_step': top['ms'] / total_ms if total_ms else None,
                     'algorithmic_gflop_per_step': top['flops'] / 1e9, 'algorithmic_mbytes_per_step': top['bytes'] / 1e6,
                     'algorithmic_mbytes_per_launch': top['bytes'] / 1e6 / max(top['launches'], 1),
                     'avg_launch_ms': top['ms'] / max(top['launches'], 1), 'peak_source': peaks['source'],
                     'peak_note': 'tensor peak = measured dense bf16 (sustained); an FP32-accurate product costs 3 kind::f16 MMAs '
                                  '(f16x2 split), so tensor_pipe_frac = 3 * achieved / peak is the share of the tensor pipe in use'})
        model_roof_ms = sum(l['roofline_ms'] for l in layers)
        line['roofline'] = roof
        line['model_roofline'] = {'sum_layer_roofline_ms': model_roof_ms, 'frac': model_roof_ms / (ms_total / args.steps),
                                  'families': {k: {'ms': v['ms'], 'share': v['ms'] / total_ms, 'tflops': v['flops'] / (v['ms'] * 1e-3) / 1e12,
                                                   'gbs': v['bytes'] / (v['ms'] * 1e-3) / 1e9} for k, v in fam.items() if v['ms'] > 0}}
        if args.layers_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
            json.dump({'workload': desc, 'batch': batch, 'layers': layers, 'families': line['model_roofline']['families']},
                      open(args.layers_out, 'w'), indent=1)
        working_set_mb = sum(w['bytes'] for w in work.values()) / 1e6
        line['config'] = {'workload': desc, 'model': model, 'batch_per_gpu': batch, 'global_batch': batch * world,
                          'input_shape': list(x.shape), 'parallelism': 'dp{} (batch-sharded replicas, no data-path collective)'.format(world),
                          'l2': 'no flush: per-step working set {:.0f} MB > 126 MB L2'.format(working_set_mb)
                          if working_set_mb > 126 else 'working set {:.0f} MB fits L2 (not flushed)'.format(working_set_mb),
                          'fused_cuda_graph': True, 'math': args.math or 'auto'}
        line['clocks'] = sampler.summary()
        line['e2e'] = {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': int(x.nbytes) * world,
                       'd2h_bytes_per_step': int(out.nbytes) * world, 'ms_per_step': e2e_s / args.steps * 1e3,
                       'api': 'Executable_Network.start_async / wait, 2 requests in flight (H2D of step i+1 overlaps step i)',
                       'sync_infer_value': images / e2e_sync_s, 'sync_infer_ms_per_step': e2e_sync_s / args.steps * 1e3}
        line['gpu_launches'] = launches_per_step * args.steps
        line['launches_per_step'] = launches_per_step
        if world == 1:
            ips, n, dt = time_cpu_port(model, args.cpu_budget)
            line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': blas_threads(), 'kind': 'port',
                                    'sample': '{} batch-1 inferences in {:.1f} s, oracle port of the reference engine, kernel_type=special, '
                                              'Const rebuilt per inference, host_cpus={}'.format(n, dt, os.cpu_count())}
        print(json.dumps(line), flush=True)
    distributed.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
