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


def set_blas_threads():
    """Give the CPU reference every host core this process may run on, whatever the launcher exported
    (torchrun sets OMP_NUM_THREADS=1): OpenBLAS' pool is resized at run time.  Returns the limiter (keep it alive)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=len(os.sched_getaffinity(0)), user_api='blas')
    except Exception:
        return None


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
    limiter = set_blas_threads()          # noqa: F841  (explicit: not whatever OMP_NUM_THREADS the launcher exported)
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
            'config': {'workload': desc, 'model': model, 'images_per_step': per_step, 'host_cpus': os.cpu_count(),
                       'blas_threads': cores, 'omp_num_threads_env': os.environ.get('OMP_NUM_THREADS')},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                             'blas_threads': cores},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def kernel_of(kind):
    """layer family -> the CUDA kernel that runs it"""
    if kind.startswith('conv') or kind == 'matmul':
        return 'conv_f16x2_kernel'
    return {'depthwise': 'dwconv3x3_tma_kernel', 'maxpool': 'pool_max_tma_kernel', 'lrn': 'lrn_vec4_kernel',
            'input_layout': 'nchw_to_nhwc4_x4_kernel'}.get(kind, kind)


def layer_table(exe, in_name, x, peaks):
    """Per-layer roofline: an eager pass of the same fused plan with CUDA events per layer."""
    from tools import roofline
    work = roofline.layer_work(exe)
    steps = exe.profile_steps({in_name: x}, iters=3)
    tensor_peak = peaks['bf16_tflops_sustained']          # tensor denominator = measured dense bf16 (stated)
    fam, layers = {}, []
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
        roof = roofline.roofline_ms(w, peaks['hbm_gbs'], tensor_peak)
        layers.append({'name': s['name'], 'kind': w['kind'], 'ms': s['ms'], 'gflop': w['flops'] / 1e9, 'mbytes': w['bytes'] / 1e6,
                       'tflops': w['flops'] / (s['ms'] * 1e-3) / 1e12 if s['ms'] > 0 else 0.0,
                       'gbs': w['bytes'] / (s['ms'] * 1e-3) / 1e9 if s['ms'] > 0 else 0.0, 'roofline_ms': roof,
                       'frac_of_roofline': roof / s['ms'] if s['ms'] > 0 else 0.0})
    kern = {}
    for kind, f in fam.items():
        k = kern.setdefault(kernel_of(kind), {'ms': 0.0, 'flops': 0, 'bytes': 0, 'launches': 0, 'families': []})
        for key in ('ms', 'flops', 'bytes', 'launches'):
            k[key] += f[key]
        k['families'].append(kind)
    return work, layers, fam, kern, total_ms


def kernel_roofline(name, k, total_ms, peaks, model):
    """The `roofline` object for one kernel (all its launches of a step)."""
    tensor_peak = peaks['bf16_tflops_sustained']
    ai = k['flops'] / max(k['bytes'], 1)
    # the f16x2 contraction spends 3 tensor-core MMAs per FP32 product: its ridge point uses peak / 3
    mma_per_product = 3 if name == 'conv_f16x2_kernel' else 1
    tensor_bound = ai > (tensor_peak / mma_per_product) * 1e12 / (peaks['hbm_gbs'] * 1e9)
    if tensor_bound:
        achieved = k['flops'] / (k['ms'] * 1e-3) / 1e12
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': achieved / tensor_peak,
                'mma_per_fp32_product': mma_per_product, 'tensor_pipe_frac': achieved * mma_per_product / tensor_peak}
    else:
        achieved = k['bytes'] / (k['ms'] * 1e-3) / 1e9
        roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs']}
    traffic = None
    tpath = os.path.join(REPO, 'profiles', 'ncu_traffic.json')      # dram bytes per launch from an ncu capture of this command
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(model, {}).get(name)
        except Exception:
            traffic = None
    roof.update({'traffic': traffic, 'kernel': name, 'families': sorted(k['families']), 'launches_per_step': k['launches'],
                 'share_of_step': k['ms'] / total_ms if total_ms else None,
                 'algorithmic_gflop_per_step': k['flops'] / 1e9, 'algorithmic_mbytes_per_step': k['bytes'] / 1e6,
                 'algorithmic_mbytes_per_launch': k['bytes'] / 1e6 / max(k['launches'], 1),
                 'avg_launch_ms': k['ms'] / max(k['launches'], 1), 'peak_source': peaks['source']})
    if name == 'conv_f16x2_kernel':
        roof['peak_note'] = ('tensor peak = measured dense bf16 (sustained); an FP32-accurate product costs 3 kind::f16 MMAs '
                             '(f16x2 split), so tensor_pipe_frac = 3 * achieved / peak is the share of the tensor pipe in use')
    return roof


def time_e2e(exe, in_name, out_name, x, steps, warmup, world, dtype=np.float32):
    """End to end through the public API with host arrays: every step's batch sits in pinned host memory (the request
    slot's own buffer, element type `dtype`) and pays its own H2D + graph replay + D2H of the result inside the timed
    region.  Two requests are kept in flight (start_async / wait), so the H2D of step i+1 overlaps the kernels of step i.
    The slot is named explicitly, so the zero-copy path does not depend on how many requests ran before."""
    import torch
    from pyopenvino_b200 import distributed
    nreq = exe.NUM_REQUESTS
    bufs = [exe.request_buffer(s, in_name, dtype) for s in range(nreq)]
    for b in bufs:
        b[...] = x
    # The all-gather of the result rows runs on its own stream: queued on the engine's stream it would sit behind the NEXT
    # request's kernels, and its (pageable) upload would block the host until they finish -- the H2D of the request after
    # that could then no longer overlap anything (round-2 finding: 8.9 ms per step at 8 GPUs = H2D + kernels in series).
    side = torch.cuda.Stream() if world > 1 else None

    def gather(res):
        if world > 1:
            with torch.cuda.stream(side):
                distributed.gather_outputs(torch.from_numpy(res[out_name]).cuda(non_blocking=True))

    for i in range(max(warmup, 1)):
        gather(exe.wait(exe.start_async({in_name: bufs[i % nreq]}, slot=i % nreq)))
    torch.cuda.synchronize()
    distributed.barrier()
    t0 = time.perf_counter()
    pending, res = None, None
    for i in range(steps):
        slot = exe.start_async({in_name: bufs[i % nreq]}, slot=i % nreq)
        if pending is not None:
            res = exe.wait(pending)
            gather(res)
        pending = slot
    res = exe.wait(pending)
    gather(res)
    torch.cuda.synchronize()
    return distributed.max_over_ranks(time.perf_counter() - t0), res[out_name]


def measure(model, batch, desc, args, peaks, rank, world, local, primary, storage='f32', light=False):
    """One workload at `world` GPUs: device-timed value, e2e (FP32 and uint8 host input), per-layer roofline.
    Returns the fields of a bench line (rank 0) or None."""
    import torch
    from pyopenvino_b200 import distributed
    from pyopenvino_b200.inference_engine import IECore
    from tools.synth_bin import ensure_model, synth_input

    if rank == 0:
        ensure_model(model, CACHE)
    distributed.barrier()
    xml = ensure_model(model, CACHE)
    ie = IECore()
    net = ie.read_network(xml, xml[:-4] + '.bin')
    exe = ie.load_network(net, 'B200', batch_size=batch, storage=storage)
    if args.math:
        exe.kernel_type = args.math
    in_name, out_name = net.inputs[0]['name'], net.outputs[0]['name']
    x = synth_input(model, batch=batch, seed=1 + rank)       # every rank owns different images
    if model == 'mnist':
        x = x * np.float32(255.0)
    exe.broadcast_constants(src=0)                            # rank 0's weights are the replica everyone uses
    out = exe.infer({in_name: x})[out_name]                   # builds the plan, warms up, captures the CUDA graph
    launches_per_step = exe.kernels_per_inference()
    out_dev = exe._static_out[out_name]
    steps, warmup = args.steps, args.warmup

    def step_resident():
        exe.replay()
        if world > 1:
            distributed.gather_outputs(out_dev.t[:out_dev.size].view(out_dev.shape[0], -1))

    def timed(nsteps, sampler=None):
        distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.__enter__()
        e0.record()
        for _ in range(nsteps):
            step_resident()
        e1.record()
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.__exit__(None, None, None)
        distributed.barrier()
        return distributed.max_over_ranks(e0.elapsed_time(e1))

    sampler = ClockSampler(local)
    res = {}
    with torch.cuda.stream(exe.stream):
        exe.stage_inputs({in_name: x})
        for _ in range(warmup):
            step_resident()
        exe.stream.synchronize()
        ms_total = timed(steps, sampler)
        ms_step = ms_total / steps
        # the same replay loop held for >= args.sustain seconds: does the number survive the power limit?
        sustained = None
        if args.sustain > 0 and (primary or args.sustain_all) and not light:
            n_sus = max(steps, int(args.sustain * 1e3 / max(ms_step, 1e-3)) + 1)
            sus_sampler = ClockSampler(local)
            ms_sus = timed(n_sus, sus_sampler)
            sustained = {'value': batch * world * n_sus / (ms_sus * 1e-3), 'unit': 'images/s', 'steps': n_sus, 'seconds': ms_sus * 1e-3,
                         'ms_per_step': ms_sus / n_sus, 'clocks': sus_sampler.summary()}
        # end to end, FP32 host arrays (headline) ...
        e2e_s, _ = time_e2e(exe, in_name, out_name, x, steps, warmup, world, np.float32)
        # ... and the same images as uint8 frames (1 byte per value over PCIe; widened in the layout kernel)
        x8 = np.clip(np.rint(x if model in ('mnist', 'ssd_mobilenet_v1_coco') else x * 255.0), 0, 255).astype(np.uint8)
        e2e8_s, _ = time_e2e(exe, in_name, out_name, x8, steps, warmup, world, np.uint8)
        exe._select_graph({in_name: x})
        e2e_sync_s = None
        if not light:
            # one synchronous infer() per step (no overlap), reported beside it
            x_pinned = exe.input_buffer(in_name)
            x_pinned[...] = x
            exe.infer({in_name: x_pinned})
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                exe.infer({in_name: x_pinned})
            torch.cuda.synchronize()
            e2e_sync_s = distributed.max_over_ranks(time.perf_counter() - t0)
    distributed.barrier()

    images = batch * world * steps
    res = {'value': images / (ms_total * 1e-3), 'unit': 'images/s', 'ms_per_step': ms_step}
    if rank != 0:
        del exe
        return None
    if light:
        res.update({'storage': storage, 'arena_mb': exe._arena.bytes() / 1e6, 'working_set_mb': exe._arena.peak_bytes() / 1e6,
                    'launches_per_step': launches_per_step, 'clocks': sampler.summary(),
                    'e2e': {'value': images / e2e_s, 'unit': 'images/s', 'input_dtype': 'float32', 'ms_per_step': e2e_s / steps * 1e3},
                    'e2e_u8': {'value': images / e2e8_s, 'unit': 'images/s', 'input_dtype': 'uint8', 'ms_per_step': e2e8_s / steps * 1e3}})
        fam = {}
        work = __import__('tools.roofline', fromlist=['layer_work']).layer_work(exe)
        for s_ in exe.profile_steps({in_name: x}, iters=3):
            w = work.get(s_['id'])
            if w is not None:
                fam[w['kind']] = fam.get(w['kind'], 0.0) + s_['ms']
        res['family_ms'] = fam
        del exe
        return res
    work, layers, fam, kern, total_ms = layer_table(exe, in_name, x, peaks)
    top_kind = max(kern, key=lambda k: kern[k]['ms'])
    res['roofline'] = kernel_roofline(top_kind, kern[top_kind], total_ms, peaks, model)
    res['kernel_rooflines'] = {k: kernel_roofline(k, v, total_ms, peaks, model) for k, v in kern.items()
                               if k in ('conv_f16x2_kernel', 'dwconv3x3_tma_kernel', 'pool_max_tma_kernel', 'lrn_vec4_kernel')
                               and k != top_kind and v['ms'] > 0}
    model_roof_ms = sum(l['roofline_ms'] for l in layers)
    res['model_roofline'] = {'sum_layer_roofline_ms': model_roof_ms, 'frac': model_roof_ms / ms_step,
                             'families': {k: {'ms': v['ms'], 'share': v['ms'] / total_ms, 'tflops': v['flops'] / (v['ms'] * 1e-3) / 1e12,
                                              'gbs': v['bytes'] / (v['ms'] * 1e-3) / 1e9} for k, v in fam.items() if v['ms'] > 0}}
    if args.layers_out and primary:
        os.makedirs(os.path.dirname(os.path.abspath(args.layers_out)), exist_ok=True)
        json.dump({'workload': desc, 'batch': batch, 'layers': layers, 'families': res['model_roofline']['families']},
                  open(args.layers_out, 'w'), indent=1)
    working_set_mb = sum(w['bytes'] for w in work.values()) / 1e6
    res['config'] = {'workload': desc, 'model': model, 'batch_per_gpu': batch, 'global_batch': batch * world,
                     'input_shape': list(x.shape), 'parallelism': 'dp{} (batch-sharded replicas, no data-path collective)'.format(world),
                     'l2': 'no flush: per-step working set {:.0f} MB > 126 MB L2'.format(working_set_mb)
                     if working_set_mb > 126 else 'working set {:.0f} MB fits L2 (not flushed)'.format(working_set_mb),
                     'fused_cuda_graph': True, 'math': args.math or 'auto', 'storage': storage,
                     'arena_mb': exe._arena.bytes() / 1e6 if exe._arena is not None else None,
                     'working_set_mb': exe._arena.peak_bytes() / 1e6 if exe._arena is not None else None}
    res['clocks'] = sampler.summary()
    if sustained is not None:
        res['sustained'] = sustained
    api = 'Executable_Network.start_async(slot=) / wait, 2 requests in flight (H2D of step i+1 overlaps step i)'
    res['e2e'] = {'value': images / e2e_s, 'unit': 'images/s', 'h2d_bytes_per_step': int(x.nbytes) * world,
                  'd2h_bytes_per_step': int(out.nbytes) * world, 'ms_per_step': e2e_s / steps * 1e3, 'input_dtype': 'float32',
                  'api': api, 'sync_infer_value': images / e2e_sync_s if e2e_sync_s else None,
                  'sync_infer_ms_per_step': e2e_sync_s / steps * 1e3 if e2e_sync_s else None}
    res['e2e_u8'] = {'value': images / e2e8_s, 'unit': 'images/s', 'h2d_bytes_per_step': int(x8.nbytes) * world,
                     'd2h_bytes_per_step': int(out.nbytes) * world, 'ms_per_step': e2e8_s / steps * 1e3, 'input_dtype': 'uint8',
                     'api': api, 'note': 'same images as uint8 frames (Parameter.py:13 accepts any array-like); widened on the device'}
    res['gpu_launches'] = launches_per_step * steps
    res['launches_per_step'] = launches_per_step
    del exe
    return res


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
    ap.add_argument('--sustain', type=float, default=3.0, help='seconds of back-to-back replays for the `sustained` field (0 = off)')
    ap.add_argument('--sustain-all', action='store_true', help='also for the secondary workloads')
    ap.add_argument('--no-secondary', action='store_true', help='measure only --workload')
    ap.add_argument('--no-f16', action='store_true', help='skip the f16_storage measurement')
    ap.add_argument('--storage', default='f32', choices=['f32', 'f16'],
                    help="feature-map storage of the headline line (the default line always adds an `f16_storage` object beside the FP32 value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    model, default_batch, desc = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, model, desc)
        return
    batch = args.batch or default_batch

    import torch
    from pyopenvino_b200 import device, distributed

    rank, world, local = distributed.init()
    assert world == args.gpus or world == 1, 'launch with torchrun --nproc-per-node {}'.format(args.gpus)
    torch.cuda.set_device(local)
    if world > 1:
        device.bind_host_thread(local, world)                  # each rank's python thread on its own slice of the host CPUs
    peaks = load_peaks()

    line = {'metric': METRIC, 'value': None, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic'}
    res = measure(model, batch, desc, args, peaks, rank, world, local, primary=True, storage=args.storage)
    if rank == 0:
        line.update(res)
        if args.storage != 'f32':
            line['dtype'] = 'f32 arithmetic, f16 feature-map storage (opt-in mode; NOT the headline configuration)'
    # opt-in FP16 feature-map storage, same workload, same run: a separately named field -- `value` above stays FP32
    if args.storage == 'f32' and not args.no_f16:
        r16 = measure(model, batch, desc, args, peaks, rank, world, local, primary=False, storage='f16', light=True)
        if rank == 0:
            r16['note'] = ("load_network(..., storage='f16'): NHWC feature maps kept in HBM as FP16, FP32 arithmetic; own tolerance "
                           "statement in tests/test_gpu_f16_storage.py; the headline `value` / `e2e` are FP32 storage")
            r16['speedup_vs_f32_storage'] = r16['value'] / line['value']
            line['f16_storage'] = r16
    # the other BASELINE.json configurations, same run, same N (the driver only ever launches the default command)
    secondary = {}
    if not args.no_secondary and args.batch is None:
        for name in ('ssd_mobilenet_v1_coco', 'mnist_bn', 'mnist'):
            if name == args.workload:
                continue
            m2, b2, d2 = WORKLOADS[name]
            r2 = measure(m2, b2, d2, args, peaks, rank, world, local, primary=False)
            if rank == 0:
                keep = ('value', 'unit', 'ms_per_step', 'e2e', 'e2e_u8', 'roofline', 'kernel_rooflines', 'model_roofline', 'config',
                        'clocks', 'launches_per_step', 'sustained')
                secondary[name] = {k: r2[k] for k in keep if k in r2}
    if rank == 0:
        if secondary:
            line['secondary'] = secondary
        if world == 1:
            limiter = set_blas_threads()      # noqa: F841
            ips, n, dt = time_cpu_port(model, args.cpu_budget)
            line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': blas_threads(), 'kind': 'port', 'blas_threads': blas_threads(),
                                    'sample': '{} batch-1 inferences in {:.1f} s, oracle port of the reference engine, kernel_type=special, '
                                              'Const rebuilt per inference, host_cpus={}'.format(n, dt, os.cpu_count())}
        print(json.dumps(line), flush=True)
    distributed.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
