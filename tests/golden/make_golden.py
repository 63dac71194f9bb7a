"""Generate the golden vectors under tests/golden/ by running the LIVE reference.

Run once in the authoring container (the reference is mounted read-only at /root/reference and
cannot travel to the GPU box):

    python tests/golden/make_golden.py

What it does
  * builds a scratch directory holding a symlink `pyopenvino -> /root/reference/pyopenvino`, the IR
    XML files and `.bin` files (`models/mnist.bin` is the shipped real one, the other three are
    synthetic, `tools/synth_bin.py` seed 0), chdir()s there (the reference finds its plugins and
    `common_def` relative to cwd, inference_engine.py:17,51) and imports the unmodified reference;
  * records (1) MNIST end-to-end on resources/mnist2.png for 'numpy' and 'special' with every node
    output, (2) the resources/node_args_6.pickle Convolution known-answer (f16 as stored + an f32
    crop), (3) per-op vectors through each reference plugin's compute(kernel_type='numpy'),
    (4) synthetic-weight end-to-end runs of mnist_bn / googlenet-v1 / ssd_mobilenet_v1_coco with the
    final outputs and a strided sample of every node output.

Only data produced by running the reference is stored; no reference source is copied.
"""
import hashlib
import json
import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
SCRATCH = '/tmp/b200ov_refscratch'


def setup():
    sys.path.insert(0, REPO)
    from tools.synth_bin import ensure_model
    os.makedirs(SCRATCH, exist_ok=True)
    link = os.path.join(SCRATCH, 'pyopenvino')
    if not os.path.islink(link):
        os.symlink(os.path.join(REF, 'pyopenvino'), link)
    for m in ('mnist', 'mnist_bn', 'googlenet-v1', 'ssd_mobilenet_v1_coco'):
        ensure_model(m, os.path.join(SCRATCH, 'models'))
    sys.path.remove(REPO)
    os.chdir(SCRATCH)
    sys.path.insert(0, SCRATCH)
    sys.dont_write_bytecode = True


def sample(arr, k=64):
    flat = np.asarray(arr).ravel()
    if flat.size <= k:
        return flat.copy()
    idx = np.linspace(0, flat.size - 1, k).astype(np.int64)
    return flat[idx].copy()


def run_model(IECore, model, blob, kernel_type):
    ie = IECore()
    net = ie.read_network('models/{}.xml'.format(model), 'models/{}.bin'.format(model))
    exe = ie.load_network(net, 'CPU')
    exe.kernel_type = kernel_type
    t0 = time.time()
    res = exe.infer({net.inputs[0]['name']: blob})
    dt = time.time() - t0
    nodes = {}
    G = net.G
    for nid in G.nodes:
        n = G.nodes[nid]
        if 'output' in n:
            p = next(iter(n['output']))
            if 'data' in n['output'][p]:
                nodes[n['name']] = np.asarray(n['output'][p]['data'])
    return res[net.outputs[0]['name']], nodes, dt


def golden_mnist(IECore):
    import cv2
    img = cv2.imread(os.path.join(REF, 'resources', 'mnist2.png'))
    blob = cv2.split(img)[0].reshape(1, 1, 28, 28).astype(np.float32)
    out = {'input': blob}
    for kt in ('numpy', 'special'):
        final, nodes, dt = run_model(IECore, 'mnist', blob, kt)
        print('mnist', kt, dt, np.argsort(final[0])[::-1])
        out['final_' + kt] = final
        if kt == 'numpy':
            names = sorted(nodes)
            out['node_names'] = np.array(json.dumps(names))
            for i, nm in enumerate(names):
                out['node_{}'.format(i)] = nodes[nm]
    # a second real image for good measure
    img = cv2.imread(os.path.join(REF, 'resources', 'mnist7.png'))
    blob7 = cv2.split(img)[0].reshape(1, 1, 28, 28).astype(np.float32)
    final7, _, _ = run_model(IECore, 'mnist', blob7, 'special')
    out['input7'] = blob7
    out['final7_special'] = final7
    np.savez_compressed(os.path.join(HERE, 'mnist_e2e.npz'), **out)


def golden_conv_kat():
    import op_plugins.Convolution as conv   # resolvable through sys.path 'pyopenvino'
    with open(os.path.join(REF, 'resources', 'node_args_6.pickle'), 'rb') as f:
        node, inputs = pickle.load(f)
    out = {'node': np.array(json.dumps({'name': node['name'], 'data': dict(node['data'])})),
           'x_f16': inputs[0], 'w_f16': inputs[1]}
    for kt in ('special', 'numpy'):
        r = conv.compute(node, inputs, kernel_type=kt)[2]
        out['sha_' + kt] = np.array(hashlib.sha256(r.tobytes()).hexdigest())
        print('conv KAT', kt, r.dtype, r.shape, out['sha_' + kt])
        if kt == 'special':
            out['y_f16_special'] = r
    # f32 crop: top-left 3x64x64 of the input -> 32x32x32 output with the same attrs (s2, pads 0/1)
    x = inputs[0][:, :, :64, :64].astype(np.float32)
    w = inputs[1].astype(np.float32)
    n32 = {'name': node['name'], 'type': 'Convolution', 'data': dict(node['data']),
           'input': {0: {'precision': 'FP32', 'dims': x.shape}, 1: {'precision': 'FP32', 'dims': w.shape}},
           'output': {2: {'precision': 'FP32', 'dims': (1, 32, 32, 32)}}}
    out['x_f32_crop'] = x
    out['w_f32'] = w
    for kt in ('special', 'numpy'):
        out['y_f32_crop_' + kt] = conv.compute(n32, {0: x, 1: w}, kernel_type=kt)[2]
    np.savez_compressed(os.path.join(HERE, 'conv_kat.npz'), **out)


def _port(arr):
    prec = {np.dtype('float32'): 'FP32', np.dtype('int64'): 'I64'}[np.asarray(arr).dtype]
    return {'precision': prec, 'dims': tuple(np.asarray(arr).shape)}


def golden_ops():
    import importlib
    rng = np.random.default_rng(1234)
    cases = []

    def f32(*shape, scale=1.0, positive=False):
        a = rng.standard_normal(shape).astype(np.float32) * np.float32(scale)
        return np.abs(a) if positive else a

    def conv_case(tag, cin, cout, k, s, pb, pe, hw, auto_pad='explicit'):
        x = f32(1, cin, hw[0], hw[1])
        w = f32(cout, cin, k, k, scale=(2.0 / (cin * k * k)) ** 0.5)
        data = {'strides': '{}, {}'.format(s, s), 'dilations': '1, 1', 'pads_begin': '{}, {}'.format(*pb),
                'pads_end': '{}, {}'.format(*pe), 'auto_pad': auto_pad}
        cases.append(dict(tag=tag, type='Convolution', data=data, ins={0: x, 1: w}, kts=['numpy', 'special']))

    conv_case('conv3x3_valid_c1', 1, 32, 3, 1, (0, 0), (0, 0), (28, 28), 'valid')
    conv_case('conv3x3_valid_32_64', 32, 64, 3, 1, (0, 0), (0, 0), (13, 13), 'valid')
    conv_case('conv3x3_same_p1', 16, 24, 3, 1, (1, 1), (1, 1), (14, 14), 'same_upper')
    conv_case('conv3x3_s2_asym', 3, 32, 3, 2, (0, 0), (1, 1), (30, 30), 'same_upper')
    conv_case('conv3x3_s2_sym_odd', 8, 12, 3, 2, (1, 1), (1, 1), (19, 19), 'explicit')
    conv_case('conv1x1', 48, 20, 1, 1, (0, 0), (0, 0), (7, 7), 'explicit')
    conv_case('conv1x1_odd_cout', 32, 273, 1, 1, (0, 0), (0, 0), (3, 3), 'explicit')
    conv_case('conv5x5_p2', 16, 32, 5, 1, (2, 2), (2, 2), (14, 14), 'explicit')
    conv_case('conv7x7_s2_p3', 3, 16, 7, 2, (3, 3), (3, 3), (32, 32), 'explicit')

    def dw_case(tag, c, s, pb, pe, hw):
        x = f32(1, c, hw[0], hw[1])
        w = f32(c, 1, 1, 3, 3, scale=(2.0 / 9) ** 0.5)
        data = {'strides': '{}, {}'.format(s, s), 'dilations': '1, 1', 'pads_begin': '{}, {}'.format(*pb),
                'pads_end': '{}, {}'.format(*pe), 'auto_pad': 'same_upper'}
        cases.append(dict(tag=tag, type='GroupConvolution', data=data, ins={0: x, 1: w}, kts=['numpy']))

    dw_case('dw_s1_p1', 32, 1, (1, 1), (1, 1), (20, 20))
    dw_case('dw_s2_asym', 16, 2, (0, 0), (1, 1), (20, 20))
    dw_case('dw_s2_sym_odd', 8, 2, (1, 1), (1, 1), (19, 19))
    dw_case('dw_s1_small', 64, 1, (1, 1), (1, 1), (3, 3))

    def mm_case(tag, a, b, ta, tb):
        cases.append(dict(tag=tag, type='MatMul', data={'transpose_a': ta, 'transpose_b': tb}, ins={0: a, 1: b}, kts=['numpy']))

    mm_case('mm_tb', f32(1, 576), f32(64, 576, scale=0.05), 'false', 'true')
    mm_case('mm_tb_rows', f32(5, 128), f32(10, 128, scale=0.1), 'false', 'true')
    mm_case('mm_plain', f32(3, 40), f32(40, 12, scale=0.1), 'false', 'false')
    mm_case('mm_ta_tb', f32(40, 3), f32(12, 40, scale=0.1), 'true', 'true')

    def pool_case(tag, typ, c, hw, k, s, pb, pe, rounding, auto_pad, positive):
        x = f32(1, c, hw[0], hw[1], positive=positive)
        data = {'strides': '{}, {}'.format(s, s), 'kernel': '{}, {}'.format(k, k), 'pads_begin': '{}, {}'.format(*pb),
                'pads_end': '{}, {}'.format(*pe), 'rounding_type': rounding, 'auto_pad': auto_pad}
        if typ == 'AvgPool':
            data['exclude-pad'] = 'false'
        cases.append(dict(tag=tag, type=typ, data=data, ins={0: x}, kts=['numpy']))

    pool_case('max3x3_s2_ceil', 'MaxPool', 8, (12, 12), 3, 2, (0, 0), (0, 0), 'ceil', 'explicit', True)
    pool_case('max3x3_s2_ceil_neg', 'MaxPool', 8, (14, 14), 3, 2, (0, 0), (0, 0), 'ceil', 'explicit', False)
    pool_case('max3x3_s1_p1', 'MaxPool', 12, (7, 7), 3, 1, (1, 1), (1, 1), 'ceil', 'explicit', True)
    pool_case('max3x3_s1_p1_neg', 'MaxPool', 12, (7, 7), 3, 1, (1, 1), (1, 1), 'ceil', 'explicit', False)
    pool_case('max2x2_s2_valid', 'MaxPool', 32, (26, 26), 2, 2, (0, 0), (0, 0), 'floor', 'valid', True)
    pool_case('max2x2_s2_valid_odd', 'MaxPool', 16, (11, 11), 2, 2, (0, 0), (0, 0), 'floor', 'valid', False)
    pool_case('avg7x7', 'AvgPool', 64, (7, 7), 7, 1, (0, 0), (0, 0), 'ceil', 'explicit', False)
    pool_case('avg3x3_s1', 'AvgPool', 4, (8, 8), 3, 1, (0, 0), (0, 0), 'floor', 'explicit', False)
    pool_case('avg2x2_s2', 'AvgPool', 4, (9, 9), 2, 2, (0, 0), (0, 0), 'floor', 'valid', False)

    def ew_case(tag, typ, ins, data=None):
        cases.append(dict(tag=tag, type=typ, data=data or {'auto_broadcast': 'numpy'}, ins=ins, kts=['numpy']))

    ew_case('add_bias', 'Add', {0: f32(1, 24, 5, 6), 1: f32(1, 24, 1, 1)})
    ew_case('add_scalar', 'Add', {0: f32(1, 3, 8, 8), 1: f32(1, 1, 1, 1)})
    ew_case('add_2d', 'Add', {0: f32(1, 10), 1: f32(1, 10)})
    ew_case('add_same', 'Add', {0: f32(1, 6, 4, 4), 1: f32(1, 6, 4, 4)})
    ew_case('mul_scale', 'Multiply', {0: f32(1, 16, 7, 7), 1: f32(1, 16, 1, 1)})
    ew_case('mul_scalar_port0', 'Multiply', {0: f32(1, 1, 1, 1), 1: f32(1, 3, 10, 10)})
    xr = f32(1, 8, 6, 6)
    xr[0, 0, 0, 0] = -0.0
    xr[0, 0, 0, 1] = 0.0
    ew_case('relu', 'ReLU', {0: xr}, {})
    ew_case('relu_2d', 'ReLU', {0: f32(1, 64)}, {})
    ew_case('clamp6', 'Clamp', {0: f32(1, 8, 6, 6, scale=4.0)}, {'min': '0', 'max': '6'})
    ew_case('softmax10', 'SoftMax', {0: f32(1, 10, scale=3.0)}, {'axis': '1'})
    ew_case('softmax1000', 'SoftMax', {0: f32(1, 1000)}, {'axis': '1'})
    ew_case('sigmoid', 'Sigmoid', {0: f32(1, 1, 50, 91, scale=3.0)}, {})
    ew_case('lrn', 'LRN', {0: f32(1, 64, 6, 6, scale=3.0), 1: np.array([1], dtype=np.int64)},
            {'alpha': '9.9999997473787516e-05', 'beta': '0.75', 'bias': '1', 'size': '5'})
    ew_case('lrn_small_c', 'LRN', {0: f32(1, 3, 4, 4, scale=10.0), 1: np.array([1], dtype=np.int64)},
            {'alpha': '0.001', 'beta': '0.75', 'bias': '2', 'size': '5'})
    ew_case('concat_c', 'Concat', {0: f32(1, 8, 5, 5), 1: f32(1, 12, 5, 5), 2: f32(1, 4, 5, 5), 3: f32(1, 4, 5, 5)}, {'axis': '1'})
    ew_case('concat_rows', 'Concat', {0: f32(1, 12, 4), 1: f32(1, 20, 4)}, {'axis': '1'})
    ew_case('transpose_nhwc', 'Transpose', {0: f32(1, 12, 5, 7), 1: np.array([0, 2, 3, 1], dtype=np.int64)}, {})
    ew_case('reshape_0_m1', 'Reshape', {0: f32(1, 5, 7, 12), 1: np.array([0, -1], dtype=np.int64)}, {'special_zero': 'true'})
    ew_case('reshape_m1_k', 'Reshape', {0: f32(1, 3, 3, 64), 1: np.array([-1, 576], dtype=np.int64)}, {'special_zero': 'false'})
    ew_case('reshape_0_m1_1_4', 'Reshape', {0: f32(1, 3, 3, 12), 1: np.array([0, -1, 1, 4], dtype=np.int64)}, {'special_zero': 'true'})
    ew_case('unsqueeze', 'Unsqueeze', {0: f32(2, 36), 1: np.array([0], dtype=np.int64)}, {})

    # SSD prior branch
    ew_case('shapeof', 'ShapeOf', {0: f32(1, 12, 19, 19)}, {})
    ew_case('strided_slice', 'StridedSlice', {0: np.array([1, 12, 19, 19], dtype=np.int64), 1: np.array([2], dtype=np.int64),
                                             2: np.array([4], dtype=np.int64), 3: np.array([1], dtype=np.int64)},
            {'begin_mask': '0', 'end_mask': '1', 'new_axis_mask': '0', 'shrink_axis_mask': '0', 'ellipsis_mask': '0'})
    ew_case('priorbox', 'PriorBoxClustered', {0: np.array([3, 3], dtype=np.int64), 1: np.array([300, 300], dtype=np.int64)},
            {'clip': 'false', 'height': '30.0, 42.42640495300293, 84.85280990600586', 'width': '30.0, 84.85280990600586, 42.42640495300293',
             'offset': '0.5', 'step': '0', 'step_h': '0', 'step_w': '0', 'variance': '0.1, 0.1, 0.2, 0.2', 'img_h': '0', 'img_w': '0'})

    # DetectionOutput: 60 priors on a jittered grid, 5 classes, ~half the priors clear the threshold
    npri, ncls = 60, 5
    centers = rng.random((npri, 2)) * 0.8 + 0.1
    sizes = rng.random((npri, 2)) * 0.25 + 0.05
    pri = np.concatenate([centers - sizes / 2, centers + sizes / 2], axis=1).astype(np.float32)
    var = np.tile(np.array([0.1, 0.1, 0.2, 0.2], dtype=np.float32), (npri, 1))
    proposals = np.stack([pri.reshape(-1), var.reshape(-1)])[None].astype(np.float32)
    loc = f32(1, npri * 4, scale=1.0)
    conf = (rng.random((1, npri * ncls)).astype(np.float32))
    ew_case('detection_output', 'DetectionOutput', {0: loc, 1: conf, 2: proposals},
            {'background_label_id': '0', 'clip_after_nms': 'true', 'clip_before_nms': 'false', 'code_type': 'caffe.PriorBoxParameter.CENTER_SIZE',
             'confidence_threshold': '0.5', 'decrease_label_id': 'false', 'input_height': '1', 'input_width': '1', 'keep_top_k': '20',
             'nms_threshold': '0.30000001192092896', 'normalized': 'true', 'num_classes': str(ncls), 'share_location': 'true', 'top_k': '100',
             'variance_encoded_in_target': 'false'})

    store = {}
    meta = []
    for i, c in enumerate(cases):
        mod = importlib.import_module('op_plugins.' + c['type'])
        ins = c['ins']
        node = {'name': c['tag'], 'type': c['type'], 'version': 'opset1', 'data': dict(c['data']),
                'input': {p: _port(a) for p, a in ins.items()}}
        outs = {}
        for kt in c['kts']:
            # out port: the IR numbers it after the inputs
            op = len(ins)
            node['output'] = {op: {'precision': 'FP32', 'dims': ()}}
            if c['type'] == 'ShapeOf':
                node['output'] = {1: {'precision': 'I64', 'dims': ()}}
            r = mod.compute(node, dict(ins), kernel_type=kt)
            r = np.asarray(next(iter(r.values())))
            outs[kt] = r
            store['c{}_out_{}'.format(i, kt)] = r
        for p, a in ins.items():
            store['c{}_in{}'.format(i, p)] = a
        meta.append({'tag': c['tag'], 'type': c['type'], 'data': c['data'], 'ports': sorted(ins), 'kts': c['kts']})
        print('op case', c['tag'], {k: (v.shape, str(v.dtype)) for k, v in outs.items()})
    store['meta'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, 'ops.npz'), **store)


def golden_models(IECore):
    sys.path.insert(0, REPO)
    from tools.synth_bin import synth_input
    sys.path.remove(REPO)
    out = {}
    for model, kts in (('mnist_bn', ['special', 'numpy']), ('googlenet-v1', ['special']), ('ssd_mobilenet_v1_coco', ['special'])):
        x = synth_input(model, batch=2, seed=1)
        for img in range(2):
            for kt in kts:
                if img == 1 and kt == 'numpy':
                    continue
                final, nodes, dt = run_model(IECore, model, x[img:img + 1], kt)
                print(model, kt, 'img', img, 'sec', round(dt, 2), 'final', final.shape, float(np.abs(final).max()))
                key = '{}|{}|{}'.format(model, kt, img)
                out[key + '|final'] = np.asarray(final)
                if img == 0:
                    names = sorted(nodes)
                    out[key + '|names'] = np.array(json.dumps(names))
                    out[key + '|samples'] = np.stack([np.resize(sample(nodes[n]).astype(np.float64), 64) for n in names])
                    out[key + '|absmax'] = np.array([float(np.abs(nodes[n].astype(np.float64)).max()) if nodes[n].size else 0.0 for n in names])
    np.savez_compressed(os.path.join(HERE, 'models_e2e.npz'), **out)


if __name__ == '__main__':
    setup()
    from pyopenvino.inference_engine import IECore      # the unmodified reference
    sys.path.append(os.path.join(SCRATCH, 'pyopenvino'))
    which = sys.argv[1:] or ['mnist', 'kat', 'ops', 'models']
    if 'mnist' in which:
        golden_mnist(IECore)
    if 'kat' in which:
        golden_conv_kat()
    if 'ops' in which:
        golden_ops()
    if 'models' in which:
        golden_models(IECore)
