"""Synthetic `.bin` generator for the IR graphs whose trained weights are not shipped.

Three of the four reference models have no `.bin` (reference `.MISSING_LARGE_BLOBS:1-3`), so every
benchmark / parity run on mnist_bn, googlenet-v1 and ssd_mobilenet_v1_coco uses weights produced here.
The recipe follows SURVEY.md Appendix C.2 / section 8(d):

* f32 constants are drawn according to the *consumer* of the Const node (found through `<edges>`):
  conv / matmul weights ~ N(0, 2/fan_in), depthwise ~ N(0, 2/9), Add operands ~ 0.05*N(0,1),
  Multiply operands (folded BN scale) ~ 1 + 0.1*N(0,1), scalar pre-processing constants 1/127.5 and -1.
* i64 constants are derived from the XML port dims (Reshape targets in batch-agnostic `[0,-1,...]`
  form, Transpose permutations, Unsqueeze / StridedSlice / LRN axes).  Model Optimizer de-duplicated
  identical i64 blobs, so several Const nodes may share one offset; the derived values are checked
  to agree.

Everything is seeded (`numpy.random.default_rng(seed)`), so the same bytes are produced here, in the
tests and on the GPU box.  This file is data tooling: it is not part of the product path.
"""
import os
import xml.etree.ElementTree as et

import numpy as np

_NP = {'f32': np.float32, 'i64': np.int64, 'i32': np.int32, 'f16': np.float16}


def _parse(xml_path):
    root = et.parse(xml_path).getroot()
    layers = {}
    for layer in root.findall('./layers/layer'):
        lid = int(layer.attrib['id'])
        d = {'type': layer.attrib['type'], 'name': layer.attrib['name'], 'in': {}, 'out': {}}
        data = layer.find('data')
        d['data'] = dict(data.attrib) if data is not None else {}
        for tag, key in (('input', 'in'), ('output', 'out')):
            node = layer.find(tag)
            if node is not None:
                for port in node.findall('port'):
                    d[key][int(port.attrib['id'])] = tuple(int(x.text) for x in port.findall('dim'))
        layers[lid] = d
    edges = [(int(e.attrib['from-layer']), int(e.attrib['from-port']),
              int(e.attrib['to-layer']), int(e.attrib['to-port'])) for e in root.findall('./edges/edge')]
    return layers, edges


def _reshape_target(in_dims, out_dims):
    if len(out_dims) >= 2 and len(in_dims) >= 1 and out_dims[0] == in_dims[0]:
        # batch-agnostic form: keep dim 0, infer dim 1
        return [0, -1] + list(out_dims[2:])
    return list(out_dims)


def _transpose_perm(in_dims, out_dims):
    for perm in ([0, 2, 3, 1], [0, 3, 1, 2], [1, 0], [0, 1, 2, 3]):
        if len(perm) == len(in_dims) and tuple(in_dims[p] for p in perm) == tuple(out_dims):
            return perm
    raise ValueError('cannot derive Transpose permutation {} -> {}'.format(in_dims, out_dims))


def synth_blob(xml_path, seed=0, cls_bias=-2.2):
    """Return the bytes of a synthetic `.bin` for `xml_path`."""
    layers, edges = _parse(xml_path)
    consumers = {}
    for fl, fp, tl, tp in edges:
        consumers.setdefault(fl, []).append((tl, tp))
    total = 0
    for lid, l in layers.items():
        if l['type'] == 'Const':
            total = max(total, int(l['data']['offset']) + int(l['data']['size']))
    blob = bytearray(total)
    written = {}
    rng = np.random.default_rng(seed)
    for lid in sorted(layers):
        l = layers[lid]
        if l['type'] != 'Const':
            continue
        off, size = int(l['data']['offset']), int(l['data']['size'])
        et_ = l['data']['element_type']
        shape = tuple(int(x) for x in l['data']['shape'].split(',')) if l['data']['shape'].strip() else ()
        count = int(np.prod(shape)) if len(shape) else 1
        cons = consumers.get(lid, [])
        assert len(cons) >= 1, 'Const {} has no consumer'.format(lid)
        tl, tp = cons[0]
        c = layers[tl]
        ctype = c['type']
        if et_ == 'f32':
            if ctype == 'Convolution':
                fan_in = int(np.prod(shape[1:]))
                val = rng.standard_normal(count) * np.sqrt(2.0 / fan_in)
            elif ctype == 'GroupConvolution':
                fan_in = int(np.prod(shape[2:]))
                val = rng.standard_normal(count) * np.sqrt(2.0 / fan_in)
            elif ctype == 'MatMul':
                fan_in = shape[-1] if c['data'].get('transpose_b', 'false') == 'true' else shape[0]
                val = rng.standard_normal(count) * np.sqrt(2.0 / fan_in)
            elif ctype == 'Multiply':
                val = np.full(count, 1.0 / 127.5) if count == 1 else 1.0 + 0.1 * rng.standard_normal(count)
            elif ctype == 'Add':
                if count == 1:
                    val = np.full(count, -1.0)
                else:
                    val = 0.05 * rng.standard_normal(count)
                    # SSD class heads: push logits down so only a handful of priors clear the
                    # DetectionOutput confidence threshold (keeps the all-pairs NMS oracle fast).
                    if 'ClassPredictor' in l['name'] or 'ClassPredictor' in c['name']:
                        val = val + cls_bias
            else:
                val = 0.05 * rng.standard_normal(count)
            data = val.astype(np.float32)
        elif et_ == 'i64':
            if ctype == 'Reshape':
                v = _reshape_target(c['in'][0], next(iter(c['out'].values())))
            elif ctype == 'Transpose':
                v = _transpose_perm(c['in'][0], next(iter(c['out'].values())))
            elif ctype == 'Unsqueeze':
                v = [0]
            elif ctype == 'StridedSlice':
                v = {1: [2], 2: [4], 3: [1]}[tp]
            elif ctype == 'LRN':
                v = [1]
            else:
                raise ValueError('no i64 rule for consumer type ' + ctype)
            assert len(v) == count, (l['name'], v, shape)
            data = np.asarray(v, dtype=np.int64)
        else:
            raise ValueError('unsupported const element type ' + et_)
        raw = data.tobytes()
        assert len(raw) == size, (l['name'], len(raw), size)
        if off in written:
            # de-duplicated blob shared by several Const nodes: keep the first, check i64 agreement
            if et_ == 'i64':
                assert bytes(blob[off:off + size]) == raw, 'shared i64 blob mismatch at {}'.format(off)
            continue
        written[off] = lid
        blob[off:off + size] = raw
    return bytes(blob)


def ensure_model(model, dst_dir, src_dir=None, seed=0):
    """Make sure `<dst_dir>/<model>.xml` and `.bin` exist; returns the xml path.

    `mnist` ships real weights (`models/mnist.bin`); the other models get a synthetic `.bin`.
    """
    if src_dir is None:
        src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'models')
    os.makedirs(dst_dir, exist_ok=True)
    src_xml = os.path.join(src_dir, model + '.xml')
    dst_xml = os.path.join(dst_dir, model + '.xml')
    dst_bin = os.path.join(dst_dir, model + '.bin')
    if os.path.abspath(src_xml) != os.path.abspath(dst_xml) and not os.path.isfile(dst_xml):
        with open(src_xml, 'rb') as f, open(dst_xml, 'wb') as g:
            g.write(f.read())
    if not os.path.isfile(dst_bin):
        src_bin = os.path.join(src_dir, model + '.bin')
        if os.path.isfile(src_bin):
            with open(src_bin, 'rb') as f, open(dst_bin, 'wb') as g:
                g.write(f.read())
        else:
            tmp = dst_bin + '.tmp{}'.format(os.getpid())
            with open(tmp, 'wb') as g:
                g.write(synth_blob(src_xml, seed=seed))
            os.replace(tmp, dst_bin)
    return dst_xml


def synth_input(model, batch=1, seed=1):
    """Seeded synthetic input batch for a model (NCHW float32), activation scale kept O(1)."""
    shapes = {'mnist': (1, 28, 28), 'mnist_bn': (1, 28, 28), 'googlenet-v1': (3, 224, 224),
              'ssd_mobilenet_v1_coco': (3, 300, 300)}
    rng = np.random.default_rng(seed)
    x = rng.random((batch,) + shapes[model], dtype=np.float32)
    if model == 'ssd_mobilenet_v1_coco':
        x = x * np.float32(255.0)     # pre-processing in the graph maps [0,255) -> [-1,1)
    return x


if __name__ == '__main__':
    import sys
    out = sys.argv[1] if len(sys.argv) > 1 else '/tmp/b200ov_models'
    for m in ('mnist', 'mnist_bn', 'googlenet-v1', 'ssd_mobilenet_v1_coco'):
        print(ensure_model(m, out))
