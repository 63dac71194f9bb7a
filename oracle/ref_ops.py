"""ORACLE (test infrastructure only) -- CPU restatement of the reference op-plugin kernels.

This module restates, in plain numpy, the arithmetic of yas-sim/pyopenvino's `op_plugins/<Type>.py`
for `kernel_type='numpy'` (and `'special'` for Convolution).  Every function cites the reference
file:line it follows (paths relative to the reference root).  It exists to CHECK the CUDA path:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
    may import it; the product package `pyopenvino_b200` never does (there is no CPU fallback);
  * it is pinned against the reference itself: `tests/golden/make_golden.py` runs the live
    reference plugins (imported from /root/reference in the authoring container) and the stored
    vectors under `tests/golden/` are replayed against these functions by `tests/test_oracle.py`
    (MNIST end-to-end known answer `README.md:69-72`, the `resources/node_args_6.pickle` conv
    known-answer, and per-op vectors including the reference's quirks).

Reference quirks that are reproduced on purpose: MaxPool pads with 0,
AvgPool('numpy') ignores pads and clips its window at h-1 / w-1, SoftMax normalises over the whole
tensor with no max-shift, LRN does not divide alpha by size, Multiply ignores kernel_type.
"""
import math

import numpy as np


# --------------------------------------------------------------------------------------------
# helpers (reference: pyopenvino/common_def.py:18-34)

DTYPES = {'f32': np.float32, 'f16': np.float16, 'i64': np.int64, 'i32': np.int32, 'i16': np.int16,
          'i8': np.int8, 'u8': np.uint8, 'FP32': np.float32, 'FP16': np.float16, 'I64': np.int64}


def ints(text):
    """'1, 1' -> (1, 1)  (common_def.py:28-30)"""
    return tuple(int(t) for t in text.split(','))


def floats(text):
    """'0.1, 0.2' -> (0.1, 0.2)  (common_def.py:32-34)"""
    return tuple(float(t) for t in text.split(','))


def truthy(text):
    """'true' / '1' -> True  (common_def.py:23-26)"""
    return text.upper() in ('TRUE', '1')


def out_hw(hw, khw, strides, pads_begin, pads_end, rounding, auto_pad, same_is_ceil):
    """Output feature-map size.

    Convolution.py:21-49 / GroupConvolution.py:22-50 (`same_*` -> ceil(h/s), same_is_ceil=True) and
    MaxPool.py:10-38 / AvgPool.py:10-38 (`same_*` -> h, same_is_ceil=False).  Uses true division
    followed by floor/ceil exactly like the reference.
    """
    assert auto_pad in ('explicit', 'valid', 'same_upper', 'same_lower')
    assert rounding in ('floor', 'ceil')
    rnd = math.floor if rounding == 'floor' else math.ceil
    out = []
    for i in range(2):
        h, k, s, pb, pe = hw[i], khw[i], strides[i], pads_begin[i], pads_end[i]
        if auto_pad == 'explicit':
            o = rnd((h + pb + pe - k) / s) + 1
        elif auto_pad == 'valid':
            o = rnd((h - k) / s) + 1
        else:
            o = math.ceil(h / s) if same_is_ceil else h
        out.append(o)
    return tuple(out)


# --------------------------------------------------------------------------------------------
# Convolution

def conv_special(x, w, strides, pads_begin, pads_end, auto_pad):
    """im2col + np.dot ('special' kernel), Convolution.py:57-87.  Handles N > 1.

    The column index is ordered (c, ky, kx) (Convolution.py:66-69) and the GEMM is
    `col[N*OH*OW, C*kh*kw] . W.reshape(K, -1).T` (Convolution.py:83-84) in float32 (OpenBLAS sgemm).
    """
    n, c, h, wd = x.shape
    kn, kc, kh, kw = w.shape
    sh, sw = strides
    oh, ow = out_hw((h, wd), (kh, kw), strides, pads_begin, pads_end, 'floor', auto_pad, True)
    img = np.pad(x, [(0, 0), (0, 0), (pads_begin[0], pads_end[0]), (pads_begin[1], pads_end[1])], 'constant')
    col = np.zeros((n, c, kh, kw, oh, ow), dtype=np.float32)
    for ky in range(kh):
        for kx in range(kw):
            col[:, :, ky, kx, :, :] = img[:, :, ky:ky + sh * oh:sh, kx:kx + sw * ow:sw]
    col = col.transpose(0, 4, 5, 1, 2, 3).reshape(n * oh * ow, -1)
    out = np.dot(col, w.reshape(kn, -1).T)
    return out.reshape(n, oh, ow, -1).transpose(0, 3, 1, 2)


def conv_numpy(x, w, strides, dilations, pads_begin, pads_end, auto_pad):
    """'numpy' kernel, Convolution.py:91-114: per (filter, oy, ox) `np.sum(patch * kernel[f])`.

    Only image 0 is computed (Convolution.py:112); dilation is used as a slice step (:112).
    Slow (python loop per output element) -- used on small shapes only.
    """
    n, c, h, wd = x.shape
    kn, kc, kh, kw = w.shape
    sh, sw = strides
    dh, dw = dilations
    oh, ow = out_hw((h, wd), (kh, kw), strides, pads_begin, pads_end, 'floor', auto_pad, True)
    xp = np.pad(x, [(0, 0), (0, 0), (pads_begin[0], pads_end[0]), (pads_begin[1], pads_end[1])], 'constant')
    out = np.zeros((n, kn, oh, ow), dtype=np.float32)
    for f in range(kn):
        wf = w[f]
        for oy in range(oh):
            for ox in range(ow):
                patch = xp[0, :, oy * sh:oy * sh + kh:dh, ox * sw:ox * sw + kw:dw]
                out[0, f, oy, ox] = np.sum(patch * wf)
    return out


def convolution(node_data, x, w, kernel_type='numpy', out_dtype=np.float32):
    """Convolution.compute, Convolution.py:149-176 (attribute parsing + dispatch + output cast)."""
    strides = ints(node_data['strides'])
    dil = ints(node_data['dilations'])
    pb = ints(node_data['pads_begin'])
    pe = ints(node_data['pads_end'])
    ap = node_data['auto_pad']
    if kernel_type == 'special':
        res = conv_special(x, w, strides, pb, pe, ap)
    else:
        res = conv_numpy(x, w, strides, dil, pb, pe, ap)
    return res.astype(out_dtype)


# --------------------------------------------------------------------------------------------
# GroupConvolution (depthwise only, like the reference)

def groupconv_numpy_loops(x, w, strides, pads_begin, pads_end, auto_pad):
    """Literal restatement of GroupConvolution.py:53-79 (image 0 only; index math `g*ci+g`)."""
    n, c, h, wd = x.shape
    grp, cho, chi, kh, kw = w.shape
    sh, sw = strides
    oh, ow = out_hw((h, wd), (kh, kw), strides, pads_begin, pads_end, 'floor', auto_pad, True)
    xp = np.pad(x, [(0, 0), (0, 0), (pads_begin[0], pads_end[0]), (pads_begin[1], pads_end[1])], 'constant')
    out = np.zeros((n, grp * cho, oh, ow), dtype=x.dtype)
    for ci in range(chi):
        for g in range(grp):
            for co in range(cho):
                flt = w[g, co, ci, :, :]
                for oy in range(oh):
                    for ox in range(ow):
                        patch = xp[0, g * ci + g, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw]
                        out[0, g * co + g, oy, ox] = np.sum(patch * flt)
    return out


def pairwise_sum_terms(terms):
    """Sum a short list of equally-shaped arrays in the order numpy's float32 pairwise reduction
    uses for one contiguous run of len(terms) elements (numpy `pairwise_sum`: n < 8 -> sequential;
    8 <= n <= 128 -> eight strided partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
    then the remainder added sequentially)."""
    n = len(terms)
    if n < 8:
        acc = terms[0] * np.float32(1)   # copy
        # numpy starts from -0.0 / 0.0 and adds; adding the first element to +-0 is exact
        for t in terms[1:]:
            acc = acc + t
        return acc
    assert n <= 128
    r = [terms[i] for i in range(8)]
    i = 8
    while i + 8 <= n:
        for j in range(8):
            r[j] = r[j] + terms[i + j]
        i += 8
    acc = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        acc = acc + terms[i]
        i += 1
    return acc


def groupconv_numpy(x, w, strides, pads_begin, pads_end, auto_pad):
    """Vectorised depthwise form of GroupConvolution.py:53-79 for the case the reference is correct
    for (C_out/G = C_in/G = 1), all images.  The kh*kw products are formed in float32 and summed in
    numpy's pairwise order, so the result is bit-identical to `np.sum(patch*flt)` per element
    (checked against the literal loops and against the live reference in tests)."""
    n, c, h, wd = x.shape
    grp, cho, chi, kh, kw = w.shape
    assert cho == 1 and chi == 1 and grp == c, 'reference GroupConvolution is only right for depthwise'
    sh, sw = strides
    oh, ow = out_hw((h, wd), (kh, kw), strides, pads_begin, pads_end, 'floor', auto_pad, True)
    xp = np.pad(x, [(0, 0), (0, 0), (pads_begin[0], pads_end[0]), (pads_begin[1], pads_end[1])], 'constant')
    terms = []
    for ky in range(kh):
        for kx in range(kw):
            win = xp[:, :, ky:ky + sh * (oh - 1) + 1:sh, kx:kx + sw * (ow - 1) + 1:sw]
            terms.append(win * w[:, 0, 0, ky, kx].reshape(1, c, 1, 1))
    return pairwise_sum_terms(terms).astype(x.dtype)


def group_convolution(node_data, x, w, kernel_type='numpy'):
    """GroupConvolution.compute, GroupConvolution.py:114-137 (no output cast)."""
    return groupconv_numpy(x, w, ints(node_data['strides']), ints(node_data['pads_begin']),
                           ints(node_data['pads_end']), node_data['auto_pad'])


# --------------------------------------------------------------------------------------------
# MatMul

def matmul(node_data, a, b):
    """MatMul.py:9-17: optional transposes selected by the literal strings 'true', then np.matmul."""
    if node_data['transpose_a'] == 'true':
        a = a.T
    if node_data['transpose_b'] == 'true':
        b = b.T
    return np.matmul(a, b)


# --------------------------------------------------------------------------------------------
# Pooling

def maxpool(node_data, x):
    """MaxPool.py:41-72 ('numpy'): zero padding takes part in the max (:53), windows that overhang
    the padded tensor (ceil mode) are clipped (:69).  Handles N > 1."""
    strides = ints(node_data['strides'])
    pb = ints(node_data['pads_begin'])
    pe = ints(node_data['pads_end'])
    kh, kw = ints(node_data['kernel'])
    n, c, h, wd = x.shape
    oh, ow = out_hw((h, wd), (kh, kw), strides, pb, pe, node_data['rounding_type'], node_data['auto_pad'], False)
    xp = np.pad(x, [(0, 0), (0, 0), (pb[0], pe[0]), (pb[1], pe[1])], 'constant')
    hp, wp = xp.shape[2:]
    out = np.zeros((n, c, oh, ow), dtype=x.dtype)
    for oy in range(oh):
        for ox in range(ow):
            win = xp[:, :, oy * strides[0]:min(hp, oy * strides[0] + kh), ox * strides[1]:min(wp, ox * strides[1] + kw)]
            out[:, :, oy, ox] = np.max(win, axis=(2, 3))
    return out


def avgpool(node_data, x):
    """AvgPool.py:41-59 ('numpy'): no padding is applied and the window is clipped at h-1 / w-1
    (:56), so the 7x7 GoogLeNet pool averages x[:, :, 0:6, 0:6]."""
    sh, sw = ints(node_data['strides'])
    pb = ints(node_data['pads_begin'])
    pe = ints(node_data['pads_end'])
    kh, kw = ints(node_data['kernel'])
    n, c, h, wd = x.shape
    oh, ow = out_hw((h, wd), (kh, kw), (sh, sw), pb, pe, node_data['rounding_type'], node_data['auto_pad'], False)
    out = np.zeros((n, c, oh, ow), dtype=x.dtype)
    for b in range(n):
        for ch in range(c):
            for oy in range(oh):
                for ox in range(ow):
                    win = x[b, ch, oy * sh:min(h - 1, oy * sh + kh), ox * sw:min(wd - 1, ox * sw + kw)]
                    out[b, ch, oy, ox] = np.average(win)
    return out


# --------------------------------------------------------------------------------------------
# Elementwise tail

def add(a, b):
    """Add.py:9-14: port 1 is broadcast to port 0's shape."""
    return a + np.broadcast_to(b, a.shape)


def multiply(a, b):
    """Multiply.py:9-17: the operand with fewer elements is broadcast to the other's shape.
    (`compute` always returns this result whatever kernel_type says, Multiply.py:46-62.)"""
    if a.size > b.size:
        b = np.broadcast_to(b, a.shape)
    else:
        a = np.broadcast_to(a, b.shape)
    return a * b


def relu(x):
    """ReLU.py:9-12: np.where(x<0, 0, x) (NaN and -0.0 pass through)."""
    return np.where(x < 0, 0, x)


def clamp(node_data, x):
    """Clamp.py:9-12,45-46."""
    return np.clip(x, float(node_data['min']), float(node_data['max']))


def softmax(x):
    """SoftMax.py:10-14: exp(x)/sum(exp(x)) over ALL elements, `axis` ignored, no max-shift."""
    return np.exp(x) / np.sum(np.exp(x))


def sigmoid(x):
    """Sigmoid.py:10-13."""
    return 1 / (1 + np.exp(-x))


def lrn(node_data, x):
    """LRN.py:10-22: den_c = (bias + alpha * sum_{c' in [c-size//2, c+size//2]} x^2)^beta, out = x/den.
    alpha is NOT divided by size; the axes input is ignored."""
    alpha = float(node_data['alpha'])
    beta = float(node_data['beta'])
    bias = float(node_data['bias'])
    size = int(node_data['size'])
    n, c, h, w = x.shape
    sq = x ** 2
    den = np.zeros_like(x)
    for ch in range(c):
        den[:, ch, :, :] = (bias + alpha * np.sum(sq[:, max(0, ch - size // 2):min(c, ch + size // 2 + 1), :, :], axis=1)) ** beta
    return x / den


# --------------------------------------------------------------------------------------------
# Layout glue

def concat(node_data, arrays):
    """Concat.py:9-13 (inputs in port order)."""
    return np.concatenate(list(arrays), axis=int(node_data['axis']))


def transpose(x, perm):
    """Transpose.py:9-13."""
    return x.transpose(perm)


def reshape(x, target):
    """Reshape.py:14-44: 0 copies the input dim (left aligned), -1 is inferred; `special_zero` is
    ignored."""
    size = x.size
    dims = []
    deferred = -1
    zero_ok = True
    for idx, d in enumerate(target):
        d = int(d)
        if d == 0:
            assert zero_ok
            d0 = x.shape[idx]
            assert size % d0 == 0
            dims.append(int(d0))
            size //= d0
        else:
            zero_ok = False
            if d == -1:
                assert deferred == -1
                deferred = idx
                dims.append(-1)
            else:
                assert size % d == 0
                dims.append(d)
                size //= d
    if deferred != -1:
        dims[deferred] = int(size)
    return x.reshape(dims)


def unsqueeze(x, axes):
    """Unsqueeze.py:9-14."""
    return np.expand_dims(x, list(int(a) for a in axes))


def shape_of(in_dims, out_dtype=np.int64):
    """ShapeOf.py:9-25: the static port dims."""
    return np.array(in_dims, dtype=out_dtype)


def strided_slice(x, begin, end, stride):
    """StridedSlice.py:8-24: plain python slicing `x[b:e:s, ...]`, every mask ignored."""
    idx = tuple(slice(int(b), int(e), int(s)) for b, e, s in zip(begin, end, stride))
    return x[idx]


def prior_box_clustered(node_data, grid_hw, image_hw):
    """PriorBoxClustered.py:10-40: python-double arithmetic, cast to float32 at the end; `clip` unused."""
    d = node_data
    width = floats(d['width']) if 'width' in d else [1.0]
    height = floats(d['height']) if 'height' in d else [1.0]
    step = int(d['step']) if 'step' in d else 0.0
    step_h = int(d['step_h']) if 'step_h' in d else 0.0
    step_w = int(d['step_w']) if 'step_w' in d else 0.0
    offset = float(d['offset'])
    variance = floats(d['variance']) if 'variance' in d else []
    img_h = float(d['img_h']) if 'img_h' in d else 0.0
    img_w = float(d['img_w']) if 'img_w' in d else 0.0
    grid_h, grid_w = [int(v) for v in grid_hw]
    image_h, image_w = [int(v) for v in image_hw]
    img_h = image_h if img_h == 0 else img_h
    img_w = image_w if img_w == 0 else img_w
    step_w = step if step_w == 0 else step_w
    step_h = step if step_h == 0 else step_h
    step_w = (img_w / grid_w) if step_w == 0 else step_w
    step_h = (img_h / grid_h) if step_h == 0 else step_h
    boxes = []
    for gy in range(grid_h):
        for gx in range(grid_w):
            cx = (gx + offset) * step_w
            cy = (gy + offset) * step_h
            for bw, bh in zip(width, height):
                boxes.extend([(cx - (bw / 2)) / img_w, (cy - (bh / 2)) / img_h,
                              (cx + (bw / 2)) / img_w, (cy + (bh / 2)) / img_h])
    var = list(np.tile(variance, grid_h * grid_w * len(width)))
    return np.array([boxes, var], dtype=np.float32)


# --------------------------------------------------------------------------------------------
# DetectionOutput (SSD post-process), DetectionOutput.py:162-260

def _iou(a, b):
    """DetectionOutput.py:12-34, float32 numpy-scalar arithmetic."""
    ax0, ay0, ax1, ay1 = a
    bx0, by0, bx1, by1 = b
    area_a = (ax1 - ax0) * (ay1 - ay0)
    area_b = (bx1 - bx0) * (by1 - by0)
    iw = min(ax1, bx1) - max(ax0, bx0)
    ih = min(ay1, by1) - max(ay0, by0)
    if iw < 0 or ih < 0:
        return 0.0
    inter = iw * ih
    with np.errstate(divide='ignore', invalid='ignore'):
        return inter / (area_a + area_b - inter)


def detection_output(node_data, loc, conf, proposals):
    """DetectionOutput.py:162-260 for the configuration the reference supports
    (share_location, normalized, CENTER_SIZE or CORNER, N == 1)."""
    d = node_data
    num_classes = int(d['num_classes'])
    keep_top_k = ints(d['keep_top_k'])
    top_k = int(d['top_k']) if 'top_k' in d else -1
    var_in_target = truthy(d['variance_encoded_in_target']) if 'variance_encoded_in_target' in d else False
    code_type = d['code_type'] if 'code_type' in d else 'caffe.PriorBoxParameter.CORNER'
    nms_thr = float(d['nms_threshold'])
    conf_thr = float(d['confidence_threshold']) if 'confidence_threshold' in d else 0
    clip_after = truthy(d['clip_after_nms']) if 'clip_after_nms' in d else False
    clip_before = truthy(d['clip_before_nms']) if 'clip_before_nms' in d else False
    normalized = truthy(d['normalized']) if 'normalized' in d else False
    assert normalized and proposals.shape[1] == 2 and loc.shape[0] == 1
    npri = proposals.shape[2] // 4
    loc_ = loc.reshape(npri, 4)
    conf_ = conf.reshape(npri, num_classes)
    pri = proposals[:, 0, :].reshape(npri, 4)
    var = proposals[:, 1, :].reshape(npri, 4)

    # top-1 class per prior (:196-201) then threshold + background rejection (:69-94)
    kept = []
    for p in range(npri):
        order = np.argsort(conf_[p])[::-1]
        cls, score = order[0], conf_[p, order[0]]
        if score > conf_thr and cls != 0:
            kept.append((p, np.float32(cls), score))
    m = len(kept)
    boxes = np.zeros((m, 4), dtype=np.float32)
    score = np.zeros((m,), dtype=np.float32)
    label = np.zeros((m,), dtype=np.float32)
    for i, (p, cls, sc) in enumerate(kept):
        score[i], label[i] = sc, cls
        pxmin, pymin, pxmax, pymax = pri[p]
        l0, l1, l2, l3 = loc_[p]
        if code_type == 'caffe.PriorBoxParameter.CORNER':
            if var_in_target:
                nb = (pxmin + l0, pymin + l1, pxmax + l2, pymax + l3)
            else:
                nb = (pxmin + var[p, 0] * l0, pymin + var[p, 1] * l1, pxmax + var[p, 2] * l2, pymax + var[p, 3] * l3)
        else:   # CENTER_SIZE (:125-144)
            pw, ph = pxmax - pxmin, pymax - pymin
            pcx, pcy = (pxmin + pxmax) / 2, (pymin + pymax) / 2
            if var_in_target:
                cx, cy = l0 * pw + pcx, l1 * ph + pcy
                bw, bh = math.exp(l2) * pw, math.exp(l3) * ph
            else:
                cx, cy = var[p, 0] * l0 * pw + pcx, var[p, 1] * l1 * ph + pcy
                bw, bh = math.exp(var[p, 2] * l2) * pw, math.exp(var[p, 3] * l3) * ph
            nb = (cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2)
        boxes[i] = nb

    def clip(b):
        for r in b:
            for j in range(4):
                r[j] = max(0, min(1, r[j]))
    if clip_before:
        clip(boxes)
    # class-agnostic all-pairs NMS (:38-63)
    keep = [True] * m
    for i in range(0, m - 1):
        for j in range(i + 1, m):
            if _iou(boxes[i], boxes[j]) > nms_thr:
                if score[i] < score[j]:
                    keep[i] = False
                else:
                    keep[j] = False
    sel = [i for i in range(m) if keep[i]]
    boxes, score, label = boxes[sel].reshape(-1, 4), score[sel], label[sel]
    if clip_after:
        clip(boxes)
    n_img = 1
    shape = (1, 1, n_img * num_classes * npri, 7)
    if keep_top_k[0] > 0:
        shape = (1, 1, n_img * keep_top_k[0], 7)
    elif keep_top_k[0] == -1 and top_k > 0:
        shape = (1, 1, n_img * top_k * num_classes, 7)
    res = np.zeros(shape, dtype=np.float32)
    order = np.argsort(score)[::-1]
    nb = len(boxes)
    for r in range(min(shape[2], nb)):
        i = order[r]
        res[0, 0, r, :] = np.array([r, label[i], score[i], boxes[i][0], boxes[i][1], boxes[i][2], boxes[i][3]], dtype=np.float32)
    if nb < shape[2]:
        res[0, 0, nb, :] = np.array([-1, 0, 0, 0, 0, 0, 0], dtype=np.float32)
    return res
