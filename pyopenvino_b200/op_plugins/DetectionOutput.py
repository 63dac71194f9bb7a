"""DetectionOutput plugin -- drop-in for `op_plugins/DetectionOutput.py` (SSD post-process).

Attribute parsing and defaults follow `DetectionOutput.py:267-300`; like the reference
(`DetectionOutput.py:177,185-186,228`) only share_location / normalized / two-row proposals are
supported.  The reference handles one image (`assert N == 1`); for a batch the kernel emits one block
of `keep_top_k` records per image, i.e. the (1, 1, N*keep_top_k, 7) shape of `DetectionOutput.py:232-237`
with each block equal to that image's batch-1 result.  Runs `b200ov_detection_output` (one CTA per image);
class ids, NMS decisions and record order are bit-compatible with the reference.
"""
from .. import _cabi, common_def, kernels, plugin_util


def name():
    print('DetectionOutput')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    d = node['data']
    sb = common_def.string_to_boolean
    num_classes = int(d['num_classes'])
    top_k = int(d['top_k']) if 'top_k' in d else -1
    variance_in_target = sb(d['variance_encoded_in_target']) if 'variance_encoded_in_target' in d else False
    keep_top_k = common_def.string_to_tuple(d['keep_top_k'])
    code_type = d['code_type'] if 'code_type' in d else 'caffe.PriorBoxParameter.CORNER'
    share_location = sb(d['share_location']) if 'share_location' in d else True
    nms_threshold = float(d['nms_threshold'])
    confidence_threshold = float(d['confidence_threshold']) if 'confidence_threshold' in d else 0
    clip_after_nms = sb(d['clip_after_nms']) if 'clip_after_nms' in d else False
    clip_before_nms = sb(d['clip_before_nms']) if 'clip_before_nms' in d else False
    normalized = sb(d['normalized']) if 'normalized' in d else False
    loc, conf, proposals = inputs[0], inputs[1], inputs[2]
    if not (share_location and normalized and proposals.shape[1] == 2):
        raise _cabi.B200ovError('DetectionOutput: only share_location / normalized / 2-row proposals are supported (like the reference)')
    priors = proposals.shape[2] // 4
    if keep_top_k[0] > 0:
        keep = keep_top_k[0]
    elif keep_top_k[0] == -1 and top_k > 0:
        keep = top_k * num_classes
    else:
        keep = num_classes * priors
    y = kernels.detection_output(loc, conf, proposals, num_classes, keep, code_type == 'caffe.PriorBoxParameter.CENTER_SIZE',
                                 variance_in_target, clip_before_nms, clip_after_nms, confidence_threshold, nms_threshold)
    return plugin_util.finish(node, inputs, y)
