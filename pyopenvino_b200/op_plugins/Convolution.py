"""Convolution plugin -- drop-in for the reference's `op_plugins/Convolution.py`.

`compute()` keeps the reference contract (`Convolution.py:149-176`): per-port validation, attribute
strings parsed on every call, output size from the explicit / valid / same_* rule
(`Convolution.py:21-49`), zero padding taken from `pads_begin` / `pads_end`, result cast to the
output port precision.  The arithmetic runs on the GPU as an implicit GEMM over NHWC feature maps
(libb200ov `b200ov_conv2d`); an Add bias and a ReLU / Clamp that follow the node can be folded into
the kernel epilogue through `fused`; `fused['pre_pool']` says that `inputs[0]` is the input of a 3x3 / stride-1
MaxPool node (`MaxPool.py:41-72`) that the executor folded into this 1x1 convolution; `fused['hl_out']` that every reader
of the result is a contraction (the result may then be left in that kernel's (hi, lo) operand form).
"""
from .. import common_def, kernels, plugin_util


def name():
    print('Convolution')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    data = node['data']
    strides = common_def.string_to_tuple(data['strides'])
    dilations = common_def.string_to_tuple(data['dilations'])
    pads_begin = common_def.string_to_tuple(data['pads_begin'])
    pads_end = common_def.string_to_tuple(data['pads_end'])
    auto_pad = data['auto_pad']
    if tuple(dilations) != (1, 1):
        # the reference 'numpy' kernel mis-handles dilation (Convolution.py:112); none of the models use it
        raise NotImplementedError('Convolution: dilations {} are not supported'.format(dilations))
    plugin_util.require_fp32_output(node)
    x, w = inputs[0], inputs[1]
    n, c, h, wd = x.shape
    kn, kc, kh, kw = w.shape
    out_hw = common_def.spatial_output_shape((h, wd), (kh, kw), strides, pads_begin, pads_end, 'floor', auto_pad, True)
    f = fused or {}
    return plugin_util.run_contraction(node, inputs, kernel_type, lambda math: kernels.conv2d(
        x, w, strides, pads_begin, out_hw, bias=f.get('bias'), act=f.get('act'), out=f.get('out'), math=math,
        pre_pool=f.get('pre_pool') is not None, hl_out=bool(f.get('hl_out'))))
