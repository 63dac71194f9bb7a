"""GroupConvolution plugin (depthwise) -- drop-in for `op_plugins/GroupConvolution.py`.

Like the reference (`GroupConvolution.py:53-79`, index math `gp*ci+gp`) only the depthwise case
C_out/G = C_in/G = 1 is meaningful; anything else is rejected loudly.  Runs `b200ov_dwconv2d`.
`kernel_type='exact'` selects the kernel whose pre-bias result is bit-identical to the reference's
`np.sum(patch*flt)`; every other kernel_type runs the packed-FMA 3x3 kernel (FP32 tolerance class).
"""
from .. import common_def, kernels, plugin_util


def name():
    print('GroupConvolution')


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
        raise NotImplementedError('GroupConvolution: dilations {} are not supported'.format(dilations))
    x, w = inputs[0], inputs[1]
    n, c, h, wd = x.shape
    grp, ch_o, ch_i, kh, kw = w.shape
    out_hw = common_def.spatial_output_shape((h, wd), (kh, kw), strides, pads_begin, pads_end, 'floor', auto_pad, True)
    f = fused or {}
    y = kernels.dwconv2d(x, w, strides, pads_begin, out_hw, bias=f.get('bias'), act=f.get('act'), out=f.get('out'),
                         exact=(kernel_type == 'exact'), hl_out=bool(f.get('hl_out')))
    return plugin_util.finish(node, inputs, y)
