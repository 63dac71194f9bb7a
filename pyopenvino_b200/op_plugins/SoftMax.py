"""SoftMax plugin -- drop-in for `op_plugins/SoftMax.py`.

The reference normalises over the whole tensor and ignores `axis` (`SoftMax.py:10-14`); with the
reference's batch-1 tensors that is one row per image, which is how the batched kernel defines it
(`b200ov_softmax`, rows = dim 0).  The kernel subtracts the row max first, so it stays finite where
the reference overflows to NaN; within tolerance they agree.
"""
from .. import common_def, kernels, plugin_util


def name():
    print('SoftMax')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    y = kernels.softmax_rows(inputs[0])
    return plugin_util.finish(node, inputs, y)
