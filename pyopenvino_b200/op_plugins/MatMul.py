"""MatMul plugin -- drop-in for `op_plugins/MatMul.py`.

`transpose_a` / `transpose_b` are compared with the literal string 'true' like the reference
(`MatMul.py:12-15`).  Runs the same GEMM kernels as the 1x1 convolution (`b200ov_matmul`); a bias
Add and a ReLU that follow can be folded into the epilogue through `fused`.
"""
from .. import common_def, kernels, plugin_util


def name():
    print('MatMul')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    data = node['data']
    f = fused or {}
    return plugin_util.run_contraction(node, inputs, kernel_type, lambda math: kernels.matmul(
        inputs[0], inputs[1], transpose_a=(data['transpose_a'] == 'true'), transpose_b=(data['transpose_b'] == 'true'),
        bias=f.get('bias'), act=f.get('act'), math=math))
