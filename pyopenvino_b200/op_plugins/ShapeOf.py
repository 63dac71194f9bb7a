"""ShapeOf plugin -- drop-in for `op_plugins/ShapeOf.py`: the static dims of the input port
(`ShapeOf.py:21`), on the host (input independent, folded once at load time)."""
import numpy as np

from .. import common_def


def name():
    print('ShapeOf')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    if inputs and all(v is not None for v in inputs.values()):   # data may have been folded away; only dims matter
        common_def.validate_inputs(node, inputs)
    port = common_def.first_output_port(node)
    dims = node['input'][next(iter(node['input']))]['dims']
    return {port: np.array(dims, dtype=common_def.type_convert_tbl[node['output'][port]['precision']])}
