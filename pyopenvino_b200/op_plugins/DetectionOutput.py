"""DetectionOutput plugin -- drop-in for `op_plugins/DetectionOutput.py` (SSD post-process)."""
from .. import _cabi, common_def


def name():
    print('DetectionOutput')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    raise _cabi.B200ovError('DetectionOutput: device kernel not built yet')
