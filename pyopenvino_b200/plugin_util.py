"""Conventions shared by the op plugins of this package.

Contract kept from the reference (`README.md:130`, e.g. `Convolution.py:149-176`):
`compute(node, inputs, kernel_type='naive', debug=False) -> {out_port: array}` with the per-port
dtype / shape validation first.  Extensions, all optional and ignored by reference-style callers:

  * arrays may be `DeviceArray`s (HBM resident).  If no input is a DeviceArray the plugin uploads
    the host arrays, runs the CUDA kernel and returns host ndarrays, so a script written for the
    reference (`test_node_sample.py`) keeps working; otherwise results stay on the device.
  * `fused=` carries what the executor folded into this node: {'bias', 'act', 'scale', 'shift',
    'out'} (see inference_engine.FusionPlan).
  * `kernel_type`: the reference's 'naive' / 'numpy' / 'special' all select the CUDA kernel (there is
    no CPU path); 'fp32', 'tf32x3', 'tf32' pin the arithmetic of the dense contractions.
"""
from . import _cabi, common_def
from .device import is_device

_MATH = {'fp32': _cabi.MATH_FP32, 'tf32x3': _cabi.MATH_TF32X3, 'tf32': _cabi.MATH_TF32, 'f16x2': _cabi.MATH_F16X2, 'safe': _cabi.MATH_SAFE}


def math_mode(kernel_type):
    return _MATH.get(kernel_type)        # None -> library default (AUTO)


def host_in_host_out(inputs):
    return not any(is_device(v) for v in inputs.values())


def finish(node, inputs, result):
    port = common_def.first_output_port(node)
    if host_in_host_out(inputs) and is_device(result):
        result = result.numpy()
    return {port: result}


def run_contraction(node, inputs, kernel_type, run):
    """Convolution / MatMul: `run(math)` launches the kernel.  In host-in / host-out mode with the default
    arithmetic the f16x2 range flag is checked here (the executor checks it once per inference instead):
    a non-finite output repeats the node with the FP32-range kernels."""
    math = math_mode(kernel_type)
    if not (host_in_host_out(inputs) and math is None):
        return finish(node, inputs, run(math))
    from . import device as dev, kernels
    kernels.status_reset()
    y = run(None)
    kernels.status_fetch()
    dev.synchronize()
    if kernels.status_value() != 0:
        y = run(_cabi.MATH_SAFE)
    return finish(node, inputs, y)


def require_fp32_output(node):
    port = common_def.first_output_port(node)
    prec = node['output'][port]['precision']
    if common_def.type_convert_tbl[prec] is not common_def.type_convert_tbl['FP32']:
        raise _cabi.B200ovError('{}: output precision {} is not supported (FP32 only)'.format(node.get('name'), prec))
