"""Helpers shared by the op plugins (host side).

Mirrors the small public surface of the reference's `pyopenvino/common_def.py` that plugins rely
on -- the precision tables (`common_def.py:13-19`) and the attribute-string parsers
(`common_def.py:23-34`) -- plus the output-size rule every spatial plugin needs
(`Convolution.py:21-49`, `MaxPool.py:10-38`).  Debug printing helpers are out of scope.
"""
import math

import numpy as np

format_config = {'FP32': ['f', 4], 'FP16': ['e', 2], 'F32': ['f', 4], 'F16': ['e', 2],
                 'I64': ['q', 8], 'I32': ['i', 4], 'I16': ['h', 2], 'I8': ['b', 1], 'U8': ['B', 1]}

type_convert_tbl = {'f32': np.float32, 'f16': np.float16, 'i64': np.int64, 'i32': np.int32, 'i16': np.int16,
                    'i8': np.int8, 'u8': np.uint8, 'FP32': np.float32, 'FP16': np.float16, 'I64': np.int64}


def string_to_boolean(bool_val: str) -> bool:
    return bool_val.upper() in ('TRUE', '1')


def string_to_tuple(string: str) -> tuple:
    return tuple(int(item) for item in string.split(','))


def string_to_tuple_float(string: str) -> tuple:
    return tuple(float(item) for item in string.split(','))


def enable_escape_sequence():
    """The reference enables ANSI colours on Windows consoles here; nothing to do on Linux."""
    return True


def spatial_output_shape(input_hw, kernel_hw, strides, pads_begin, pads_end, rounding_type, auto_pad, same_is_ceil):
    """(oh, ow) of a conv / pool window sweep.

    explicit: rnd((h + pb + pe - k) / s) + 1; valid: rnd((h - k) / s) + 1;
    same_*  : ceil(h / s) for convolutions (`Convolution.py:45-47`), h for pools (`MaxPool.py:34-36`).
    """
    assert auto_pad in ('explicit', 'valid', 'same_upper', 'same_lower')
    assert rounding_type in ('floor', 'ceil')
    rnd = math.floor if rounding_type == 'floor' else math.ceil
    res = []
    for h, k, s, pb, pe in zip(input_hw, kernel_hw, strides, pads_begin, pads_end):
        if auto_pad == 'explicit':
            res.append(rnd((h + pb + pe - k) / s) + 1)
        elif auto_pad == 'valid':
            res.append(rnd((h - k) / s) + 1)
        else:
            res.append(math.ceil(h / s) if same_is_ceil else h)
    return tuple(res)


def validate_inputs(node: dict, inputs: dict):
    """The validation convention every reference plugin starts with (e.g. `Convolution.py:154-157`):
    dtype and shape of each input must equal the IR port's precision and dims."""
    for port, data in inputs.items():
        spec = node['input'][port]
        assert data.dtype == type_convert_tbl[spec['precision']], \
            '{}: port {} dtype {} != {}'.format(node.get('name'), port, data.dtype, spec['precision'])
        assert tuple(data.shape) == tuple(spec['dims']), \
            '{}: port {} shape {} != {}'.format(node.get('name'), port, tuple(data.shape), spec['dims'])


def first_output_port(node: dict):
    return next(iter(node['output']))
