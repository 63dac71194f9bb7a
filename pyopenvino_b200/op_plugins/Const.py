"""Const plugin -- drop-in for `op_plugins/Const.py`.

The reference rebuilds every constant from a python tuple on every inference (`Const.py:13`, 45% of
its GoogLeNet time).  Here a float constant is uploaded to HBM once and stays resident (the packed
/ split forms the kernels need are cached on the same handle); integer constants (shape targets,
permutations, axes) stay on the host where the glue plugins read them.
"""
import numpy as np

from .. import common_def, kernels


def name():
    print('Const')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    shape = node['data']['shape']
    precision = common_def.type_convert_tbl[node['data']['element_type']]
    const = node['const']
    if precision is not np.float32:
        if 'host' not in const:
            const['host'] = np.array(const['data'], dtype=precision).reshape(shape)
        return {0: const['host']}
    if 'device' not in const:
        from .. import device as dev
        host = np.array(const['data'], dtype=precision).reshape(shape)
        arena = dev.current_arena()
        dev.set_arena(None)              # resident for the life of the network: never from the per-inference arena
        try:
            const['device'] = kernels.upload(host, keep_host=host.size <= 4096)
        finally:
            dev.set_arena(arena)
    return {0: const['device']}
