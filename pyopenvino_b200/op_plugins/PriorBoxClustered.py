"""PriorBoxClustered plugin -- drop-in for `op_plugins/PriorBoxClustered.py`.

Input independent (depends on static shapes only), so the executor folds it once at load time on
the host.  Boxes are computed in python doubles and cast to float32 at the end, `clip` is parsed but
unused, exactly like the reference (`PriorBoxClustered.py:10-40`).
"""
import numpy as np

from .. import common_def


def name():
    print('PriorBoxClustered')


def compute(node: dict, inputs: dict = None, kernel_type: str = 'naive', debug: bool = False, fused: dict = None):
    if debug:
        print(node)
    common_def.validate_inputs(node, inputs)
    d = node['data']
    width = common_def.string_to_tuple_float(d['width']) if 'width' in d else [1.0]
    height = common_def.string_to_tuple_float(d['height']) if 'height' in d else [1.0]
    step = int(d['step']) if 'step' in d else 0.0
    step_h = int(d['step_h']) if 'step_h' in d else 0.0
    step_w = int(d['step_w']) if 'step_w' in d else 0.0
    offset = float(d['offset'])
    variance = common_def.string_to_tuple_float(d['variance']) if 'variance' in d else []
    img_h = float(d['img_h']) if 'img_h' in d else 0.0
    img_w = float(d['img_w']) if 'img_w' in d else 0.0
    grid_h, grid_w = (int(v) for v in np.asarray(inputs[0]).reshape(-1))
    image_h, image_w = (int(v) for v in np.asarray(inputs[1]).reshape(-1))
    img_h = img_h or image_h
    img_w = img_w or image_w
    step_w = step_w or step
    step_h = step_h or step
    step_w = step_w or img_w / grid_w
    step_h = step_h or img_h / grid_h
    count = grid_h * grid_w * len(width)
    boxes = np.empty((count, 4), dtype=np.float64)
    i = 0
    for gy in range(grid_h):
        cy = (gy + offset) * step_h
        for gx in range(grid_w):
            cx = (gx + offset) * step_w
            for bw, bh in zip(width, height):
                boxes[i] = ((cx - (bw / 2)) / img_w, (cy - (bh / 2)) / img_h, (cx + (bw / 2)) / img_w, (cy + (bh / 2)) / img_h)
                i += 1
    var = np.tile(np.asarray(variance, dtype=np.float64), count)
    res = np.stack([boxes.reshape(-1), var]).astype(np.float32)
    return {common_def.first_output_port(node): res}
