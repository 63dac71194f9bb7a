# A script in the style of the reference's test_pyopenvino.py (imports, call sequence and result handling follow
# /root/reference/test_pyopenvino.py:7-52) with the cv2 file/display calls replaced by an .npy input.
# It must run from the repo root with NO knowledge of pyopenvino_b200: only the reference's names appear.
import sys
import time

import numpy as np
from pyopenvino.inference_engine import IECore

model = 'models/mnist'

ie = IECore()
net = ie.read_network(model + '.xml', model + '.bin')
print('inputs:', [n['name'] for n in net.inputs])
print('outputs:', [n['name'] for n in net.outputs])
output_node_name = net.outputs[0]['name']
exenet = ie.load_network(net, 'CPU')

inblob = np.load(sys.argv[1])                      # uint8 28x28 like cv2.split(cv2.imread(...))[0]
inblob = inblob.reshape(1, 1, 28, 28).astype(np.float32)

exenet.kernel_type = 'naive'

atime = 0
nitr = 2
for i in range(nitr):
    stime = time.time()
    res = exenet.infer({net.inputs[0]['name']: inblob}, verbose=False)
    etime = time.time()
    atime += etime - stime
print(atime / nitr, 'sec/inf')

m = np.argsort(res[output_node_name][0])[::-1]
result = res[output_node_name][0][m]
print('RESULT', ' '.join(str(int(v)) for v in m))
print(result)
