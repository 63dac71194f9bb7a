# In the style of the reference's test_node_sample.py (/root/reference/test_node_sample.py:1-16): a pickled
# (node, inputs) pair is run through one operator plugin imported as top-level `op_plugins.<Type>`.
import pickle
import sys
sys.path.append('pyopenvino')

with open(sys.argv[1], 'rb') as f:
    node, inputs = pickle.load(file=f)

import op_plugins.MaxPool as op

print('Node name={} ,type={} is running.'.format(node['name'], node['type']))
res = op.compute(node, inputs)
import numpy as np
np.save(sys.argv[2], res[next(iter(res))])
