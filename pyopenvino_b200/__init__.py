"""pyopenvino_b200 -- B200-native (sm_100a) implementation of pyOpenVINO's op-plugin hot path behind
the reference's IECore / read_network / load_network / infer API.

    from pyopenvino_b200.inference_engine import IECore
"""
__version__ = '0.1.0'
