"""sparsernns_b200 - B200-native (sm_100a) drop-in for the S5 / NDNS forward path of stevenabreu7/SparseRNNs.

The hot path lives in ``csrc/`` (hand-written CUDA behind the C ABI of ``include/s5b200.h``); the Python
modules mirror the reference's call structure (model / layer / SSM / fixed-point model) on top of it.
"""

__version__ = "0.1.0"
