"""spvipes_b200 — B200-native (sm_100a) implementation of the spVIPES per-minibatch training hot path.

The arithmetic lives in hand-written CUDA kernels behind the C ABI of include/spvipes_b200.h
(libspvipes_b200.so, built in-tree by `python -m spvipes_b200.build`); this package holds the Python host side that
mirrors the reference's module interface.  There is no CPU fallback.
"""
__version__ = "0.1.0"
