"""dreamgnn_b200 -- B200-native (sm_100a) implementation of DREAM-GNN's message-passing hot path.

Host side: Python / PyTorch modules that mirror the reference's `model.py`, `layers.py`,
`augmentation.py` and the graph-building part of `data_loader.py`. Device side: hand-written CUDA
kernels behind the C ABI in include/dreamgnn.h (libdreamgnn.so, built in-tree by
`python -m dreamgnn_b200.build`). There is no CPU fallback and no dependency on DGL.
"""
from . import _lib  # noqa: F401

__version__ = '0.1.0'
__all__ = ['graph', 'layers', 'model', 'augmentation', 'graph_build', 'ops', 'utils']
