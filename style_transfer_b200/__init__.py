"""B200-native engine for the per-tile hot path of crowsonkb/style_transfer.

Host side (this package) mirrors the reference's operator interface; all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``libstyle_b200.so``
(include/style_b200.h).  There is no CPU fallback.
"""

from . import _lib, netdesc, weights                       # noqa: F401
from ._lib import StError                                   # noqa: F401

__all__ = ['netdesc', 'weights', 'StError', 'TileEngine', 'StyleTransfer', 'AdamOptimizer',
           'LBFGSOptimizer']


def __getattr__(name):
    # torch is imported lazily so that tooling which only needs the graph/flag code stays light
    if name in ('TileEngine', 'ContentData', 'StyleData'):
        from . import engine
        return getattr(engine, name)
    if name in ('StyleTransfer', 'parse_weights', 'scale_ladder'):
        from . import transfer
        return getattr(transfer, name)
    if name in ('AdamOptimizer', 'LBFGSOptimizer'):
        from . import optimizers
        return getattr(optimizers, name)
    raise AttributeError(name)
