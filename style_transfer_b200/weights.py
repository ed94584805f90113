"""VGG weights for the tile engine.

The reference loads ``vgg16/19.caffemodel`` through Caffe (style_transfer.py:370); those files are
downloaded by download_models.sh and are not available offline.  BASELINE.json prescribes
random-init weights for the benchmark: He-normal N(0, sqrt(2/(9*Cin))) float32 OIHW, zero bias,
RandomState(1234).  A .caffemodel reader is a later item.
"""

from collections import OrderedDict

import numpy as np


def he_normal(net, seed=1234):
    rng = np.random.RandomState(seed)
    params = OrderedDict()
    for _, layer in net.conv_layers():
        std = np.float32(np.sqrt(2 / (9 * layer.cin)))
        w = rng.randn(layer.cout, layer.cin, 3, 3).astype(np.float32) * std
        params[layer.name] = (w, np.zeros(layer.cout, np.float32))
    return params


def load_npz(path):
    """Weights saved as ``<layer>_w`` / ``<layer>_b`` arrays in an .npz file."""
    data = np.load(path)
    names = sorted({k[:-2] for k in data.files})
    return OrderedDict((n, (data[n + '_w'], data[n + '_b'])) for n in names)


def save_npz(path, params):
    """Writes weights in the layout ``load_npz`` reads (``<layer>_w`` OIHW f32, ``<layer>_b``)."""
    arrays = {}
    for name, (w, b) in params.items():
        arrays[name + '_w'], arrays[name + '_b'] = np.float32(w), np.float32(b)
    np.savez(path, **arrays)
