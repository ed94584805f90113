"""Oracle (TEST INFRASTRUCTURE): a miniature of the pycaffe ``Net`` surface the reference uses.

Restates (external, unpinned Caffe -- see oracle/__init__.py) exactly the calls made at
style_transfer.py:370 (``caffe.Net(deploy, 1, weights=...)``), :423-425 / :559-566
(``blobs['data'].reshape``, ``forward(end=...)``) and :608-610 (``backward(start=, end=)``):

* layers are addressed BY NAME and both ends of a forward/backward range are inclusive;
* ReLU layers are separate layers working in place on the conv's top blob
  (vgg19.prototxt:27-32), so ``forward(end='conv4_2')`` stops BEFORE ``relu4_2``;
* ``backward(start='conv4_2')`` starts AT the conv layer, i.e. whatever sits in
  ``diff['conv4_2']`` is taken as the gradient w.r.t. the conv output, with no ReLU mask;
* conv backward overwrites its bottom diff; ``force_backward: true`` (vgg19.prototxt:2) makes the
  chain reach ``diff['data']``.

The layer graph is generated here from the VGG configuration (the six bundled prototxts differ
only in depth, pool type and the ``_big`` re-wiring of conv2_1, reference vgg19_big.prototxt:62).
"""

from collections import OrderedDict

import numpy as np

from . import caffe_ops as ops

VGG_CFG = {
    'vgg16': (2, 2, 3, 3, 3),
    'vgg19': (2, 2, 4, 4, 4),
}
VGG_WIDTH = (64, 128, 256, 512, 512)


def vgg_layers(arch='vgg19', pool='max', big=False):
    """Returns the ordered layer list [(kind, name, bottom, top, extra)] of a bundled prototxt."""
    layers = []
    bottom = 'data'
    cin = 3
    for b, (n, cout) in enumerate(zip(VGG_CFG[arch], VGG_WIDTH), start=1):
        for i in range(1, n + 1):
            name = 'conv%d_%d' % (b, i)
            layers.append(('conv', name, bottom, name, (cin, cout)))
            layers.append(('relu', 'relu%d_%d' % (b, i), name, name, None))
            bottom, cin = name, cout
        pname = 'pool%d' % b
        layers.append(('pool', pname, bottom, pname, pool))
        if big and b == 1:
            continue            # conv2_1 keeps reading conv1_2; pool1 becomes a dead end
        bottom = pname
    return layers


def model_layers(model_name):
    """Maps a bundled prototxt file name (e.g. 'vgg19_avgpool.prototxt') to its layer list."""
    stem = model_name.rsplit('/', 1)[-1].replace('.prototxt', '')
    arch, _, variant = stem.partition('_')
    return vgg_layers(arch, pool='ave' if variant == 'avgpool' else 'max', big=variant == 'big')


def he_normal_weights(layers, seed=1234):
    """Synthetic VGG weights: N(0, sqrt(2/(9 Cin))) OIHW f32, zero bias."""
    rng = np.random.RandomState(seed)
    params = OrderedDict()
    for kind, name, _, _, extra in layers:
        if kind == 'conv':
            cin, cout = extra
            w = rng.randn(cout, cin, 3, 3).astype(np.float32) * np.float32(np.sqrt(2 / (9 * cin)))
            params[name] = (w, np.zeros(cout, np.float32))
    return params


class OracleNet:
    """Blobs ``data`` / ``diff`` keyed by blob name + ``forward`` / ``backward`` keyed by layer."""

    def __init__(self, layers, params, compute_weight_grads=False):
        self.layers = list(layers)
        self.params = params
        self.index = {name: i for i, (_, name, _, _, _) in enumerate(self.layers)}
        self.data = {}
        self.diff = {}
        self._argmax = {}
        # The reference's Caffe computes dW/db on every backward and never uses them
        # (SURVEY 8a6); the CPU-baseline timing leg switches this on to pay the same cost.
        self.compute_weight_grads = compute_weight_grads

    def blob_names(self):
        return ['data'] + [top for kind, _, _, top, _ in self.layers if kind != 'relu']

    def set_input(self, img):
        """``blobs['data'].reshape(1, 3, h, w); data['data'] = img``."""
        self.data = {'data': np.ascontiguousarray(img, dtype=np.float32)}
        self.diff = {}
        self._argmax = {}

    def forward(self, end):
        for kind, name, bottom, top, extra in self.layers[:self.index[end] + 1]:
            if kind == 'conv':
                w, b = self.params[name]
                self.data[top] = ops.conv3x3_forward(self.data[bottom], w, b)
            elif kind == 'relu':
                ops.relu_forward_(self.data[top])
            elif extra == 'max':
                self.data[top], self._argmax[name] = ops.maxpool_forward(self.data[bottom])
            else:
                self.data[top] = ops.avepool_forward(self.data[bottom])
            if kind != 'relu':
                self.diff.setdefault(top, np.zeros_like(self.data[top]))
        self.diff.setdefault('data', np.zeros_like(self.data['data']))

    def backward(self, start, end=None):
        lo = 0 if end is None else self.index[end]
        for kind, name, bottom, top, extra in reversed(self.layers[lo:self.index[start] + 1]):
            if top not in self.data:
                continue                    # dead-end branch that forward never reached
            if kind == 'conv':
                w, _ = self.params[name]
                if self.compute_weight_grads:
                    ops.conv3x3_backward_weight(self.diff[top], self.data[bottom])
                self.diff[bottom] = ops.conv3x3_backward_data(self.diff[top], w)
            elif kind == 'relu':
                ops.relu_backward_(self.diff[top], self.data[top])
            else:
                shape = self.data[bottom].shape
                if extra == 'max':
                    d = ops.maxpool_backward(self.diff[top], self._argmax[name], shape)
                else:
                    d = ops.avepool_backward(self.diff[top], shape)
                if self._shares_bottom(name):
                    self.diff[bottom] += d
                else:
                    self.diff[bottom] = d

    def _shares_bottom(self, pool_name):
        """True for pool1 of the ``_big`` nets, whose bottom (conv1_2) also feeds conv2_1: Caffe
        inserts a Split layer there whose backward SUMS the two top diffs, so the pool's
        contribution is added to what conv2_1's backward already wrote."""
        i = self.index[pool_name]
        bottom = self.layers[i][2]
        return any(b == bottom and k == 'conv' for k, _, b, _, _ in self.layers[i + 1:])
