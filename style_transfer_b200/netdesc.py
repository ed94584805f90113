"""Network descriptions for the tile engine.

The reference ships six Caffe deploy files (vgg{16,19}{,_avgpool,_big}.prototxt) and tabulates the
blob shapes of the first four at style_transfer.py:1030-1073; everything else about the graph lives
in Caffe's prototxt parser.  Here the graph is data: a list of 3x3/pad-1 convolutions (each followed
by an in-place ReLU, vgg19.prototxt:27-32) and 2x2/stride-2 pools.  Bundled model names are
generated from the VGG configuration; any other ``--model`` file goes through ``parse_prototxt``,
a reader for exactly these layer types.
"""

import ctypes as C
import os
import re
from dataclasses import dataclass

from . import _lib

VGG_CFG = {'vgg16': (2, 2, 3, 3, 3), 'vgg19': (2, 2, 4, 4, 4)}
VGG_WIDTH = (64, 128, 256, 512, 512)
BUNDLED = tuple('%s%s.prototxt' % (a, v) for a in VGG_CFG for v in ('', '_avgpool', '_big'))


@dataclass
class Layer:
    kind: str          # 'conv' | 'pool'
    name: str          # layer name == top blob name
    bottom: str
    cin: int
    cout: int
    pool: str = ''     # 'max' | 'ave'


class NetDesc:
    def __init__(self, layers, name=''):
        self.name = name
        self.layers = list(layers)
        self.blob_names = ['data'] + [l.name for l in self.layers]
        self._index = {n: i for i, n in enumerate(self.blob_names)}
        if len(self._index) != len(self.blob_names):
            raise ValueError('duplicate blob names in network')
        # Blob shapes at the canonical 224x224 input, as in VGG16_SHAPES / VGG19_SHAPES
        # (style_transfer.py:1030-1073) or as probed by init_model (:1013-1027).
        self.shapes = {}
        size = {'data': 224}
        chans = {'data': 3}
        for l in self.layers:
            if l.bottom not in size:
                raise ValueError('layer %s reads unknown blob %s' % (l.name, l.bottom))
            if l.cin != chans[l.bottom]:
                raise ValueError('layer %s: channel mismatch' % l.name)
            size[l.name] = size[l.bottom] if l.kind == 'conv' else (size[l.bottom] + 1) // 2
            chans[l.name] = l.cout
            self.shapes[l.name] = (l.cout, size[l.name], size[l.name])
        self.last_layer = self.layers[-1].name

    def layer_names(self):
        """``CaffeModel.layers()`` (style_transfer.py:403-413)."""
        return list(self.shapes)

    def blob_index(self, name):
        if name not in self._index:
            raise KeyError('unknown layer %r (known: %s)' % (name, ', '.join(self.blob_names[1:])))
        return self._index[name]

    def layer_info(self, name):
        """(scale vs. the image, channels) -- style_transfer.py:415-419."""
        shape = self.shapes[name]
        return 224 // shape[1], shape[0]

    def feature_hw(self, name, h, w):
        """Spatial size of blob ``name`` for an h x w input (ceil-mode pooling)."""
        sizes = {'data': (h, w)}
        for l in self.layers:
            bh, bw = sizes[l.bottom]
            sizes[l.name] = (bh, bw) if l.kind == 'conv' else ((bh + 1) // 2, (bw + 1) // 2)
            if l.name == name:
                break
        return sizes[name]

    def to_ctypes(self):
        arr = (_lib.LayerDesc * len(self.layers))()
        for i, l in enumerate(self.layers):
            kind = _lib.ST_CONV3X3 if l.kind == 'conv' else (
                _lib.ST_POOL_MAX if l.pool == 'max' else _lib.ST_POOL_AVE)
            arr[i] = _lib.LayerDesc(kind, self._index[l.bottom], l.cin, l.cout)
        return arr

    def conv_layers(self):
        return [(i, l) for i, l in enumerate(self.layers) if l.kind == 'conv']


def vgg(arch='vgg19', pool='max', big=False):
    layers, bottom, cin = [], 'data', 3
    for b, (n, cout) in enumerate(zip(VGG_CFG[arch], VGG_WIDTH), start=1):
        for i in range(1, n + 1):
            name = 'conv%d_%d' % (b, i)
            layers.append(Layer('conv', name, bottom, cin, cout))
            bottom, cin = name, cout
        layers.append(Layer('pool', 'pool%d' % b, bottom, cin, cin, pool))
        if not (big and b == 1):        # *_big: conv2_1 reads conv1_2 (vgg19_big.prototxt:62)
            bottom = 'pool%d' % b
    return NetDesc(layers, arch + ('_avgpool' if pool == 'ave' else '') + ('_big' if big else ''))


def from_model(path):
    """``--model`` handling: bundled names are table-driven like the reference's shape tables
    (style_transfer.py:1098-1101); other files are parsed."""
    base = os.path.basename(str(path))
    if base in BUNDLED and not os.path.exists(str(path)):
        stem = base[:-len('.prototxt')]
        arch, _, variant = stem.partition('_')
        return vgg(arch, 'ave' if variant == 'avgpool' else 'max', variant == 'big')
    with open(str(path)) as f:
        return parse_prototxt(f.read(), name=base)


# ---- a reader for the subset of prototxt the deploy files use ---------------------------------------
_TOKEN = re.compile(r'\s*(?:#[^\n]*\n)?\s*([{}]|[A-Za-z_][\w.]*\s*:?|"[^"]*"|\'[^\']*\'|[-+.\w]+)')


def _tokens(text):
    pos, out = 0, []
    text = re.sub(r'#[^\n]*', '', text)
    while True:
        m = _TOKEN.match(text, pos)
        if not m:
            if text[pos:].strip():
                raise ValueError('prototxt syntax error near %r' % text[pos:pos + 30])
            return out
        out.append(m.group(1).strip())
        pos = m.end()


def _parse_block(toks, i):
    """Parses ``key: value`` / ``key { ... }`` pairs until the closing brace; repeated keys
    collect into lists."""
    msg = {}
    while i < len(toks) and toks[i] != '}':
        key = toks[i].rstrip(':').strip()
        i += 1
        if toks[i] == '{':
            val, i = _parse_block(toks, i + 1)
            i += 1
        else:
            val = toks[i].strip('"\'')
            i += 1
        msg.setdefault(key, []).append(val)
    return msg, i


def parse_prototxt(text, name=''):
    root, _ = _parse_block(_tokens(text), 0)
    layers, chans = [], {}
    for spec in root.get('layer', []) + root.get('layers', []):
        kind = spec['type'][0]
        lname = spec['name'][0]
        top = spec.get('top', [None])[0]
        bottom = spec.get('bottom', [None])[0]
        if kind == 'Input':
            dims = spec['input_param'][0]['shape'][0]['dim']
            if top != 'data' or int(dims[1]) != 3:
                raise ValueError('the input blob must be a 3-channel blob called "data"')
            chans['data'] = 3
        elif kind == 'Convolution':
            p = spec['convolution_param'][0]
            if int(p['kernel_size'][0]) != 3 or int(p.get('pad', [0])[0]) != 1 or \
                    int(p.get('stride', [1])[0]) != 1 or top != lname:
                raise ValueError('%s: only 3x3 / pad 1 / stride 1 convolutions named after their '
                                 'top blob are supported' % lname)
            cout = int(p['num_output'][0])
            layers.append(Layer('conv', lname, bottom, chans[bottom], cout))
            chans[top] = cout
        elif kind == 'ReLU':
            if top != bottom or not layers or layers[-1].name != top or layers[-1].kind != 'conv':
                raise ValueError('%s: ReLU must run in place directly after a convolution' % lname)
        elif kind == 'Pooling':
            p = spec['pooling_param'][0]
            if int(p['kernel_size'][0]) != 2 or int(p.get('stride', [1])[0]) != 2 or \
                    int(p.get('pad', [0])[0]) != 0 or top != lname:
                raise ValueError('%s: only 2x2 / stride 2 pooling is supported' % lname)
            mode = {'MAX': 'max', 'AVE': 'ave'}[p.get('pool', ['MAX'])[0]]
            layers.append(Layer('pool', lname, bottom, chans[bottom], chans[bottom], mode))
            chans[top] = chans[bottom]
        else:
            raise ValueError('unsupported layer type %s (%s)' % (kind, lname))
    convs = [l for l in layers if l.kind == 'conv']
    relus = sum(1 for s in root.get('layer', []) + root.get('layers', []) if s['type'][0] == 'ReLU')
    if relus != len(convs):
        raise ValueError('every convolution must be followed by an in-place ReLU')
    return NetDesc(layers, name)


def to_prototxt(net):
    """Writes a deploy file in the dialect parse_prototxt reads (and Caffe would)."""
    out = ['name: "%s"' % net.name, 'force_backward: true',
           'layer {\n  top: "data"\n  name: "input"\n  type: "Input"\n  input_param {\n'
           '    shape {\n      dim: 1\n      dim: 3\n      dim: 224\n      dim: 224\n    }\n  }\n}']
    for l in net.layers:
        if l.kind == 'conv':
            out.append('layer {\n  bottom: "%s"\n  top: "%s"\n  name: "%s"\n  type: "Convolution"\n'
                       '  convolution_param {\n    num_output: %d\n    pad: 1\n    kernel_size: 3\n'
                       '  }\n}' % (l.bottom, l.name, l.name, l.cout))
            out.append('layer {\n  bottom: "%s"\n  top: "%s"\n  name: "%s"\n  type: "ReLU"\n}' %
                       (l.name, l.name, l.name.replace('conv', 'relu')))
        else:
            out.append('layer {\n  bottom: "%s"\n  top: "%s"\n  name: "%s"\n  type: "Pooling"\n'
                       '  pooling_param {\n    pool: %s\n    kernel_size: 2\n    stride: 2\n  }\n}' %
                       (l.bottom, l.name, l.name, l.pool.upper()))
    return '\n'.join(out) + '\n'
