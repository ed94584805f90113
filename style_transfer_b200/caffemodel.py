"""Reader for the weights of a ``.caffemodel`` (a serialized Caffe ``NetParameter``), without
protobuf or Caffe: the reference loads these files through ``caffe.Net(deploy, 1, weights=...)``
(style_transfer.py:370).  Only what the engine needs is decoded -- layer names and their blobs --
straight from the protobuf wire format (there is no ``protoc`` in this environment).

Fields used (caffe.proto, BVLC Caffe):
  NetParameter      : layers = 2 (V1LayerParameter, the format of the published VGG files),
                      layer = 100 (LayerParameter)
  V1LayerParameter  : name = 4, blobs = 6
  LayerParameter    : name = 1, blobs = 7
  BlobProto         : num/channels/height/width = 1..4, data = 5 (packed float), shape = 7,
                      double_data = 8
  BlobShape         : dim = 1 (packed int64)
"""

from collections import OrderedDict

import numpy as np


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one message; length-delimited values are
    memoryview slices (no copies)."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + n], pos + n
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield field, wt, val


def _packed_varints(buf):
    out, pos = [], 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(v)
    return out


def _blob(buf):
    legacy = {}
    shape, data = None, []
    for field, wt, val in _fields(buf):
        if field in (1, 2, 3, 4) and wt == 0:
            legacy[field] = val
        elif field == 5:                                    # float data, packed or repeated
            data.append(np.frombuffer(val, dtype='<f4'))
        elif field == 8:                                    # double data
            data.append(np.frombuffer(val, dtype='<f8').astype(np.float32))
        elif field == 7 and wt == 2:
            for f2, wt2, v2 in _fields(val):
                if f2 == 1:
                    shape = _packed_varints(v2) if wt2 == 2 else (shape or []) + [v2]
    arr = np.concatenate(data) if data else np.zeros(0, np.float32)
    if shape is None:
        shape = [legacy.get(i, 1) for i in (1, 2, 3, 4)]
    return arr.astype(np.float32, copy=False).reshape(shape)


def _layer(buf, name_field, blobs_field):
    name, blobs = None, []
    for field, wt, val in _fields(buf):
        if field == name_field and wt == 2:
            name = bytes(val).decode('utf-8')
        elif field == blobs_field and wt == 2:
            blobs.append(_blob(val))
    return name, blobs


def read_blobs(path):
    """{layer name: [blob arrays]} of every layer that carries blobs."""
    with open(path, 'rb') as f:
        buf = memoryview(f.read())
    layers = OrderedDict()
    for field, wt, val in _fields(buf):
        if wt != 2 or field not in (2, 100):
            continue
        name, blobs = _layer(val, 4, 6) if field == 2 else _layer(val, 1, 7)
        if name is not None and blobs:
            layers[name] = blobs
    return layers


def load_caffemodel(path, net=None):
    """Conv weights as the engine wants them: {layer: (OIHW float32, bias float32)}.  With ``net``
    (a NetDesc) the shapes are checked and only the net's convolution layers are returned."""
    params = OrderedDict()
    for name, blobs in read_blobs(path).items():
        if len(blobs) < 2:
            continue
        w = blobs[0]
        if w.ndim != 4:
            continue
        params[name] = (np.ascontiguousarray(w), np.ascontiguousarray(blobs[1].reshape(-1)))
    if net is not None:
        out = OrderedDict()
        for _, layer in net.conv_layers():
            if layer.name not in params:
                raise KeyError('layer %s is missing from %s' % (layer.name, path))
            w, b = params[layer.name]
            if w.shape != (layer.cout, layer.cin, 3, 3) or b.shape != (layer.cout,):
                raise ValueError('%s: blob shapes %s / %s do not match the network' %
                                 (layer.name, w.shape, b.shape))
            out[layer.name] = (w, b)
        return out
    return params
