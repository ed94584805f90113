"""The .caffemodel reader against files written by a minimal protobuf encoder (both the V1 layout of
the published VGG files and the current LayerParameter layout, packed and legacy blob shapes)."""

import struct

import numpy as np
import pytest


def varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def field(num, wt, payload):
    return varint((num << 3) | wt) + (varint(len(payload)) + payload if wt == 2 else payload)


def blob(arr, modern):
    arr = np.asarray(arr, np.float32)
    data = field(5, 2, arr.astype('<f4').tobytes())
    if modern:
        shape = field(7, 2, field(1, 2, b''.join(varint(d) for d in arr.shape)))
        return shape + data
    dims = (list(arr.shape) if arr.ndim == 4 else [1, 1, 1, arr.size])
    return b''.join(field(i + 1, 0, varint(d)) for i, d in enumerate(dims)) + data


@pytest.mark.parametrize('modern', [False, True])
def test_caffemodel_roundtrip(tmp_path, modern):
    from style_transfer_b200 import caffemodel, netdesc, weights
    net = netdesc.from_model('vgg16.prototxt')
    params = weights.he_normal(net, seed=3)
    for name in params:
        params[name] = (params[name][0], np.random.RandomState(1).randn(params[name][1].size).astype(np.float32))
    msg = field(1, 2, b'VGG_test')
    for name, (w, b) in params.items():
        if modern:
            layer = field(1, 2, name.encode()) + field(2, 2, b'Convolution') + \
                field(7, 2, blob(w, True)) + field(7, 2, blob(b, True))
            msg += field(100, 2, layer)
            msg += field(100, 2, field(1, 2, ('relu_' + name).encode()) + field(2, 2, b'ReLU'))
        else:
            layer = field(4, 2, name.encode()) + field(5, 0, varint(4)) + \
                field(6, 2, blob(w, False)) + field(6, 2, blob(b, False))
            msg += field(2, 2, layer)
    path = tmp_path / 'net.caffemodel'
    path.write_bytes(msg)
    got = caffemodel.load_caffemodel(str(path), net)
    assert list(got) == list(params)
    for name in params:
        assert np.array_equal(got[name][0], params[name][0])
        assert np.array_equal(got[name][1], params[name][1])
    with pytest.raises(KeyError):                # vgg19 has layers this file lacks
        caffemodel.load_caffemodel(str(path), netdesc.from_model('vgg19.prototxt'))


def test_caffemodel_shape_mismatch(tmp_path):
    from style_transfer_b200 import caffemodel, netdesc
    w = np.zeros((64, 3, 5, 5), np.float32)
    layer = field(1, 2, b'conv1_1') + field(7, 2, blob(w, True)) + field(7, 2, blob(np.zeros(64), True))
    path = tmp_path / 'bad.caffemodel'
    path.write_bytes(field(100, 2, layer))
    with pytest.raises((ValueError, KeyError)):
        caffemodel.load_caffemodel(str(path), netdesc.from_model('vgg16.prototxt'))


@pytest.mark.parametrize('tag', ['v1', 'v1_unpacked', 'v2'])
def test_caffemodel_fixtures_from_the_protobuf_library(golden_dir, tag):
    """Independent bytes: tests/golden/caffemodel_*.bin were serialized by Google's protobuf library
    from the field numbers of BVLC Caffe's caffe.proto (tests/golden/make_caffemodel_fixture.py), not
    by any encoder of this repository.  v1: ``layers = 2`` / V1LayerParameter with legacy
    num/channels/height/width shapes (the format of the published VGG files, download_models.sh:5-6);
    v1_unpacked: the same with ``data`` stored one key per float; v2: ``layer = 100`` /
    LayerParameter with BlobShape, one bias stored as double_data.  Every file also carries fields the
    reader must skip (diff, blobs_lr, convolution_param, bottom / top, ReLU layers without blobs)."""
    import os
    from style_transfer_b200 import caffemodel
    want = np.load(os.path.join(golden_dir, 'caffemodel_expected.npz'))
    got = caffemodel.load_caffemodel(os.path.join(golden_dir, 'caffemodel_%s.bin' % tag))
    assert list(got) == ['conv1_1', 'conv1_2', 'conv2_1']
    for name, (w, b) in got.items():
        assert w.dtype == np.float32 and np.array_equal(w, want[name + '_w']), name
        assert b.dtype == np.float32 and np.array_equal(b, want[name + '_b']), name
    blobs = caffemodel.read_blobs(os.path.join(golden_dir, 'caffemodel_%s.bin' % tag))
    assert all(not k.startswith('relu') for k in blobs)
