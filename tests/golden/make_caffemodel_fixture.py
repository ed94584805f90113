#!/usr/bin/env python3
"""Writes tests/golden/caffemodel_*.bin: small serialized Caffe ``NetParameter`` messages produced by
Google's protobuf library (NOT by this repository's wire-format code) from a schema that restates the
relevant subset of BVLC Caffe's ``caffe.proto`` with its field numbers:

    message BlobShape        { repeated int64 dim = 1 [packed = true]; }
    message BlobProto        { optional BlobShape shape = 7;
                               repeated float data = 5 [packed = true]; repeated float diff = 6 [packed = true];
                               repeated double double_data = 8 [packed = true];
                               optional int32 num = 1, channels = 2, height = 3, width = 4; }
    message ConvolutionParameter { optional uint32 num_output = 1; repeated uint32 pad = 3, kernel_size = 4; }
    message LayerParameter   { optional string name = 1, type = 2; repeated string bottom = 3, top = 4;
                               repeated BlobProto blobs = 7; optional ConvolutionParameter convolution_param = 106; }
    message V1LayerParameter { repeated string bottom = 2, top = 3; optional string name = 4;
                               optional LayerType type = 5 (CONVOLUTION = 4, RELU = 18, POOLING = 17);
                               repeated BlobProto blobs = 6; repeated float blobs_lr = 7;
                               optional ConvolutionParameter convolution_param = 10; }
    message NetParameter     { optional string name = 1; repeated V1LayerParameter layers = 2;
                               repeated string input = 3; repeated int32 input_dim = 4;
                               repeated LayerParameter layer = 100; }

The published VGG ILSVRC files (download_models.sh:5-6) are V1 files (``layers = 2``, legacy
num/channels/height/width blob shapes); files saved by current Caffe use ``layer = 100`` and
``BlobShape``.  A third fixture stores ``data`` UNPACKED (one key per float), which old writers may
emit and every protobuf reader must accept.  The expected arrays are stored beside the bytes.
"""
import os

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

OUT = os.path.dirname(os.path.abspath(__file__))
F = descriptor_pb2.FieldDescriptorProto


def schema(packed_data=True):
    fd = descriptor_pb2.FileDescriptorProto(name='caffe_subset_%d.proto' % packed_data,
                                            package='caffe%d' % packed_data, syntax='proto2')
    pkg = '.caffe%d.' % packed_data

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, extra in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if 'type_name' in extra:
                f.type_name = pkg + extra['type_name']
            if 'packed' in extra:
                f.options.packed = extra['packed']
    REP, OPT = F.LABEL_REPEATED, F.LABEL_OPTIONAL
    msg('BlobShape', [('dim', 1, F.TYPE_INT64, REP, {'packed': True})])
    msg('BlobProto', [('num', 1, F.TYPE_INT32, OPT, {}), ('channels', 2, F.TYPE_INT32, OPT, {}),
                      ('height', 3, F.TYPE_INT32, OPT, {}), ('width', 4, F.TYPE_INT32, OPT, {}),
                      ('data', 5, F.TYPE_FLOAT, REP, {'packed': packed_data}),
                      ('diff', 6, F.TYPE_FLOAT, REP, {'packed': packed_data}),
                      ('shape', 7, F.TYPE_MESSAGE, OPT, {'type_name': 'BlobShape'}),
                      ('double_data', 8, F.TYPE_DOUBLE, REP, {'packed': True})])
    msg('ConvolutionParameter', [('num_output', 1, F.TYPE_UINT32, OPT, {}),
                                 ('pad', 3, F.TYPE_UINT32, REP, {}),
                                 ('kernel_size', 4, F.TYPE_UINT32, REP, {})])
    msg('LayerParameter', [('name', 1, F.TYPE_STRING, OPT, {}), ('type', 2, F.TYPE_STRING, OPT, {}),
                           ('bottom', 3, F.TYPE_STRING, REP, {}), ('top', 4, F.TYPE_STRING, REP, {}),
                           ('blobs', 7, F.TYPE_MESSAGE, REP, {'type_name': 'BlobProto'}),
                           ('convolution_param', 106, F.TYPE_MESSAGE, OPT,
                            {'type_name': 'ConvolutionParameter'})])
    msg('V1LayerParameter', [('bottom', 2, F.TYPE_STRING, REP, {}), ('top', 3, F.TYPE_STRING, REP, {}),
                             ('name', 4, F.TYPE_STRING, OPT, {}), ('type', 5, F.TYPE_INT32, OPT, {}),
                             ('blobs', 6, F.TYPE_MESSAGE, REP, {'type_name': 'BlobProto'}),
                             ('blobs_lr', 7, F.TYPE_FLOAT, REP, {}),
                             ('convolution_param', 10, F.TYPE_MESSAGE, OPT,
                              {'type_name': 'ConvolutionParameter'})])
    msg('NetParameter', [('name', 1, F.TYPE_STRING, OPT, {}),
                         ('layers', 2, F.TYPE_MESSAGE, REP, {'type_name': 'V1LayerParameter'}),
                         ('input', 3, F.TYPE_STRING, REP, {}), ('input_dim', 4, F.TYPE_INT32, REP, {}),
                         ('layer', 100, F.TYPE_MESSAGE, REP, {'type_name': 'LayerParameter'})])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('caffe%d.NetParameter' % packed_data))


def main():
    rs = np.random.RandomState(424242)
    shapes = [('conv1_1', (8, 3, 3, 3)), ('conv1_2', (8, 8, 3, 3)), ('conv2_1', (16, 8, 3, 3))]
    arrays = {}
    for name, shp in shapes:
        arrays[name + '_w'] = rs.randn(*shp).astype(np.float32)
        arrays[name + '_b'] = rs.randn(shp[0]).astype(np.float32)

    def fill_blob(b, arr, modern, as_double=False):
        if modern:
            b.shape.dim.extend(arr.shape)
        else:
            dims = arr.shape if arr.ndim == 4 else (1, 1, 1, arr.size)
            b.num, b.channels, b.height, b.width = (int(d) for d in dims)
        if as_double:
            b.double_data.extend(float(v) for v in arr.ravel())
        else:
            b.data.extend(float(v) for v in arr.ravel())
            b.diff.extend([0.0] * min(arr.size, 5))          # training leftovers: must be skipped

    for tag, packed, v1 in (('v1', True, True), ('v1_unpacked', False, True), ('v2', True, False)):
        Net = schema(packed)
        net = Net(name='VGG_fixture')
        net.input.append('data')
        net.input_dim.extend([1, 3, 32, 32])
        prev = 'data'
        for i, (name, shp) in enumerate(shapes):
            w, b = arrays[name + '_w'], arrays[name + '_b']
            if v1:
                l = net.layers.add(name=name, type=4)
                l.bottom.append(prev), l.top.append(name), l.blobs_lr.extend([1.0, 2.0])
                relu = net.layers.add(name='relu' + name[4:], type=18)
                relu.bottom.append(name), relu.top.append(name)
            else:
                l = net.layer.add(name=name, type='Convolution')
                l.bottom.append(prev), l.top.append(name)
                relu = net.layer.add(name='relu' + name[4:], type='ReLU')
                relu.bottom.append(name), relu.top.append(name)
            l.convolution_param.num_output = shp[0]
            l.convolution_param.pad.append(1), l.convolution_param.kernel_size.append(3)
            fill_blob(l.blobs.add(), w, modern=not v1)
            fill_blob(l.blobs.add(), b, modern=not v1, as_double=(not v1 and i == 2))
            prev = name
        with open(os.path.join(OUT, 'caffemodel_%s.bin' % tag), 'wb') as f:
            f.write(net.SerializeToString())
    np.savez(os.path.join(OUT, 'caffemodel_expected.npz'), **arrays)
    print('wrote', sorted(p for p in os.listdir(OUT) if p.startswith('caffemodel')))


if __name__ == '__main__':
    main()
