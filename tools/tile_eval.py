"""Runs N evaluations of one 512x512 VGG-19 tile (5 style + 1 content layer) through the C ABI:
the unit profiled with ncu (launch list / --set full captures under profiles/)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from style_transfer_b200 import netdesc, weights
from style_transfer_b200.engine import ContentData, StyleData, TileEngine

STYLE = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--precision', default='fp16')
    p.add_argument('--size', type=int, default=512)
    p.add_argument('--evals', type=int, default=3)
    p.add_argument('--model', default='vgg19.prototxt')
    a = p.parse_args()
    net = netdesc.from_model(a.model)
    eng = TileEngine(net, weights.he_normal(net), precision=a.precision)
    rs = np.random.RandomState(0)
    h = w = a.size
    img = torch.from_numpy(rs.rand(3, h, w).astype(np.float32) * 255 - 120).cuda()
    tgt = torch.from_numpy(rs.rand(3, h, w).astype(np.float32) * 255 - 120).cuda()
    f = eng.eval_features_tile(tgt, STYLE + ['conv4_2'])
    eng.set_contents_and_styles([ContentData({'conv4_2': f['conv4_2']})],
                                [StyleData({l: eng.gram_matrix(f[l]) for l in STYLE})])
    lw = {l: 1.0 for l in eng.layers()}
    cw, sw = {'conv4_2': 0.05}, {l: 0.2 for l in STYLE}
    layers = eng.ordered_layers(STYLE, ['conv4_2'])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(a.evals):
        if i == a.evals - 1:
            e0.record()
        loss, g = eng.eval_sc_grad_tile(img, (0, 0), layers, ['conv4_2'], STYLE, [], lw, cw, sw, {})
    e1.record()
    torch.cuda.synchronize()
    print('tile-eval %dx%d %s %s: last eval %.3f ms, loss %.6e' %
          (h, w, a.model, a.precision, e0.elapsed_time(e1), loss), flush=True)


if __name__ == '__main__':
    main()
