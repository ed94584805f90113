#!/usr/bin/env python
"""tc32 (split fp16 hi+lo tensor-core convolutions) against the fp32 SIMT mode of the same engine:
features and gradient of one tile, then step timing of both modes."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from style_transfer_b200 import netdesc, weights
from style_transfer_b200.engine import TileEngine, ContentData, StyleData

def l2rel(a, b):
    return float((a - b).double().norm() / b.double().norm())

def run(model, hw, c_layers, s_layers):
    net = netdesc.from_model(model)
    params = weights.he_normal(net)
    rs = np.random.RandomState(0)
    engs = {p: TileEngine(net, params, precision=p) for p in ('fp32', 'tc32')}
    img, content, style = (engs['fp32'].pil_to_image(rs.randint(0, 256, hw + (3,))) for _ in range(3))
    layers = ['conv1_2', 'conv2_2', 'conv3_3', 'conv4_2']
    f = {p: e.eval_features_tile(img, layers) for p, e in engs.items()}
    for l in layers:
        print('  features %-8s rel L2 %.3e  max rel %.3e' % (l, l2rel(f['tc32'][l], f['fp32'][l]),
              float((f['tc32'][l] - f['fp32'][l]).abs().max() / f['fp32'][l].abs().max())))
    ref = engs['fp32']
    ref.contents, ref.styles = [], []
    ref.preprocess_images([content], [style], c_layers, s_layers, max(hw))
    out = {}
    for p, e in engs.items():
        e.set_contents_and_styles(ref.contents, ref.styles)
        lw = {l: 1.0 for l in e.layers()}
        ol = e.ordered_layers(c_layers, s_layers)
        cw = {l: 0.05 / len(c_layers) for l in c_layers}
        sw = {l: 1.0 / len(s_layers) for l in s_layers}
        torch.cuda.synchronize(); t = time.perf_counter()
        out[p] = e.eval_sc_grad_tile(img, (0, 0), ol, c_layers, s_layers, [], lw, cw, sw, {})
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        t = time.perf_counter()
        for _ in range(3):
            e.eval_sc_grad_tile(img, (0, 0), ol, c_layers, s_layers, [], lw, cw, sw, {})
        torch.cuda.synchronize()
        print('  %s: loss %.8e  first call %.1f ms, then %.2f ms per evaluation' %
              (p, out[p][0], dt * 1e3, (time.perf_counter() - t) / 3 * 1e3))
    g32, gtc = out['fp32'][1], out['tc32'][1]
    err = (gtc - g32).abs() / g32.abs().max()
    print('  gradient tc32 vs fp32: rel L2 %.3e  max rel %.3e  frac > 1e-3: %.2e  loss rel %.2e' %
          (l2rel(gtc, g32), float(err.max()), float((err > 1e-3).float().mean()),
           abs(out['tc32'][0] - out['fp32'][0]) / abs(out['fp32'][0])))

if __name__ == '__main__':
    print('vgg16 96x128'); run('vgg16.prototxt', (96, 128), ['conv4_2'], ['conv1_1', 'conv3_1'])
    print('vgg19 256x256'); run('vgg19.prototxt', (256, 256), ['conv4_2'], ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1'])
    print('vgg19 512x512'); run('vgg19.prototxt', (512, 512), ['conv4_2'], ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1'])
