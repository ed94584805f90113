"""GPU check of the tcgen05 convolution against the SIMT convolution (both bf16 storage), plus a
first timing of one 512x512 VGG-19 tile evaluation.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from style_transfer_b200 import netdesc, weights
from style_transfer_b200.engine import TileEngine, ContentData, StyleData

def engine(model, precision, tc=True):
    if tc: os.environ.pop('ST_DISABLE_TC', None)
    else: os.environ['ST_DISABLE_TC'] = '1'
    net = netdesc.from_model(model)
    return TileEngine(net, weights.he_normal(net), precision=precision)

def rel(a, b): return float((a - b).norm() / b.norm())

LAYERS = ['conv1_1', 'conv1_2', 'pool1', 'conv2_1', 'conv2_2', 'conv3_1', 'conv3_4', 'conv4_1', 'conv4_2', 'conv5_1']
STYLE = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']

def targets(eng, h, w, rs):
    img = torch.from_numpy(rs.rand(3, h, w).astype(np.float32) * 255 - 120).cuda()
    f = eng.eval_features_tile(img, STYLE + ['conv4_2'])
    eng.set_contents_and_styles([ContentData({'conv4_2': f['conv4_2']})],
                                [StyleData({l: eng.gram_matrix(f[l]) for l in STYLE})])

def main():
    rs = np.random.RandomState(0)
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'check'):
        for (h, w) in [(64, 64), (96, 160), (45, 77), (256, 256)]:
            img = torch.from_numpy(rs.rand(3, h, w).astype(np.float32) * 255 - 120).cuda()
            e_tc, e_simt, e_f32 = engine('vgg19.prototxt', 'bf16', True), engine('vgg19.prototxt', 'bf16', False), engine('vgg19.prototxt', 'fp32')
            e_h = engine('vgg19.prototxt', 'fp16', True)
            f_tc, f_s, f_32 = (e.eval_features_tile(img, LAYERS) for e in (e_tc, e_simt, e_f32))
            torch.cuda.synchronize()
            print('features %dx%d' % (h, w), ' '.join('%s tc/simt %.1e tc/f32 %.1e |' % (l, rel(f_tc[l], f_s[l]), rel(f_tc[l], f_32[l])) for l in LAYERS[1:]), flush=True)
            lw = {l: 1.0 for l in e_tc.layers()}
            cw, sw = {'conv4_2': 0.05}, {l: 0.2 for l in STYLE}
            f_h = e_h.eval_features_tile(img, LAYERS)
            print('   fp16/f32 ', ' '.join('%s %.1e |' % (l, rel(f_h[l], f_32[l])) for l in LAYERS[1:]), flush=True)
            grads = []
            for e in (e_tc, e_simt, e_f32, e_h):
                targets(e, h, w, np.random.RandomState(1))
                layers = e.ordered_layers(STYLE, ['conv4_2'])
                loss, g = e.eval_sc_grad_tile(img, (0, 0), layers, ['conv4_2'], STYLE, [], lw, cw, sw, {})
                grads.append((loss, g.clone()))
            print('  sc_grad loss tc %.6e simt %.6e f32 %.6e fp16 %.6e | grad tc/simt %.2e tc/f32 %.2e simt/f32 %.2e fp16/f32 %.2e' % (
                grads[0][0], grads[1][0], grads[2][0], grads[3][0], rel(grads[0][1], grads[1][1]), rel(grads[0][1], grads[2][1]), rel(grads[1][1], grads[2][1]), rel(grads[3][1], grads[2][1])), flush=True)
    if which in ('all', 'time'):
        for prec in ('bf16', 'fp32'):
            e = engine('vgg19.prototxt', prec)
            h = w = 512
            img = torch.from_numpy(rs.rand(3, h, w).astype(np.float32) * 255 - 120).cuda()
            targets(e, h, w, np.random.RandomState(1))
            lw = {l: 1.0 for l in e.layers()}
            cw, sw = {'conv4_2': 0.05}, {l: 0.2 for l in STYLE}
            layers = e.ordered_layers(STYLE, ['conv4_2'])
            for _ in range(2): e.eval_sc_grad_tile(img, (0, 0), layers, ['conv4_2'], STYLE, [], lw, cw, sw, {})
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            n = 5
            t0.record()
            for _ in range(n): e.eval_sc_grad_tile(img, (0, 0), layers, ['conv4_2'], STYLE, [], lw, cw, sw, {})
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / n
            print('tile-eval 512x512 vgg19 %s: %.3f ms  (conv %.1f TFLOP/s incl. everything)' % (prec, ms, 378.7e9 / ms / 1e9), flush=True)

if __name__ == '__main__':
    main()
