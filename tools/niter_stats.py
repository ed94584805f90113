import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from test_gpu_parity import engine_for, rand_img, default_args, OracleTransfer
from style_transfer_b200.transfer import StyleTransfer
for opt, iters in (('adam', 6), ('adam', 20), ('lbfgs', 5)):
    for precision in ('fp32', 'fp16', 'bf16'):
        eng, ora = engine_for('vgg16.prototxt', precision, mean=(103.939, 116.779, 123.68))
        rs = np.random.RandomState(21)
        H, W = 64, 80
        content, style = rand_img(rs, H, W), rand_img(rs, H, W)
        args = default_args(tile_size=48, optimizer=opt, content_layers=['conv4_2'], style_layers=['conv3_1'])
        ot = OracleTransfer(ora, args); np.random.seed(0); ot.init_first_scale(H, W)
        want = ot.run(iters, [content], [style]).copy()
        st = StyleTransfer(eng, args); np.random.seed(0); st.init_first_scale(H, W)
        got = st.transfer(iters, [content], [style])
        err = np.abs(got.cpu().numpy() - want)
        print(opt, iters, precision, 'max %.3g rms %.3g q50/90/99 %s frac>8 %.3g' % (err.max(), np.sqrt((err.astype(np.float64)**2).mean()), np.round(np.quantile(err,[.5,.9,.99]),3), (err>8).mean()), flush=True)
