"""Runs the five BASELINE.json configurations (single GPU unless launched under torchrun) for a few
iterations each through the public Python surface and prints iterations/s: evidence that every
configured shape -- VGG-16/19, max/average pooling, Adam/L-BFGS, TV, Deep-Dream layer, ragged tile
grids of the sqrt(2) scale ladder -- runs on the engine.  Synthetic images, random He-normal weights."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

from style_transfer_b200 import netdesc, weights
from style_transfer_b200.cli import transfer_multiscale
from style_transfer_b200.engine import TileEngine
from style_transfer_b200.transfer import StyleTransfer, default_args

CONFIGS = {
    'cfg1': dict(model='vgg16.prototxt', size=256, min_size=256, tile_size=512, optimizer='adam',
                 content_layers=['conv4_2'], style_layers=['conv3_1'], iterations=[20]),
    'cfg2': dict(model='vgg19.prototxt', size=512, min_size=512, optimizer='lbfgs', iterations=[20]),
    'cfg3': dict(model='vgg19.prototxt', size=2048, min_size=2048, tile_size=512, optimizer='adam',
                 iterations=[10]),
    'cfg4': dict(model='vgg19_avgpool.prototxt', size=4096, min_size=4096, tile_size=1024,
                 optimizer='lbfgs', tv_weight=5.0, iterations=[6]),
    'cfg5': dict(model='vgg19.prototxt', size=2048, min_size=256, tile_size=512, optimizer='adam',
                 dd_layers=['conv5_1'], dd_weight=0.1, iterations=[4]),
}


def main():
    p = argparse.ArgumentParser()
    p.add_argument('configs', nargs='*', default=list(CONFIGS))
    p.add_argument('--precision', default='fp16')
    a = p.parse_args()
    rs = np.random.RandomState(0)
    out = {}
    for name in a.configs:
        cfg = CONFIGS[name]
        args = default_args(**cfg)
        args.style_scale, args.max_style_size, args.style_scale_up = 1.0, None, False
        net = netdesc.from_model(args.model)
        eng = TileEngine(net, weights.he_normal(net), mean=args.mean, precision=a.precision)
        st = StyleTransfer(eng, args)
        content = Image.fromarray(rs.randint(0, 256, (args.size, args.size, 3)).astype(np.uint8))
        style = Image.fromarray(rs.randint(0, 256, (args.size, args.size, 3)).astype(np.uint8))
        np.random.seed(0)
        marks = []

        def cb(step, loss, scale, size, **kw):
            torch.cuda.synchronize()
            marks.append((scale, size, step, time.perf_counter(), float(loss)))
        t0 = time.perf_counter()
        transfer_multiscale(st, args, [content], [style], callback=cb)
        torch.cuda.synchronize()
        total = time.perf_counter() - t0
        # steady-state rate of the LAST scale: skip its first iteration (first-touch allocations)
        last = [m for m in marks if m[0] == marks[-1][0]]
        its = (len(last) - 1) / (last[-1][3] - last[0][3]) if len(last) > 1 else float('nan')
        # the same loop body without a per-iteration callback (nothing synchronises the host)
        n_free = 30 if args.size <= 1024 else 8
        for _ in range(3):
            st.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        h0 = time.perf_counter()
        e0.record()
        for _ in range(n_free):
            st.step()
        e1.record()
        h1 = time.perf_counter()
        torch.cuda.synchronize()
        free = n_free / (e0.elapsed_time(e1) * 1e-3)
        out[name] = {'iterations_per_s_free_running': free,
                     'host_enqueue_ms_per_step': (h1 - h0) * 1e3 / n_free,
                     'iterations_per_s_last_scale': its, 'last_scale_size': list(last[-1][1]),
                     'scales': marks[-1][0], 'total_s': total, 'final_loss': last[-1][4],
                     'finite': bool(np.isfinite(last[-1][4]))}
        print(name, json.dumps(out[name]), flush=True)
        del st, eng
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
