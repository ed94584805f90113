"""Runs N optimizer iterations of the bench workload (2048^2 image, 512^2 tiles, VGG-19, Adam) with a
short preprocessing (one feature pass instead of ten): the unit profiled with ncu."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from style_transfer_b200 import netdesc, weights
from style_transfer_b200.engine import ContentData, StyleData, TileEngine
from style_transfer_b200.transfer import StyleTransfer, default_args, parse_weights


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--size', type=int, default=2048)
    p.add_argument('--tile-size', type=int, default=512)
    p.add_argument('--steps', type=int, default=3)
    p.add_argument('--precision', default='fp16')
    p.add_argument('--optimizer', default='adam')
    p.add_argument('--profile-last', action='store_true',
                   help='bracket the last step with cudaProfilerStart/Stop (ncu --profile-from-start off)')
    a = p.parse_args()
    args = default_args(size=a.size, min_size=a.size, tile_size=a.tile_size, optimizer=a.optimizer)
    net = netdesc.from_model(args.model)
    eng = TileEngine(net, weights.he_normal(net), mean=args.mean, precision=a.precision)
    st = StyleTransfer(eng, args)
    rs = np.random.RandomState(1)
    np.random.seed(0)
    st.init_first_scale(a.size, a.size)
    st.c_layers, st.c_weight = parse_weights(args.content_layers, args.content_weight)
    st.s_layers, st.s_weight = parse_weights(args.style_layers, 1)
    st.d_layers, st.d_weight = parse_weights(args.dd_layers, args.dd_weight)
    st.jitter_scale = 8
    saved = eng.img
    eng.img = eng.to_device(eng.pil_to_image(rs.randint(0, 256, (a.size, a.size, 3))))
    feats = eng.eval_features_once(st.s_layers + st.c_layers, a.tile_size)
    eng.img = saved
    eng.set_contents_and_styles([ContentData({l: feats[l] for l in st.c_layers})],
                                [StyleData({l: eng.gram_matrix(feats[l]) for l in st.s_layers})])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(a.steps):
        if i == a.steps - 1:
            e0.record()
            if a.profile_last:
                torch.cuda.synchronize()
                torch.cuda.profiler.start()
        avg, loss = st.step()
    e1.record()
    torch.cuda.synchronize()
    if a.profile_last:
        torch.cuda.profiler.stop()
    # host time to ENQUEUE one step (no synchronisation inside): the launch-bound floor of a step
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        st.step()
    t_host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    print('host enqueue time per step: %.3f ms' % (t_host * 1e3), flush=True)
    print('step %dx%d/%d %s: last step %.3f ms, loss %.6e' %
          (a.size, a.size, a.tile_size, a.precision, e0.elapsed_time(e1), float(loss)), flush=True)


if __name__ == '__main__':
    main()
