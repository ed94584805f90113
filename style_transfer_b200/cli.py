"""Command-line front end with the reference's surface (style_transfer.py:1076-1167): parse the
flags, run ``transfer_multiscale`` (:832-909) on the CUDA engine, save the result.

    python style_transfer.py -ci CONTENT -si STYLE [-s 2048 --tile-size 512 --devices 0 1 2 3 ...]

``--devices`` with several entries re-launches the script under ``torch.distributed.run`` with one
rank per listed GPU (the reference fork()s one worker per device, :179-181); every rank holds the
image and the optimizer state, tiles are dealt round-robin, gradients are all-gathered.
"""

import csv
import json
import os
import sys
import time

import numpy as np

from . import config_system


def resize_to_fit(image, size, div=1, scale_up=False):
    """Resizes a PIL image to fit into a size-by-size square (style_transfer.py:963-976)."""
    from PIL import Image
    size = int(round(size)) // div * div
    w, h = image.size
    if not scale_up and max(w, h) <= size:
        return image
    if w > h:
        new_w, new_h = size, int(round(size * h / w)) // div * div
    else:
        new_h, new_w = size, int(round(size * w / h)) // div * div
    return image.resize((new_w, new_h), Image.LANCZOS)


def resize_f32(tensor, hw, method='lanczos'):
    """``num_utils.resize`` (:90-108): per-channel float resampling of a CUDA f32 [C][H][W] tensor.
    Default: on the device (``st_resize_f32``: Pillow's float resampling restated, bit-identical to
    PIL on the B200 -- tests/test_gpu_parity.py::test_device_resize_matches_oracle, and the two-scale
    run of tests/test_gpu_parity_configs.py).  ``ST_HOST_RESIZE=1`` goes through PIL 'F' images on
    the host instead, exactly as the reference does (once per scale)."""
    import torch
    if os.environ.get('ST_HOST_RESIZE') != '1' and tensor.is_cuda:
        return resize_f32_device(tensor, hw, method)
    from PIL import Image
    m = {'lanczos': Image.LANCZOS, 'bilinear': Image.BILINEAR}[method]
    a = tensor.detach().cpu().numpy().astype(np.float32)
    out = np.stack([np.asarray(Image.fromarray(ch).resize((hw[1], hw[0]), m), dtype=np.float32)
                    for ch in a])
    return torch.from_numpy(np.ascontiguousarray(out)).to(tensor.device)


def resize_f32_device(tensor, hw, method='lanczos'):
    """The scale change without leaving the GPU (``st_resize_f32``)."""
    import ctypes as C
    import torch
    from . import _lib
    t = tensor.detach().contiguous().float()
    c, h, w = t.shape
    out = torch.empty((c, hw[0], hw[1]), dtype=torch.float32, device=t.device)
    tmp = torch.empty((c, h, hw[1]), dtype=torch.float32, device=t.device)
    _lib.call('st_resize_f32', C.c_void_p(t.data_ptr()), c, h, w, int(hw[0]), int(hw[1]),
              {'lanczos': 0, 'bilinear': 1}[method], C.c_void_p(out.data_ptr()),
              C.c_void_p(tmp.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def style_multiscale_variants(image, vmin, vmax, div=1):
    """The scaled copies of one style image whose Gram matrices ``--style-multiscale MIN MAX``
    averages (preprocess_images :501-524): sizes from MAX down by sqrt(2) while >= max(32, MIN),
    visited smallest first, stopping after the first size that no longer shrinks the image and
    skipping copies whose short side falls below 32 pixels."""
    size, sizes = vmax, [vmax]
    while True:
        size = int(round(size / np.sqrt(2)))
        if size < max(32, vmin):
            break
        sizes.append(size)
    out, too_big = [], False
    for size in reversed(sizes):
        if too_big:
            break
        scaled = resize_to_fit(image, size, div)
        if max(scaled.size) == max(image.size):
            too_big = True
        if min(scaled.size) < 32:
            continue
        out.append(scaled)
    return out


class StatLogger:
    """Per-iteration CSV with the reference's columns (style_transfer.py:101-130)."""
    columns = ['iteration', 'scale', 'step', 'time', 'content_h', 'content_w', 'update_size', 'loss',
               'tv_norm']

    def __init__(self, path):
        self.path, self.rows, self.t0 = path, [], time.perf_counter()

    def add(self, **row):
        row['time'] = time.perf_counter() - self.t0
        self.rows.append(row)

    def dump(self):
        with open(self.path, 'w', newline='') as f:
            w = csv.DictWriter(f, self.columns)
            w.writeheader()
            w.writerows(self.rows)


def transfer_multiscale(st, args, content_images, style_images, initial_image=None, aux_image=None,
                        callback=None):
    """Scale ladder of the reference (:832-909).  ``st`` is a transfer.StyleTransfer."""
    from PIL import Image
    from .transfer import scale_ladder
    model = st.model
    sizes = scale_ladder(args.size, args.min_size)
    output = None
    for i, size in enumerate(sizes):
        content_scaled = [resize_to_fit(im, size, args.div, scale_up=True) for im in content_images]
        if any(im.size != content_scaled[0].size for im in content_scaled):
            raise ValueError('All of the content images must be the same size')
        w, h = content_scaled[0].size
        if model.rank == 0:
            print('\nScale %d, image size %dx%d.\n' % (i + 1, w, h), flush=True)
        style_scaled = []
        multiscale = getattr(args, 'style_multiscale', None)
        for im in style_images:
            if multiscale:
                # --style-multiscale scales the ORIGINAL style image itself (:857-858)
                style_scaled.append(im)
            elif args.style_scale >= 32:
                style_scaled.append(resize_to_fit(im, args.style_scale, args.div, scale_up=True))
            else:
                style_size = round(size * args.style_scale)
                if args.max_style_size is not None:
                    style_size = min(style_size, args.max_style_size)
                style_scaled.append(resize_to_fit(im, style_size, args.div, args.style_scale_up))
        if aux_image is not None:
            st.aux_image = model.to_device(model.pil_to_image(aux_image.resize((w, h), Image.LANCZOS)))
        if output is not None:                                       # not the first scale
            # the next scale starts from the resampled AVERAGED iterate (:877-881)
            model.img = resize_f32(st.current_raw, (h, w))
            st.optimizer.set_params(model.img, resize=resize_f32)
        else:
            init = initial_image.resize((w, h), Image.LANCZOS) if initial_image is not None else None
            st.init_first_scale(h, w, init)
        if multiscale:
            # every style image becomes the list of its scaled copies; the Grams are computed once,
            # at the first scale, and kept (:754-755: styles are only reset without the flag)
            style_arrays = [[model.pil_to_image(v) for v in
                             style_multiscale_variants(im, multiscale[0], multiscale[1], args.div)]
                            for im in style_scaled]
        else:
            model.styles = []                                        # recomputed at every scale
            style_arrays = [model.pil_to_image(im) for im in style_scaled]
        iters = args.iterations[min(i, len(args.iterations) - 1)]
        output = st.transfer(iters, [model.pil_to_image(im) for im in content_scaled], style_arrays,
                             callback=(lambda **kw: callback(scale=i + 1, size=(h, w), **kw))
                             if callback else None)
    return output


def relaunch_multi_device(args, argv):
    devices = [d for d in args.devices if d >= 0]
    if len(devices) <= 1 or 'RANK' in os.environ:
        return
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=','.join(str(d) for d in devices))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', str(len(devices)), '--master-addr', '127.0.0.1', '--master-port',
           str(29500 + os.getpid() % 1000), sys.argv[0]] + list(argv)
    os.execvpe(cmd[0], cmd, env)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    import argparse
    state = argparse.Namespace()                 # the reference's STATE (:48): scale, step, steps, img_size
    args = config_system.parse_args(argv, state)
    from . import netdesc, weights
    net = netdesc.from_model(args.model)
    if args.list_layers:
        print('Layers:')
        for name, shape in net.shapes.items():
            print('% 25s %s' % (name, shape))
        return 0
    for flag, bad in (('--swt-weight', args.swt_weight),):
        if bad:
            raise SystemExit('%s is outside the scope of this engine (see DESIGN.md section 7)' % flag)
    relaunch_multi_device(args, argv)

    import torch
    import torch.distributed as dist
    from PIL import Image
    from .engine import TileEngine
    from .transfer import StyleTransfer
    rank, world, local = (int(os.environ.get(k, d)) for k, d in
                          (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    device = local if world > 1 else max(args.devices[0], 0)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', device))
    if args.weights == 'random':
        params = weights.he_normal(net)
    elif not os.path.exists(args.weights):
        # the reference fails hard here too (caffe.Net raises on a missing .caffemodel, :370)
        raise SystemExit('weights file %s not found (pass --weights random for He-normal random '
                         'weights)' % args.weights)
    elif args.weights.endswith('.npz'):
        params = weights.load_npz(args.weights)
    else:
        from .caffemodel import load_caffemodel
        params = load_caffemodel(args.weights, net)
    layer_weights = None
    if args.layer_weights:
        with open(args.layer_weights) as f:
            layer_weights = json.load(f)
    eng = TileEngine(net, params, mean=args.mean, device=device, precision=args.precision,
                     rank=rank, world=world)
    eng.init_comm()              # multi-device: the exchange step runs from the C ABI over NCCL
    st = StyleTransfer(eng, args, layer_weights)
    st.state = state
    if args.display != 'none' and rank == 0:
        print('note: --display %s is not available in this engine (no web / GUI display); running '
              'headless' % args.display)
    content = Image.open(args.content_image).convert('RGB')
    styles = [Image.open(p).convert('RGB') for p in args.style_images]
    init = Image.open(args.init_image).convert('RGB') if args.init_image else None
    aux = Image.open(args.aux_image).convert('RGB') if args.aux_image else None
    run = time.strftime('%Y-%m-%d_%H-%M-%S')
    stats = StatLogger(run + '_log.csv')
    state = {'n': 0, 't': None}

    def progress(step, update_size, loss, tv_loss, image, scale, size):
        now = time.perf_counter()
        dt = now - state['t'] if state['t'] is not None else 0.0
        state['t'], state['n'] = now, state['n'] + 1
        if rank == 0:
            print('Step %d, time: %.2f s, update: %.2f, loss: %e, tv: %.2f' %
                  (step, dt, update_size, loss, tv_loss), flush=True)          # :950
            stats.add(iteration=state['n'], scale=scale, step=step, content_h=size[0],
                      content_w=size[1], update_size=update_size, loss=loss, tv_norm=tv_loss)
            if args.save_every and state['n'] % args.save_every == 0:
                Image.fromarray(eng.get_image_array(image)).save(run + '_out_%04d.png' % state['n'])

    np.random.seed(args.seed)
    try:
        output = transfer_multiscale(st, args, [content], styles, init, aux, progress)
    except KeyboardInterrupt:
        output = st.current_raw
    if rank == 0:
        stats.dump()
        out_path = args.output_image or (run + '_out.png')
        Image.fromarray(eng.get_image_array(output)).save(out_path)
        print('Saved %s' % out_path)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
