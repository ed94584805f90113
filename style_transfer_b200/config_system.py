"""Flag table of the reference (config_system.py:42-119): same names, short forms, types and
defaults, so command lines written for the reference run unchanged.  Flags of subsystems that are
out of scope here (web / GUI display, SWT regulariser, Caffe path) are accepted and ignored with a
note; ``--precision`` is the one addition (fp32 parity mode vs. bf16 tensor-core mode).

Precedence as in the reference (:121-136): defaults < ``config.py`` in the working directory <
command line (only values that differ from the default) < ``--config FILE``.  A config file is
Python source executed with the already-parsed flags visible; every public name it defines that
matches a flag overrides it.
"""

import argparse
from fractions import Fraction
from pathlib import Path

CONFIG_PY = Path('config.py')


def ffloat(s):
    """Parses fractional or floating point input strings (config_system.py:27-29)."""
    return float(Fraction(s))


def build_parser():
    p = argparse.ArgumentParser(description='Neural style transfer on B200 (libstyle_b200).',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    arg = p.add_argument
    arg('--content-image', '-ci', help='the content image')
    arg('--style-images', '-si', nargs='+', default=[], metavar='STYLE_IMAGE', help='the style images')
    arg('--output-image', '-oi', help='the output image')
    arg('--init-image', '-ii', metavar='IMAGE', help='the initial image')
    arg('--aux-image', '-ai', metavar='IMAGE', help='the auxiliary image')
    arg('--config', type=Path, help='a Python source file containing configuration options')
    arg('--list-layers', action='store_true', help="list the model's layers")
    arg('--caffe-path', help='ignored (no Caffe in this engine)')
    arg('--devices', nargs='+', metavar='DEVICE', type=int, default=[-1],
        help='GPU device numbers to use (one rank per device; -1 = device 0)')
    arg('--iterations', '-i', nargs='+', type=int, default=[200, 100], help='the number of iterations')
    arg('--size', '-s', type=int, default=256, help='the output size')
    arg('--min-size', type=int, default=182, help="the minimum scale's size")
    arg('--style-scale', '-ss', type=ffloat, default=1, help='the style scale factor')
    arg('--max-style-size', type=int, help='the maximum style size')
    arg('--style-scale-up', default=False, action='store_true', help='allow scaling style images up')
    arg('--style-multiscale', '-sm', type=int, nargs=2, metavar=('MIN_SCALE', 'MAX_SCALE'),
        default=None, help='average the style Gram matrices over copies of each style image '
        'scaled from MAX_SCALE down by sqrt(2) to MIN_SCALE (computed once, at the first scale)')
    arg('--tile-size', type=int, default=512, help='the maximum rendering tile size')
    arg('--optimizer', '-o', default='adam', choices=['adam', 'lbfgs'], help='the optimizer to use')
    arg('--step-size', '-st', type=ffloat, default=15, help='the initial step size for Adam')
    arg('--step-decay', '-sd', nargs=2, metavar=('DECAY', 'POWER'), type=ffloat, default=[0.05, 0.5],
        help='on step i, divide step_size by (1 + DECAY * i)^POWER')
    arg('--avg-window', type=ffloat, default=20, help='the iterate averaging window size')
    arg('--layer-weights', help='a json file containing per-layer weight scaling factors')
    arg('--content-weight', '-cw', type=ffloat, default=0.05, help='the content image factor')
    arg('--dd-weight', '-dw', type=ffloat, default=0, help='the Deep Dream factor')
    arg('--tv-weight', '-tw', type=ffloat, default=5, help='the TV smoothing factor')
    arg('--tv-power', '-tp', metavar='BETA', type=ffloat, default=2, help='the TV smoothing exponent')
    arg('--swt-weight', '-ww', metavar='WEIGHT', type=ffloat, default=0,
        help='the SWT smoothing factor (only 0 is supported)')
    arg('--swt-wavelet', '-wt', metavar='WAVELET', default='haar', help='ignored')
    arg('--swt-levels', '-wl', metavar='LEVELS', default=1, type=int, help='ignored')
    arg('--swt-power', '-wp', metavar='P', default=2, type=ffloat, help='ignored')
    arg('--p-weight', '-pw', type=ffloat, default=2, help='the p-norm regularizer factor')
    arg('--p-power', '-pp', metavar='P', type=ffloat, default=6, help='the p-norm exponent')
    arg('--aux-weight', '-aw', type=ffloat, default=10, help='the auxiliary image factor')
    arg('--content-layers', nargs='*', default=['conv4_2'], metavar='LAYER',
        help='the layers to use for content')
    arg('--style-layers', nargs='*', metavar='LAYER',
        default=['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1'],
        help='the layers to use for style')
    arg('--dd-layers', nargs='*', metavar='LAYER', default=[], help='the layers to use for Deep Dream')
    arg('--port', '-p', type=int, default=8000, help='ignored (no web interface)')
    arg('--display', default='none', choices=['browser', 'gui', 'none'], help='only "none" is supported')
    arg('--browser', default=None, help='ignored')
    arg('--model', default='vgg19.prototxt', help='the deploy.prototxt of the model to use')
    arg('--weights', default='vgg19.caffemodel',
        help='the .caffemodel of the model (or an .npz written by weights.save_npz, or "random")')
    arg('--mean', nargs=3, metavar=('B_MEAN', 'G_MEAN', 'R_MEAN'), type=float,
        default=(103.939, 116.779, 123.68), help='the per-channel means of the model (BGR order)')
    arg('--save-every', metavar='N', type=int, default=0, help='save the image every n steps')
    arg('--seed', type=int, default=0, help='the random seed')
    arg('--div', metavar='FACTOR', type=int, default=1, help='ensure all images are divisible by FACTOR')
    arg('--jitter', action='store_true',
        help='roll by whole pixels and recompute the content features from the rolled content '
        'image in every iteration (slower; avoids the feature-grid quantisation of the roll)')
    arg('--debug', action='store_true', help='enable debug messages')
    arg('--precision', default='fp16', choices=['bf16', 'fp16', 'fp32'],
        help='bf16 / fp16 (fp16 forward, bf16 backward): tcgen05 tensor cores; fp32: exact SIMT parity mode')
    return p


def eval_config(path, visible):
    """Executes a config file; returns the public names it defines (config_system.py:151-178)."""
    scope = dict(visible)
    before = set(scope)
    exec(compile(Path(path).read_text(), str(path), 'exec'), scope)
    return {k: v for k, v in scope.items()
            if not k.startswith('_') and (k not in before or scope[k] is not visible.get(k))}


def parse_args(argv=None):
    parser = build_parser()
    defaults = vars(parser.parse_args([]))
    args = dict(defaults)
    known = set(defaults)
    if CONFIG_PY.exists():
        args.update({k: v for k, v in eval_config(CONFIG_PY, args).items() if k in known})
    sysv = vars(parser.parse_args(argv))
    for k, v in sysv.items():
        if defaults[k] != v:
            args[k] = v
    if sysv['config']:
        args.update({k: v for k, v in eval_config(sysv['config'], args).items() if k in known})
    ns = argparse.Namespace(**args)
    if not ns.list_layers and (not ns.content_image or not ns.style_images):
        parser.print_help()
        raise SystemExit(1)
    return ns
