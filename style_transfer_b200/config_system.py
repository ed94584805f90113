"""Flag table of the reference (config_system.py:42-119): same names, short forms, types and
defaults, so command lines written for the reference run unchanged.  Flags of subsystems that are
out of scope here (web / GUI display, SWT regulariser, Caffe path) are accepted and ignored with a
note; ``--precision`` is the one addition (tensor-core modes vs. the fp32 SIMT parity mode).

Precedence as in the reference (:121-136): defaults < ``config.py`` beside the entry script (the
reference looks beside its ``config_system.py``, which sits beside ``style_transfer.py``; :14) <
command line (only values that differ from the default) < ``--config FILE``.  A config file is
Python source executed with ``CONFIG_GLOBALS`` = {detect_devices, math, np} visible (:187-194);
every name it defines becomes an argument.  A callable value is called with the state object each
time it is read (``AutocallNamespace``, :151-184), so a config can say
``step_size = lambda st: 15 if st.scale < 3 else 10``.
"""

import argparse
from fractions import Fraction
import math
import os
from pathlib import Path
import re
import subprocess

import numpy as np

# beside style_transfer.py (the drop-in entry script at the repository root)
CONFIG_PY = Path(__file__).resolve().parent.parent / 'config.py'


def detect_devices():
    """GPU indices reported by ``nvidia-smi -L``, or [-1] (config_system.py:17-24)."""
    try:
        gpu_list = subprocess.run(['nvidia-smi', '-L'], stdout=subprocess.PIPE, check=True,
                                  universal_newlines=True)
        gpus = [int(g) for g in re.findall(r'^GPU (\d+)', gpu_list.stdout, re.M)]
        return gpus if gpus else [-1]
    except (subprocess.CalledProcessError, FileNotFoundError):
        return [-1]


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
    arg('--display', default='browser', choices=['browser', 'gui', 'none'],
        help='accepted for compatibility; this engine has no web / GUI display (always "none")')
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
    arg('--precision', default='fp16', choices=['bf16', 'fp16', 'tc32', 'fp32'],
        help='bf16 / fp16 (fp16 forward, bf16 backward): tcgen05 tensor cores, 16-bit operands; tc32: '
        'tensor cores with split fp16 hi+lo operands (fp32-class results); fp32: exact SIMT parity mode')
    return p


class ValuePlaceholder:
    """What a callable argument evaluates to while the state it asks for does not exist yet
    (config_system.py:147-148)."""


class AutocallNamespace:
    """Argument namespace whose callable values are called with ``state_obj`` on every read
    (config_system.py:151-184)."""

    def __init__(self, state_obj, **kwargs):
        self.state_obj = state_obj
        self.ns = argparse.Namespace(**kwargs)

    def __getattr__(self, name):
        value = getattr(self.ns, name)
        if callable(value):
            try:
                return value(self.state_obj)
            except AttributeError:
                return ValuePlaceholder()
        return value

    def __setattr__(self, name, value):
        if name in ('state_obj', 'ns'):
            object.__setattr__(self, name, value)
            return
        setattr(self.ns, name, value)

    def __iter__(self):
        yield from vars(self.ns)

    def __contains__(self, key):
        return key in self.ns

    def __repr__(self):
        return 'Autocall' + repr(self.ns)


CONFIG_GLOBALS = dict(detect_devices=detect_devices, math=math, np=np)


def eval_config(config_file):
    """Executes a config file with CONFIG_GLOBALS visible; returns the names it defines
    (config_system.py:190-194)."""
    config_file = Path(config_file)
    code = compile(config_file.read_text(), config_file.name, 'exec')
    locs = {}
    exec(code, dict(CONFIG_GLOBALS), locs)
    return locs


def parse_args(argv=None, state_obj=None):
    parser = build_parser()
    defaults = vars(parser.parse_args([]))
    config_args = eval_config(CONFIG_PY) if CONFIG_PY.exists() else {}
    sysv = vars(parser.parse_args(argv))
    config2_args = eval_config(sysv['config']) if sysv['config'] else {}
    args = dict(defaults)
    args.update(config_args)
    for k, v in sysv.items():
        if defaults[k] != v:
            args[k] = v
    args.update(config2_args)
    ns = AutocallNamespace(state_obj, **args)
    if ns.debug:
        os.environ['DEBUG'] = '1'
    if not ns.list_layers and (not ns.content_image or not ns.style_images):
        parser.print_help()
        raise SystemExit(1)
    return ns
