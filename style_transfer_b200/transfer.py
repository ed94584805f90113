"""Objective assembly and the optimisation loop: mirror of the reference's ``StyleTransfer``
(style_transfer.py:664-909) driving the CUDA tile engine.

The loop structure, the order of draws from the global numpy RNG (``np.random.uniform`` at :889,
:475 and :784) and the ``//jitter_scale`` quantisation of the roll are kept exactly, because the
result depends on them.  Per-iteration statistics (:808-817) are computed lazily: only when the
caller's callback asks for them is anything copied off the device.
"""

import ctypes as C
from fractions import Fraction

import numpy as np
import torch

from . import _lib
from .optimizers import AdamOptimizer, LBFGSOptimizer


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ffloat(s):
    """Parses fractional or floating point input strings (config_system.py:27-29)."""
    return float(Fraction(s))


def parse_weights(args, master_weight):
    """Parses name[:number] pairs into a normalized dict of weights (style_transfer.py:685-698)."""
    names, weights, total = [], {}, 0
    for arg in args:
        name, _, w = arg.partition(':')
        names.append(name)
        weights[name] = ffloat(w) if w else 1
        total += abs(weights[name])
    return names, {name: weight * master_weight / total for name, weight in weights.items()}


def default_args(**overrides):
    """The hot-path flags with the defaults of the reference's flag table
    (config_system.py:46-119); the CLI front end fills the same namespace from argv."""
    from types import SimpleNamespace
    args = dict(
        size=256, min_size=182, tile_size=512, devices=[-1], iterations=[200, 100],
        optimizer='adam', step_size=15.0, step_decay=[0.05, 0.5], avg_window=20.0,
        content_weight=0.05, dd_weight=0.0, tv_weight=5.0, tv_power=2.0, p_weight=2.0,
        p_power=6.0, aux_weight=10.0, content_layers=['conv4_2'],
        style_layers=['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1'], dd_layers=[],
        model='vgg19.prototxt', weights='vgg19.caffemodel', mean=(103.939, 116.779, 123.68),
        seed=0, div=1, jitter=False)
    args.update(overrides)
    return SimpleNamespace(**args)


def scale_ladder(size, min_size):
    """Sizes from --size down by sqrt(2) while >= --min-size, smallest first (:840-846, :854)."""
    sizes = [size]
    while True:
        size = round(size / np.sqrt(2))
        if size < min_size:
            break
        sizes.append(size)
    return list(reversed(sizes))


class StyleTransfer:
    """Performs style transfer on the device.  ``args`` carries the reference's flag names
    (config_system.py:46-119)."""

    def __init__(self, model, args, layer_weights=None):
        self.model = model
        self.args = args
        self.layer_weights = {layer: 1.0 for layer in model.layers() + ['data']}
        if layer_weights:
            self.layer_weights.update(layer_weights)
        self.aux_image = None            # CUDA f32[3,H,W] in pil_to_image format
        self.state = None                # the reference's STATE namespace (:48, :740-746), if any
        self.current_raw = None
        self.optimizer = None
        self._mean = (C.c_float * 3)(*[float(m) for m in np.ravel(model.mean)])

    # ---- objective ------------------------------------------------------------------------------
    def eval_loss_and_grad(self, img, sc_grad_args):
        """Returns the summed loss and gradient (:700-736) as CUDA tensors (float64[1], f32[3,H,W]).
        ``sc_grad_args`` = (roll_xy_pixels, content_layers, style_layers, dd_layers, layer_weights,
        content_weight, style_weight, dd_weight, tile_size) as in the reference minus the pool."""
        a = self.args
        lw = self.layer_weights['data']
        tv_w = lw * a.tv_weight if a.tv_weight else 0.0
        p_w = lw * a.p_weight if a.p_weight else 0.0
        aux_w = lw * a.aux_weight if self.aux_image is not None else 0.0
        reg = None
        if tv_w or p_w or self.aux_image is not None:
            reg = (self._mean, tv_w, a.tv_power, p_w, a.p_power, self.aux_image, aux_w)
        # the regularisers ride on the pass that stitches the gradient tiles (st_unpack_regularize)
        return self.model.eval_sc_grad(*sc_grad_args, img=img, regularizers=reg)

    # ---- first-scale initialisation (:882-901) ---------------------------------------------------
    def init_first_scale(self, h, w, initial_image=None):
        a = self.args
        biased_g1 = initial_image is not None
        if initial_image is None:
            initial_image = np.random.uniform(0, 255, size=(h, w, 3))       # RNG draw (:889)
        self.model.set_image(initial_image)
        if a.optimizer == 'adam':
            self.optimizer = AdamOptimizer(
                self.model.img, step_size=a.step_size, bp1=1 - (1 / a.avg_window),
                decay=a.step_decay[0], power=a.step_decay[1], biased_g1=biased_g1)
        elif a.optimizer == 'lbfgs':
            self.optimizer = LBFGSOptimizer(self.model.img)
        else:
            raise ValueError('unknown optimizer %r' % a.optimizer)

    # ---- one scale (:738-830) ---------------------------------------------------------------------
    def prepare(self, content_images, style_images):
        """Preprocessing part of ``transfer`` (:748-766): weights, targets, hand-off to the GPU."""
        a, model = self.args, self.model
        self.c_layers, self.c_weight = parse_weights(a.content_layers, a.content_weight)
        self.s_layers, self.s_weight = parse_weights(a.style_layers, 1)
        self.d_layers, self.d_weight = parse_weights(a.dd_layers, a.dd_weight)
        model.contents = []
        if not getattr(a, 'style_multiscale', None):                 # (:754-755)
            model.styles = []
        self.jitter = bool(getattr(a, 'jitter', False))
        self._jitter_primed = False
        if self.jitter:
            # --jitter (:757-759): only the styles now; the content features are recomputed from the
            # rolled content image(s) in every iteration
            self._content_images = [model.to_device(c) for c in content_images]
            model.preprocess_images([], style_images, [], self.s_layers, a.tile_size)
        else:
            model.preprocess_images(content_images, style_images, self.c_layers, self.s_layers,
                                    a.tile_size)
        model.set_contents_and_styles()
        deepest_content = [l for l in reversed(model.layers()) if l in self.c_layers]
        self.jitter_scale = model.layer_info(deepest_content[0])[0] if deepest_content else 1
        if self.jitter:
            self.jitter_scale = 1                                    # (:780-781)

    def step(self):
        """One pass of the loop body (:777-806): draw the roll, update, roll back.  Returns
        (averaged iterate, loss) on the device."""
        a, model = self.args, self.model
        if getattr(self, 'jitter', False):
            return self._step_jitter()
        js = self.jitter_scale
        img_size = np.array(model.img.shape[-2:])
        xy = np.int32(np.random.uniform(-0.5, 0.5, size=2) * img_size) // js       # (:784)
        model.roll(xy, jitter_scale=js)
        self.optimizer.roll(xy * js)
        sc_args = (xy * js, self.c_layers, self.s_layers, self.d_layers, self.layer_weights,
                   self.c_weight, self.s_weight, self.d_weight, a.tile_size)
        avg_img, loss = self.optimizer.update(
            lambda params: self.eval_loss_and_grad(params, sc_args))
        model.roll(-xy, jitter_scale=js)
        self.optimizer.roll(-xy * js)
        return avg_img, loss

    def _step_jitter(self):
        """One iteration under ``--jitter`` (:778-797): the roll is drawn at PIXEL granularity
        (jitter_scale = 1), the content features are recomputed from the content image(s) rolled by
        it (one pass, no RNG draw), and the objective is evaluated with no feature roll.  The
        reference rolls the iterate and the optimizer state in place and back; here the objective
        sees a rolled COPY of the parameters and its gradient is rolled back, which is the same
        thing for element-wise optimizer updates (the un-rolled frame never moves).  The aux image
        stays un-rolled against the rolled iterate, as in the reference (:730-733)."""
        a, model = self.args, self.model
        img_size = np.array(model.img.shape[-2:])
        xy = np.int32(np.random.uniform(-0.5, 0.5, size=2) * img_size) // 1        # (:784)
        sh = (int(xy[1]), int(xy[0]))        # roll2: xy[0] along the width, xy[1] along the height
        model.contents = []
        model.preprocess_images([torch.roll(c, sh, dims=(-2, -1)) for c in self._content_images], [],
                                self.c_layers, [], a.tile_size, content_passes=1)
        # only the content features change: overwritten in place, the style Grams stay on the device
        model.set_contents_and_styles(contents_only=self._jitter_primed)
        self._jitter_primed = True
        sc_args = (np.zeros(2, dtype=np.int64), self.c_layers, self.s_layers, self.d_layers,
                   self.layer_weights, self.c_weight, self.s_weight, self.d_weight, a.tile_size)

        def opfunc(params):
            loss, grad = self.eval_loss_and_grad(torch.roll(params, sh, dims=(-2, -1)), sc_args)
            return loss, torch.roll(grad, (-sh[0], -sh[1]), dims=(-2, -1))
        return self.optimizer.update(opfunc)

    def transfer(self, iterations, content_images, style_images, callback=None):
        """Performs style transfer at the current scale; returns the averaged raw iterate."""
        st = self.state
        if st is not None:                       # STATE bookkeeping read by callable config values
            st.scale = st.scale + 1 if 'scale' in st else 0                  # (:740-743)
            st.step, st.steps = 0, iterations
            st.img_size = tuple(self.model.img.shape[1:])
        self.prepare(content_images, style_images)
        old_img = self.model.img.clone()
        stats = torch.zeros(2, dtype=torch.float64, device=old_img.device)
        avg_img = None
        for step in range(1, iterations + 1):
            if st is not None:
                st.step = step - 1                                           # (:772)
            avg_img, loss = self.step()
            if callback is not None:
                update_size, tv_loss = self.iter_stats(avg_img, old_img, stats)
                callback(step=step, update_size=update_size, loss=float(loss), tv_loss=tv_loss,
                         image=avg_img)
            self.current_raw = avg_img
        return avg_img

    @staticmethod
    def iter_stats_async(avg_img, old_img, stats):
        """The update-size and total-variation sums of :808-815 in one device pass (st_iter_stats),
        which also performs ``old_img[...] = avg_img``: ``stats`` (two device doubles) receives
        sum |avg - old| and the sum of the squared periodic differences.  Nothing is synchronised."""
        h, w = avg_img.shape[-2:]
        _lib.call('st_iter_stats', _ptr(avg_img), _ptr(old_img), h, w, _ptr(stats), _stream())
        return stats

    def output_step(self, avg_img, old_img, stats, picture=True):
        """The loop's output step (:808-821) in ONE device pass (st_output_step): the two statistic
        sums into ``stats``, ``old_img[...] = avg_img`` and the uint8 RGB picture of ``avg_img``
        (CaffeModel.get_image :378-386; returned as a CUDA tensor [H,W,3], or None)."""
        h, w = avg_img.shape[-2:]
        pic = torch.empty((h, w, 3), dtype=torch.uint8, device=avg_img.device) if picture else None
        _lib.call('st_output_step', _ptr(avg_img), _ptr(old_img), h, w, self._mean,
                  1 if self.model.bgr else 0, _ptr(stats), _ptr(pic), _stream())
        return pic

    @staticmethod
    def iter_stats(avg_img, old_img, stats):
        """(update_size, tv_loss) of :808-815 as Python floats (copies two doubles to the host)."""
        s = StyleTransfer.iter_stats_async(avg_img, old_img, stats).cpu().numpy()
        n = float(avg_img.numel())
        return float(s[0] / n), float(np.sqrt(s[1] / n))
