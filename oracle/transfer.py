"""Oracle (TEST INFRASTRUCTURE): objective assembly and the single-scale optimisation loop.

Follows reference ``style_transfer.py``:
  * ``StyleTransfer.parse_weights``       :685-698  -> ``parse_weights``
  * ``StyleTransfer.eval_loss_and_grad``  :700-736  -> ``OracleTransfer.loss_and_grad``
        (SWT term :716-720 out of scope: weight defaults to 0, needs pywt)
  * ``StyleTransfer.transfer``            :738-830  -> ``OracleTransfer.run`` (non ``--jitter``)
  * first-scale initialisation            :883-901  -> ``OracleTransfer.init_first_scale``
  * ``CaffeModel.pil_to_image``           :388-393  -> ``to_params``
The scale ladder / Lanczos resampling between scales (:840-881) is a "next" item.
"""

from fractions import Fraction
from types import SimpleNamespace

import numpy as np

from .numeric import norm2, p_norm, tv_norm
from .optimizers import Adam, Lbfgs

DEFAULT_MEAN = (103.939, 116.779, 123.68)        # config_system.py:108-110


def default_args(**overrides):
    """The hot-path flags with the defaults of config_system.py:46-119."""
    args = dict(
        tile_size=512, optimizer='adam', step_size=15.0, step_decay=(0.05, 0.5), avg_window=20.0,
        content_weight=0.05, dd_weight=0.0, tv_weight=5.0, tv_power=2.0, p_weight=2.0,
        p_power=6.0, aux_weight=10.0, content_layers=['conv4_2'],
        style_layers=['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1'], dd_layers=[],
        mean=DEFAULT_MEAN, seed=0)
    args.update(overrides)
    return SimpleNamespace(**args)


def parse_weights(specs, master_weight):
    """['conv4_2:2', 'conv5_1'] -> names, {name: w * master / sum|w|}."""
    names, weights, total = [], {}, 0
    for spec in specs:
        name, _, w = spec.partition(':')
        names.append(name)
        weights[name] = float(Fraction(w)) if w else 1
        total += abs(weights[name])
    return names, {n: w * master_weight / total for n, w in weights.items()}


def to_params(rgb_hwc, mean=DEFAULT_MEAN):
    """uint8/float RGB [H,W,3] -> mean-subtracted BGR f32 [3,H,W]."""
    arr = np.float32(rgb_hwc).transpose((2, 0, 1))[::-1]
    return np.ascontiguousarray(arr - np.float32(mean).reshape(3, 1, 1))


def iter_stats(avg_img, old_img):
    """Per-iteration statistics of ``StyleTransfer.transfer`` (style_transfer.py:808-815): returns
    (update_size, tv_loss) and performs ``old_img[...] = avg_img`` like :810."""
    update_size = np.mean(abs(avg_img - old_img))
    old_img[...] = avg_img
    x_diff = avg_img - np.roll(avg_img, -1, axis=-1)
    y_diff = avg_img - np.roll(avg_img, -1, axis=-2)
    tv_loss = np.sqrt(np.mean(x_diff**2 + y_diff**2))
    return float(update_size), float(tv_loss)


def get_image_array(params, mean=DEFAULT_MEAN, bgr=True):
    """``CaffeModel.get_image`` (style_transfer.py:378-386) up to the PIL wrapper: uint8 HxWx3."""
    arr = params + np.float32(mean).reshape((3, 1, 1))
    if bgr:
        arr = arr[::-1]
    arr = arr.transpose((1, 2, 0))
    return np.uint8(np.clip(arr, 0, 255))


class OracleTransfer:
    def __init__(self, model, args, layer_weights=None):
        self.model = model
        self.args = args
        self.layer_weights = {layer: 1.0 for layer in model.layers() + ['data']}
        self.layer_weights.update(layer_weights or {})
        self.aux_image = None
        self.optimizer = None
        self.mean = np.float32(args.mean).reshape(3, 1, 1)

    def loss_and_grad(self, img, sc_args):
        a = self.args
        old_img, self.model.img = self.model.img, img
        lw = self.layer_weights['data']
        loss, grad = self.model.sc_grad(*sc_args)
        if a.tv_weight:                                                        # :710-713
            tv_loss, tv_grad = tv_norm(img / np.float32(127.5), beta=a.tv_power)
            loss += lw * a.tv_weight * tv_loss
            grad += np.float32(lw * a.tv_weight) * tv_grad
        if a.p_weight:                                                         # :723-727
            p_loss, p_grad = p_norm((img + self.mean - np.float32(127.5)) / np.float32(127.5),
                                    p=a.p_power)
            loss += lw * a.p_weight * p_loss
            grad += np.float32(lw * a.p_weight) * np.float32(p_grad)
        if self.aux_image is not None:                                         # :730-733
            aux_grad = (img - self.aux_image) / np.float32(127.5)
            loss += lw * a.aux_weight * norm2(aux_grad)
            grad += np.float32(lw * a.aux_weight) * aux_grad
        self.model.img = old_img
        return loss, grad

    def init_first_scale(self, h, w, init_rgb=None):
        """:883-901.  Without an init image the RNG draw is uniform(0,255,(h,w,3))."""
        a = self.args
        biased_g1 = init_rgb is not None
        if init_rgb is None:
            init_rgb = np.random.uniform(0, 255, size=(h, w, 3))
        self.model.img = to_params(init_rgb, a.mean)
        if a.optimizer == 'adam':
            self.optimizer = Adam(self.model.img, step_size=a.step_size,
                                  bp1=1 - (1 / a.avg_window), decay=a.step_decay[0],
                                  power=a.step_decay[1], biased_g1=biased_g1)
        else:
            self.optimizer = Lbfgs(self.model.img)

    def run(self, iterations, content_imgs, style_imgs, callback=None):
        """One scale of ``transfer``; returns the (averaged, un-rolled) raw iterate."""
        a, model = self.args, self.model
        params = model.img
        c_layers, c_weight = parse_weights(a.content_layers, a.content_weight)
        s_layers, s_weight = parse_weights(a.style_layers, 1)
        d_layers, d_weight = parse_weights(a.dd_layers, a.dd_weight)
        jitter = bool(getattr(a, 'jitter', False))
        model.contents, model.styles = [], []
        if jitter:                                                              # :757-759
            model.preprocess([], style_imgs, [], s_layers, a.tile_size)
        else:
            model.preprocess(content_imgs, style_imgs, c_layers, s_layers, a.tile_size)
        model.publish()
        model.img = params
        avg_img = None
        for step in range(1, iterations + 1):
            js, _ = model.layer_info([l for l in reversed(model.layers()) if l in c_layers][0])
            if jitter:                                                          # :780-782
                js = 1
                model.contents = []
            img_size = np.array(model.img.shape[-2:])
            xy = np.int32(np.random.uniform(-0.5, 0.5, size=2) * img_size) // js   # :784
            model.roll(xy, jitter_scale=js)
            self.optimizer.roll(xy * js)
            xy_ = xy
            if jitter:                                                          # :788-794
                model.preprocess(content_imgs, [], c_layers, [], a.tile_size, roll=xy)
                model.img = params
                model.publish()
                xy_ = np.asarray((0, 0))
            sc_args = (xy_ * js, c_layers, s_layers, d_layers, self.layer_weights, c_weight,
                       s_weight, d_weight, a.tile_size)
            avg_img, loss = self.optimizer.update(lambda p: self.loss_and_grad(p, sc_args))
            model.roll(-xy, jitter_scale=js)
            self.optimizer.roll(-xy * js)
            if callback is not None:
                callback(step=step, loss=loss, image=avg_img)
        return avg_img
