"""Oracle (TEST INFRASTRUCTURE): the per-tile operator and the tile loop, restated on OracleNet.

Follows reference ``style_transfer.py``:
  * ``CaffeModel.eval_features_tile``  :421-427   -> ``OracleModel.features_tile``
  * ``CaffeModel.eval_features_once``  :429-464   -> ``OracleModel.features_once``
  * ``CaffeModel.prepare_features``    :466-486   -> ``OracleModel.prepare_features``
  * ``CaffeModel.preprocess_images``   :488-554   -> ``OracleModel.preprocess`` (arrays, no PIL;
        the ``--style-multiscale`` branch :501-527 is out of scope)
  * ``CaffeModel.eval_sc_grad_tile``   :556-612   -> ``OracleModel.sc_grad_tile``
  * ``CaffeModel.eval_sc_grad``        :614-645   -> ``OracleModel.sc_grad``
  * ``CaffeModel.roll/roll_features``  :647-661   -> ``OracleModel.roll`` / ``roll_features``
  * ``CaffeModel.layer_info``          :415-419   -> ``OracleModel.layer_info``
The fork()/queue/shm transport (:169-337) carries no arithmetic: tiles are simply evaluated one
after the other here, and summed in row-major request order.
"""

import numpy as np

from . import caffe_ops as ops
from .caffe_net import OracleNet, model_layers
from .numeric import gram_lower, norm2, normalize_, roll2_, symm_times


def tile_grid(img_size, tile_size):
    """Tile boxes [(start(y,x), end(y,x))] in row-major order (style_transfer.py:431-450,
    619-631): n = (size-1)//tile + 1 tiles of size//n, the last row/column absorbing the rest."""
    img_size = np.array(img_size)
    ntiles = (img_size - 1) // tile_size + 1
    tile = img_size // ntiles
    boxes = []
    for y in range(ntiles[0]):
        for x in range(ntiles[1]):
            start = np.array([y, x]) * tile
            end = start + tile
            if y == ntiles[0] - 1:
                end[0] = img_size[0]
            if x == ntiles[1] - 1:
                end[1] = img_size[1]
            boxes.append((start, end))
    return boxes


class OracleModel:
    def __init__(self, model_name, params, compute_weight_grads=False):
        self.layer_list = model_layers(model_name)
        self.net = OracleNet(self.layer_list, params, compute_weight_grads)
        self.contents = []      # master copy: list of {layer: f32[C,Hf,Wf]} (ContentData.features)
        self.styles = []        # master copy: list of {layer: f32[C,C] lower} (StyleData.grams)
        self.w_contents = []    # the workers' private copies, made by publish()
        self.w_styles = []
        self.img = None
        # Blob shapes at the canonical 224x224 input, as tabulated at style_transfer.py:1030-1073.
        self.shapes = {}
        size = {'data': 224}
        for kind, _, bottom, top, extra in self.layer_list:
            if kind == 'conv':
                size[top] = size[bottom]
                self.shapes[top] = (extra[1], size[top], size[top])
            elif kind == 'pool':
                size[top] = ops.pooled_size(size[bottom])
                self.shapes[top] = (self.shapes[bottom][0], size[top], size[top])
        self.last_layer = list(self.shapes)[-1]

    def layers(self):
        return list(self.shapes)

    def layer_info(self, layer):
        return 224 // self.shapes[layer][1], self.shapes[layer][0]

    # ---- feature extraction -----------------------------------------------------------------
    def features_tile(self, img, layers):
        self.net.set_input(img)
        self.net.forward(end=self.last_layer)                 # always to the last layer (:425)
        return {layer: self.net.data[layer].copy() for layer in layers}

    def features_once(self, layers, tile_size=512):
        img_size = np.array(self.img.shape[-2:])
        feats = {}
        for layer in layers:
            scale, ch = self.layer_info(layer)
            feats[layer] = np.zeros((ch,) + tuple(np.int32(np.ceil(img_size / scale))), np.float32)
        for start, end in tile_grid(img_size, tile_size):
            tile = self.features_tile(self.img[:, start[0]:end[0], start[1]:end[1]], layers)
            for layer, feat in tile.items():
                s = start // self.layer_info(layer)[0]
                e = s + np.array(feat.shape[-2:])
                feats[layer][:, s[0]:e[0], s[1]:e[1]] = feat
        return feats

    def prepare_features(self, layers, tile_size=512, passes=10):
        img_size = np.array(self.img.shape[-2:])
        if max(*img_size) <= tile_size:
            passes = 1
        feats = {}
        for i in range(passes):
            xy = np.array((0, 0))
            if i > 0:
                xy = np.int32(np.random.uniform(size=2) * img_size) // 32       # RNG draw (:475)
            self.roll(xy)
            self.roll_features(feats, xy)
            once = self.features_once(layers, tile_size)
            for layer in layers:
                if i == 0:
                    feats[layer] = once[layer] / passes
                else:
                    feats[layer] += np.float32(1 / passes) * once[layer]
            self.roll(-xy)
            self.roll_features(feats, -xy)
        return feats

    def preprocess(self, content_imgs, style_imgs, content_layers, style_layers, tile_size=512,
                   roll=None):
        """content_imgs / style_imgs: lists of preprocessed f32[3,H,W] arrays.  A style entry may be
        a list of arrays: the scaled copies of ``--style-multiscale`` (style_transfer.py:501-524),
        each adding its Gram matrices and counting once in the average."""
        if not self.styles:
            grams, count = {}, 0
            for entry in style_imgs:
                for img in (entry if isinstance(entry, (list, tuple)) else [entry]):
                    self.img = img.copy()
                    feats = self.prepare_features(style_layers, tile_size, passes=1)
                    for layer in feats:
                        g = gram_lower(feats[layer])
                        grams[layer] = g if layer not in grams else grams[layer] + g
                    count += 1
            for g in grams.values():
                g /= count
            self.styles.append(grams)
        # ``roll`` given = the per-iteration preprocessing of --jitter (:526, :545-552): the image is
        # rolled before its features are taken, in ONE pass
        for img in content_imgs:
            self.img = img.copy()
            if roll is not None:
                roll2_(self.img, roll)
            self.contents.append(self.prepare_features(content_layers, tile_size,
                                                       passes=10 if roll is None else 1))

    def publish(self):
        """``TileWorkerPool.set_contents_and_styles`` (:309-332): every worker receives its own
        COPY of the features / Grams; later rolls of the master's copy do not reach them."""
        self.w_contents = [{k: v.copy() for k, v in c.items()} for c in self.contents]
        self.w_styles = [{k: v.copy() for k, v in g.items()} for g in self.styles]

    # ---- loss + gradient ----------------------------------------------------------------------
    def ordered_layers(self, *layer_sets):
        """Deepest-first list of the requested layers (style_transfer.py:231-233)."""
        wanted = set().union(*layer_sets)
        return [l for l in reversed(self.layers()) if l in wanted]

    def sc_grad_tile(self, img, start, layers, content_layers, style_layers, dd_layers,
                     layer_weights, content_weight, style_weight, dd_weight):
        net = self.net
        net.set_input(img)
        loss = 0
        net.forward(end=layers[0])
        ops.relu_forward_(net.data[layers[0]])                       # manual ReLU (:567)
        for layer in layers:
            net.diff[layer][...] = 0                                 # :564-565
        for i, layer in enumerate(layers):
            lw = layer_weights[layer]
            scale, _ = self.layer_info(layer)
            data = net.data[layer]
            s0 = np.asarray(start) // scale
            e0 = s0 + np.array(data.shape[-2:])
            if layer in content_layers:                              # :575-580
                for content in self.w_contents:
                    target = content[layer][:, s0[0]:e0[0], s0[1]:e0[1]]
                    c_grad = data - target
                    loss += lw * content_weight[layer] * norm2(c_grad)
                    net.diff[layer] += np.float32(lw * content_weight[layer]) * normalize_(c_grad)
            if layer in style_layers:                                # :582-593
                for grams in self.w_styles:
                    n = data.shape[0]
                    gram_diff = gram_lower(data) - grams[layer]
                    s_grad = symm_times(gram_diff, data.reshape(n, -1)).reshape(data.shape)
                    loss += lw * style_weight[layer] * norm2(gram_diff) / len(self.w_styles)
                    net.diff[layer] += np.float32(lw * style_weight[layer] / len(self.w_styles)) * \
                        normalize_(s_grad)
            if layer in dd_layers:                                   # :602-604
                loss -= lw * dd_weight[layer] * norm2(data)
                net.diff[layer] += np.float32(-lw * dd_weight[layer]) * normalize_(data)
            if i + 1 == len(layers):                                 # :607-610
                net.backward(start=layer)
            else:
                net.backward(start=layer, end=layers[i + 1])
        return loss, net.diff['data']

    def sc_grad(self, roll, content_layers, style_layers, dd_layers, layer_weights,
                content_weight, style_weight, dd_weight, tile_size):
        """``roll`` is the pixel roll already applied to self.img; the worker applies the same
        roll to its content features around the tile evaluation (:234, :240)."""
        loss = 0
        grad = np.zeros_like(self.img)
        layers = self.ordered_layers(content_layers, style_layers, dd_layers)
        for start, end in tile_grid(self.img.shape[-2:], tile_size):
            tile = np.ascontiguousarray(self.img[:, start[0]:end[0], start[1]:end[1]])
            self.roll_features_all(self.w_contents, roll, 1)
            loss_tile, grad_tile = self.sc_grad_tile(
                tile, start, layers, content_layers, style_layers, dd_layers, layer_weights,
                content_weight, style_weight, dd_weight)
            self.roll_features_all(self.w_contents, -np.asarray(roll), 1)
            loss += loss_tile
            grad[:, start[0]:end[0], start[1]:end[1]] = grad_tile
        return loss, grad

    # ---- roll ---------------------------------------------------------------------------------
    def roll_features(self, feats, xy, jitter_scale=32):
        xy = np.asarray(xy) * jitter_scale
        for layer, feat in feats.items():
            roll2_(feat, xy // self.layer_info(layer)[0])
        return feats

    def roll_features_all(self, contents, xy, jitter_scale):
        for content in contents:
            self.roll_features(content, xy, jitter_scale)

    def roll(self, xy, jitter_scale=32):
        """Master-side roll (:657-661): its own feature copies and the image."""
        self.roll_features_all(self.contents, xy, jitter_scale)
        roll2_(self.img, np.asarray(xy) * jitter_scale)
