"""Host-side mirror of the reference's ``CaffeModel`` operator interface on top of libstyle_b200.

Same method names, argument meaning and error behaviour as style_transfer.py:356-661, with three
deliberate differences that the B200 design needs:

* tensors live in HBM (torch CUDA tensors are used purely as storage; host numpy arrays are
  accepted everywhere and copied through pinned memory);
* the per-iteration roll (:647-661, :784-786) is *virtual*: ``roll()`` only records the offset and
  the kernels index the un-rolled image / feature maps circularly, so no data moves;
* tiles are evaluated by this rank's GPU in round-robin order (tile i -> rank i % world, as
  ``TileWorkerPool.request`` :284-298 does with worker processes) and the gradient tiles are
  exchanged with one all-gather instead of queues + shared memory.

There is no CPU fallback: constructing a TileEngine without CUDA or without the built library
raises.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib, sharding
from .netdesc import NetDesc

PRECISIONS = {'fp32': _lib.ST_PREC_FP32, 'bf16': _lib.ST_PREC_BF16, 'fp16': _lib.ST_PREC_FP16,
              'tc32': _lib.ST_PREC_TC32}


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class ContentData:
    """``ContentData`` message (style_transfer.py:165): features[layer] = f32[C,Hf,Wf]."""

    def __init__(self, features):
        self.features = features


class StyleData:
    """``StyleData`` message (style_transfer.py:166): grams[layer] = f32[C,C], lower triangle."""

    def __init__(self, grams):
        self.grams = grams


class TileEngine:
    def __init__(self, net, weights, mean=(0, 0, 0), device=0, precision='fp32', rank=0, world=1,
                 group=None):
        if not torch.cuda.is_available():
            raise _lib.StError('TileEngine needs a CUDA device (sm_100a); there is no CPU path')
        if not isinstance(net, NetDesc):
            raise TypeError('net must be a NetDesc')
        self.lib = _lib.load()
        self.net = net
        self.device = torch.device('cuda', device)
        self.precision = precision
        self.rank, self.world, self.group = rank, world, group
        self.mean = np.float32(mean).reshape((3, 1, 1))
        self.bgr = True
        self.shapes = net.shapes
        self.last_layer = net.last_layer
        self.contents, self.styles = [], []
        self.img = None
        self.roll_px = np.zeros(2, dtype=np.int64)          # accumulated (x, y) pixel roll
        ctx = C.c_void_p()
        _lib.call('st_create', device, PRECISIONS[precision], len(net.layers), net.to_ctypes(),
                  C.byref(ctx))
        self.ctx = ctx
        for i, layer in net.conv_layers():
            w, b = weights[layer.name]
            w = np.ascontiguousarray(w, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            if w.shape != (layer.cout, layer.cin, 3, 3) or b.shape != (layer.cout,):
                raise ValueError('weights of %s have the wrong shape' % layer.name)
            _lib.call('st_set_conv_params', ctx, i, w.ctypes.data_as(C.c_void_p),
                      b.ctypes.data_as(C.c_void_p))
        self._loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._packed = None
        self._packed_all = None
        self._comm_ready = False
        self._host_img, self._copy_stream, self._slabs = None, None, None

    def __del__(self):
        ctx, self.ctx = getattr(self, 'ctx', None), None
        if ctx:
            try:
                self.lib.st_destroy(ctx)
            except Exception:       # interpreter shutdown
                pass

    # ---- helpers ------------------------------------------------------------------------------
    def to_device(self, arr):
        """numpy / torch (any device) -> contiguous float32 CUDA tensor."""
        if isinstance(arr, torch.Tensor):
            return arr.to(self.device, torch.float32).contiguous()
        host = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        return host.pin_memory().to(self.device, non_blocking=True)

    def layers(self):
        return self.net.layer_names()

    def layer_info(self, layer):
        return self.net.layer_info(layer)

    def pil_to_image(self, img):
        """RGB HxWx3 array / PIL image -> mean-subtracted BGR f32[3,H,W] (:388-393)."""
        arr = np.float32(img).transpose((2, 0, 1))
        if self.bgr:
            arr = arr[::-1]
        return np.ascontiguousarray(arr - self.mean)

    def set_image(self, img):
        self.img = self.to_device(self.pil_to_image(img))

    def get_image_u8(self, params=None):
        """uint8 RGB [H][W][3] CUDA tensor of the current (or given) parameters (:378-386):
        mean added, BGR -> RGB, clipped and truncated on the device (st_get_image_u8)."""
        params = self.img if params is None else params
        params = params.contiguous()
        h, w = params.shape[-2:]
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=params.device)
        mean = (C.c_float * 3)(*[float(m) for m in np.ravel(self.mean)])
        _lib.call('st_get_image_u8', C.c_void_p(params.data_ptr()), h, w, mean, 1 if self.bgr else 0,
                  C.c_void_p(out.data_ptr()),
                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out

    def get_image_array(self, params=None):
        """The same picture on the host (numpy uint8 HxWx3): a quarter of the bytes of the f32
        parameters cross PCIe."""
        return self.get_image_u8(params).cpu().numpy()

    def _specs(self, layers, content_layers, style_layers, dd_layers, layer_weights,
               content_weight, style_weight, dd_weight):
        arr = (_lib.LossSpec * len(layers))()
        for i, layer in enumerate(layers):
            lw = layer_weights[layer]
            c, s, d = layer in content_layers, layer in style_layers, layer in dd_layers
            arr[i] = _lib.LossSpec(
                self.net.blob_index(layer), int(c), int(s), int(d),
                lw * content_weight[layer] if c else 0.0,
                lw * style_weight[layer] if s else 0.0,
                lw * dd_weight[layer] if d else 0.0)
        return arr

    def ordered_layers(self, *layer_sets):
        """Deepest-first list of the requested layers (TileWorker, :231-233)."""
        wanted = set().union(*[set(s) for s in layer_sets])
        return [l for l in reversed(self.layers()) if l in wanted]

    # ---- feature extraction ------------------------------------------------------------------
    def eval_features_tile(self, img, layers):
        """Computes a single tile in a set of feature maps (:421-427)."""
        img = self.to_device(img)
        h, w = img.shape[-2:]
        ids = (C.c_int32 * len(layers))(*[self.net.blob_index(l) for l in layers])
        outs = {}
        for layer in layers:
            hf, wf = self.net.feature_hw(layer, h, w)
            outs[layer] = torch.empty((self.shapes[layer][0], hf, wf), dtype=torch.float32,
                                      device=self.device)
        ptrs = (C.c_void_p * len(layers))(*[outs[l].data_ptr() for l in layers])
        _lib.call('st_eval_features_tile', self.ctx, _ptr(img), h, w, len(layers), ids, ptrs,
                  _stream())
        return outs

    def tile_boxes(self, img_size, tile_size):
        """[(start(y,x), end(y,x))] row-major (:431-450, :619-631)."""
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

    def rolled_image(self):
        """The image as the reference would hold it after its physical rolls."""
        if not self.roll_px.any():
            return self.img
        return torch.roll(self.img, (int(self.roll_px[0]), int(self.roll_px[1])), dims=(-1, -2))

    def eval_features_once(self, layers, tile_size=512):
        """Computes the set of feature maps for the (rolled) image, tile by tile (:429-464)."""
        img = self.rolled_image()
        img_size = np.array(img.shape[-2:])
        features = {}
        for layer in layers:
            scale, channels = self.layer_info(layer)
            shape = (channels,) + tuple(int(v) for v in np.ceil(img_size / scale))
            features[layer] = torch.zeros(shape, dtype=torch.float32, device=self.device)
        for start, end in self.tile_boxes(img_size, tile_size):
            tile = img[:, start[0]:end[0], start[1]:end[1]].contiguous()
            feats = self.eval_features_tile(tile, layers)
            for layer, feat in feats.items():
                scale, _ = self.layer_info(layer)
                s = start // scale
                e = s + np.array(feat.shape[-2:])
                features[layer][:, s[0]:e[0], s[1]:e[1]] = feat
        return features

    def prepare_features(self, layers, tile_size=512, passes=10):
        """Averages feature maps over randomly rolled passes to obscure tiling (:466-486).
        Draws from the global numpy RNG exactly like the reference."""
        img_size = np.array(self.img.shape[-2:])
        if max(*img_size) <= tile_size:
            passes = 1
        features = {}
        for i in range(passes):
            xy = np.array((0, 0))
            if i > 0:
                xy = np.int32(np.random.uniform(size=2) * img_size) // 32
            self.roll(xy)
            self.roll_features(features, xy)
            feats = self.eval_features_once(layers, tile_size)
            for layer in layers:
                if i == 0:
                    features[layer] = feats[layer] / passes
                else:
                    features[layer] += feats[layer] * np.float32(1 / passes)
            self.roll(-xy)
            self.roll_features(features, -xy)
        return features

    def gram_matrix(self, feat):
        """``num_utils.gram_matrix`` on the device: lower-triangular f32[C,C]."""
        feat = self.to_device(feat)
        c = feat.shape[0]
        hw = feat[0].numel()
        out = torch.empty((c, c), dtype=torch.float32, device=self.device)
        _lib.call('st_gram', self.ctx, _ptr(feat), c, hw, _ptr(out), _stream())
        return out

    def preprocess_images(self, content_images, style_images, content_layers, style_layers,
                          tile_size=512, content_passes=10):
        """Style Grams and content features (:488-554; arrays in ``pil_to_image`` format).  An entry
        of ``style_images`` may be a LIST of arrays -- the scaled copies of one style image under
        ``--style-multiscale`` (:501-524, built by cli.style_multiscale_variants): every copy adds its
        Gram matrices and counts once in the average.  ``content_passes`` = 1 is the ``roll is not
        None`` case of the reference (:545-549): the per-iteration preprocessing of ``--jitter``."""
        saved_img, saved_roll = self.img, self.roll_px.copy()
        self.roll_px[:] = 0
        if not self.styles:
            grams, count = {}, 0
            for entry in style_images:
                for image in (entry if isinstance(entry, (list, tuple)) else [entry]):
                    self.img = self.to_device(image)
                    feats = self.prepare_features(style_layers, tile_size, passes=1)
                    for layer in feats:
                        gram = self.gram_matrix(feats[layer])
                        grams[layer] = gram if layer not in grams else grams[layer] + gram
                    count += 1
            for gram in grams.values():
                gram /= count
            self.styles.append(StyleData(grams))
        for image in content_images:
            self.img = self.to_device(image)
            feats = self.prepare_features(content_layers, tile_size, passes=content_passes)
            self.contents.append(ContentData(feats))
        self.img, self.roll_px = saved_img, saved_roll

    def set_contents_and_styles(self, contents=None, styles=None, contents_only=False):
        """``TileWorkerPool.set_contents_and_styles`` (:309-332): hands the targets to the
        worker -- here, copies them into the library context of this rank's GPU.
        ``contents_only``: the per-iteration update of ``--jitter`` (:788-794) -- the content feature
        maps are overwritten in place (same shapes: no free / allocate, no device synchronisation),
        the style Grams stay as they are."""
        contents = self.contents if contents is None else contents
        if contents_only:
            for i, content in enumerate(contents):
                for layer, feat in content.features.items():
                    feat = self.to_device(feat)
                    _lib.call('st_set_content_features', self.ctx, i, self.net.blob_index(layer),
                              _ptr(feat), feat.shape[1], feat.shape[2], _stream())
            return
        styles = self.styles if styles is None else styles
        _lib.call('st_clear_targets', self.ctx)
        for i, content in enumerate(contents):
            for layer, feat in content.features.items():
                feat = self.to_device(feat)
                _lib.call('st_set_content_features', self.ctx, i, self.net.blob_index(layer),
                          _ptr(feat), feat.shape[1], feat.shape[2], _stream())
        for i, style in enumerate(styles):
            for layer, gram in style.grams.items():
                gram = self.to_device(gram)
                _lib.call('st_set_style_gram', self.ctx, i, self.net.blob_index(layer), _ptr(gram),
                          _stream())
        torch.cuda.current_stream().synchronize()

    # ---- loss + gradient ----------------------------------------------------------------------
    def eval_sc_grad_tile(self, img, start, layers, content_layers, style_layers, dd_layers,
                          layer_weights, content_weight, style_weight, dd_weight, roll=(0, 0)):
        """Evaluates an individual style+content gradient tile (:556-612).  ``roll`` = (x, y) is
        the pixel roll the worker would have applied to its content features (req.roll, :234).
        Returns (loss, grad) as a Python float and a CUDA tensor."""
        img = self.to_device(img)
        h, w = img.shape[-2:]
        specs = self._specs(layers, content_layers, style_layers, dd_layers, layer_weights,
                            content_weight, style_weight, dd_weight)
        grad = torch.empty_like(img)
        self._loss.zero_()
        _lib.call('st_eval_sc_grad_tile', self.ctx, _ptr(img), h, w, int(start[0]), int(start[1]),
                  int(roll[1]), int(roll[0]), len(layers), specs, _ptr(self._loss), _ptr(grad),
                  h * w, w, _stream())
        return float(self._loss.item()), grad

    def eval_sc_grad(self, roll, content_layers, style_layers, dd_layers, layer_weights,
                     content_weight, style_weight, dd_weight, tile_size, img=None, regularizers=None):
        """Evaluates the summed style and content gradients (:614-645) of ``img`` (default
        ``self.img``) under the virtual roll ``roll`` = (x, y) pixels.  Returns (loss, grad) as
        CUDA tensors (float64[1], float32[3,H,W]) in the UN-rolled frame; nothing is synchronised."""
        img = self.img if img is None else img
        H, W = img.shape[-2:]
        rx, ry = int(roll[0]), int(roll[1])
        layers = self.ordered_layers(content_layers, style_layers, dd_layers)
        specs = self._specs(layers, content_layers, style_layers, dd_layers, layer_weights,
                            content_weight, style_weight, dd_weight)
        nfl = sharding.packed_floats(H, W, tile_size, self.world)
        if self._packed is None or self._packed.numel() != nfl:
            self._packed = torch.zeros(nfl, dtype=torch.float32, device=self.device)
            self._packed_all = (torch.empty((self.world, nfl), dtype=torch.float32, device=self.device)
                                if self.world > 1 else None)
        # this rank's loss accumulates in the tail of its chunk: gradients and losses travel in ONE
        # all-gather
        self._packed[-4:].zero_()
        tail = C.c_void_p(self._packed.data_ptr() + (nfl - 4) * 4)
        staged, self._host_img = self._host_img, None
        if staged is not None and staged.shape == img.shape:
            self._eval_tiles_from_host(staged, img, ry, rx, tile_size, layers, specs, tail)
        else:
            _lib.call('st_eval_sc_grad_tiles', self.ctx, _ptr(img), H, W, ry, rx, tile_size,
                      self.rank, self.world, len(layers), specs, tail, _ptr(self._packed), _stream())
        if self.world == 1:
            packed_all = self._packed
        elif self._comm_ready:
            packed_all = self._packed_all
            _lib.call('st_allgather_grad', self.ctx, _ptr(self._packed), _ptr(packed_all), nfl,
                      _stream())
        else:
            packed_all = sharding.exchange(self._packed, self.world, self.group)
        loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        grad = torch.empty_like(img)
        if regularizers is None:
            _lib.call('st_unpack_grad', _ptr(packed_all), H, W, ry, rx, tile_size, self.world,
                      _ptr(grad), _ptr(loss), _stream())
        else:
            # StyleTransfer.eval_loss_and_grad's full-image terms (:700-736) in the same pass
            mean, tv_w, tv_beta, p_w, p_pow, aux, aux_w = regularizers
            _lib.call('st_unpack_regularize', _ptr(packed_all), _ptr(img), H, W, ry, rx, tile_size,
                      self.world, mean, tv_w, tv_beta, p_w, p_pow,
                      C.c_void_p(aux.data_ptr()) if aux is not None else None, aux_w, _ptr(loss),
                      _ptr(grad), _stream())
        return loss, grad

    # ---- image arriving from the host ----------------------------------------------------------
    def stage_host_image(self, host_img):
        """The next ``eval_sc_grad`` takes its image from ``host_img`` (pinned f32 [3,H,W], the
        layout of ``self.img``) instead of the resident copy -- the situation of a caller that owns
        the parameters on the host, as the reference's master process does.  One GPU: the rows the
        first tile row reads are uploaded first and the rest streams in while those tiles are being
        evaluated.  Several GPUs: every rank uploads only its 1/world slab of rows
        through its own PCIe link and the slabs are all-gathered over NVLink."""
        self._host_img = host_img

    def _copy_rows(self, host_img, img, start, count):
        """img[:, r] = host_img[:, r] for the circular row range [start, start + count) (mod H),
        plane by plane (contiguous pinned -> device copies on the current stream)."""
        H = img.shape[-2]
        for r0, r1 in ((start, min(start + count, H)), (0, max(start + count - H, 0))):
            if r1 > r0:
                for c in range(img.shape[0]):
                    img[c, r0:r1].copy_(host_img[c, r0:r1], non_blocking=True)

    def _eval_tiles_from_host(self, host_img, img, ry, rx, tile_size, layers, specs, tail):
        H, W = img.shape[-2:]
        cur = torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        cs.wait_stream(cur)                   # earlier readers of the resident image are done
        if self.world > 1 and H % self.world == 0:
            import torch.distributed as dist
            rows = H // self.world
            if self._slabs is None or self._slabs.shape != (self.world, img.shape[0], rows, W):
                self._slabs = torch.empty((self.world, img.shape[0], rows, W), dtype=torch.float32,
                                          device=self.device)
            with torch.cuda.stream(cs):
                self._copy_rows(host_img[:, self.rank * rows:(self.rank + 1) * rows],
                                self._slabs[self.rank], 0, rows)
            cur.wait_stream(cs)
            dist.all_gather_into_tensor(self._slabs.view(-1), self._slabs[self.rank].view(-1),
                                        group=self.group)
            img.view(img.shape[0], self.world, rows, W).copy_(self._slabs.permute(1, 0, 2, 3))
            split = [(0, -1, None)]
        else:
            nty, ntx, th, _, _, _ = sharding.tile_grid(H, W, tile_size)
            n_local = len(range(self.rank, nty * ntx, self.world))
            # the first phase is ONE tile row when there are three or more (its upload is the exposed
            # part; the rest streams in behind the evaluation of that row), else half of the rows
            rows_a = (1 if nty >= 3 else nty // 2) if self.world == 1 else 0
            with torch.cuda.stream(cs):
                if rows_a > 0:
                    # rolled rows [0, rows_a * th) = un-rolled rows starting at (-ry) mod H
                    self._copy_rows(host_img, img, (-ry) % H, rows_a * th)
                    ev_a = torch.cuda.Event()
                    ev_a.record(cs)
                    self._copy_rows(host_img, img, (rows_a * th - ry) % H, H - rows_a * th)
                else:
                    self._copy_rows(host_img, img, 0, H)
                ev_b = torch.cuda.Event()
                ev_b.record(cs)
            split = ([(0, rows_a * ntx, ev_a), (rows_a * ntx, n_local - rows_a * ntx, ev_b)]
                     if rows_a > 0 else [(0, -1, ev_b)])
        for first, count, ev in split:
            if ev is not None:
                cur.wait_event(ev)
            _lib.call('st_eval_sc_grad_tile_range', self.ctx, _ptr(img), H, W, ry, rx, tile_size,
                      self.rank, self.world, first, count, len(layers), specs, tail,
                      _ptr(self._packed), _stream())

    def init_comm(self):
        """Creates the context's NCCL communicator (``st_comm_init``) so that the exchange step runs
        from the C ABI on the caller's stream.  The 128-byte id is made by rank 0 and broadcast
        through the ``torch.distributed`` group the ranks already share."""
        if self.world == 1 or self._comm_ready:
            return
        import torch.distributed as dist
        uid = torch.zeros(_lib.ST_COMM_ID_BYTES, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * _lib.ST_COMM_ID_BYTES)()
            _lib.call('st_comm_unique_id', buf)
            uid = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        uid = uid.to(self.device)
        dist.broadcast(uid, src=0, group=self.group)
        host = bytes(uid.cpu().numpy().tobytes())
        ok = torch.ones(1, dtype=torch.int32, device=self.device)
        try:
            _lib.call('st_comm_init', self.ctx, host, self.rank, self.world)
        except _lib.StError as e:                 # e.g. no libnccl.so.2 in this process
            ok.zero_()
            err = e
        # all ranks take the same path: the C-ABI collective only if every rank has a communicator
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        self._comm_ready = bool(ok.item())
        if not self._comm_ready and self.rank == 0:
            print('note: st_comm_init unavailable on some rank (%s); the exchange step runs through '
                  'torch.distributed instead' % (err if 'err' in locals() else 'another rank'))

    # ---- roll -----------------------------------------------------------------------------------
    def roll_features(self, feats, xy, jitter_scale=32):
        """Rolls an individual set of feature maps in place (:647-655)."""
        xy = np.asarray(xy) * jitter_scale
        for layer, feat in feats.items():
            scale, _ = self.layer_info(layer)
            sh = xy // scale
            if sh.any():
                feat.copy_(torch.roll(feat, (int(sh[0]), int(sh[1])), dims=(-1, -2)))
        return feats

    def roll(self, xy, jitter_scale=32):
        """Rolls the image (:657-661) -- virtually: only the offset is recorded.  (The reference
        also rolls the master's copy of the content features here, which no worker ever sees.)"""
        # not reduced modulo the image size: feature rolls are floor(roll / scale), which differs
        # between congruent rolls when the image size is not a multiple of the layer scale
        self.roll_px += np.asarray(xy, dtype=np.int64) * jitter_scale
