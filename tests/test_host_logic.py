"""CPU tests of the host side: the C-ABI library loads and exports what include/style_b200.h
declares, the tile geometry / sharding logic matches the reference's arithmetic, the network
descriptions match the reference's deploy files, and a world-size-2 gloo run of the dispatcher
(CPU oracle as a stand-in for the per-tile operator) reproduces the single-process result."""

import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI ------------------------------------------------------------------------------------------
def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'style_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return re.findall(r'ST_API\s+[\w\s\*]+?\b(st_\w+)\s*\(', text)


def test_header_symbols_are_exported_and_bound():
    from style_transfer_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25 and len(set(names)) == len(names)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), 'libstyle_b200.so does not export %s' % name
        assert name in _lib.PROTOTYPES, 'no ctypes prototype for %s' % name
    assert set(_lib.PROTOTYPES) == set(names)
    assert _lib.load().st_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the context cannot be created and the engine refuses to construct."""
    import torch
    from style_transfer_b200 import _lib, netdesc, weights
    from style_transfer_b200.engine import TileEngine
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    net = netdesc.from_model('vgg16.prototxt')
    with pytest.raises(_lib.StError):
        TileEngine(net, weights.he_normal(net))
    ctx = C.c_void_p()
    rc = _lib.load().st_create(0, 0, len(net.layers), net.to_ctypes(), C.byref(ctx))
    assert rc < 0 and _lib.load().st_last_error()


@pytest.mark.parametrize('H,W,tile', [(2048, 2048, 512), (1448, 1448, 512), (724, 724, 512),
                                      (75, 52, 40), (64, 96, 32), (256, 256, 512), (513, 512, 512)])
def test_tile_grid_matches_reference_arithmetic(H, W, tile):
    from oracle.tile_operator import tile_grid as oracle_grid
    from style_transfer_b200 import _lib, sharding
    boxes = sharding.tile_boxes(H, W, tile)
    want = [(int(s[0]), int(s[1]), int(e[0]), int(e[1])) for s, e in oracle_grid((H, W), tile)]
    assert boxes == want
    nty, ntx, thm, twm = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _lib.call('st_tile_grid', H, W, tile, C.byref(nty), C.byref(ntx), C.byref(thm), C.byref(twm))
    g = sharding.tile_grid(H, W, tile)
    assert (nty.value, ntx.value, thm.value, twm.value) == (g[0], g[1], g[4], g[5])
    assert thm.value == max(b[2] - b[0] for b in boxes) and twm.value == max(b[3] - b[1] for b in boxes)
    # round-robin placement: every tile exactly once, slots consecutive per rank
    for world in (1, 2, 3, 4, 8):
        seen = []
        for rank in range(world):
            local = sharding.local_tiles(H, W, tile, rank, world)
            assert [s for s, _ in local] == list(range(len(local)))
            seen += [b for _, b in local]
        assert sorted(seen) == sorted(boxes)
        assert sharding.packed_shape(H, W, tile, world)[0] == -(-len(boxes) // world)
        nfl = sharding.packed_floats(H, W, tile, world)
        assert nfl == _lib.load().st_packed_floats(H, W, tile, world)
        assert nfl % 4 == 0 and nfl >= int(np.prod(sharding.packed_shape(H, W, tile, world))) + 4


def test_scale_ladder_and_weights():
    from style_transfer_b200.transfer import default_args, parse_weights, scale_ladder
    assert scale_ladder(2048, 256) == [256, 362, 512, 724, 1024, 1448, 2048]
    assert scale_ladder(256, 182) == [256]
    names, w = parse_weights(['conv4_2:2', 'conv5_1'], 0.05)
    assert names == ['conv4_2', 'conv5_1']
    assert w['conv4_2'] == pytest.approx(0.05 * 2 / 3) and w['conv5_1'] == pytest.approx(0.05 / 3)
    a = default_args()
    assert (a.size, a.tile_size, a.optimizer, a.step_size, a.avg_window) == (256, 512, 'adam', 15.0, 20.0)
    assert a.style_layers == ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']


# ---- network descriptions -----------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['vgg16', 'vgg19', 'vgg16_avgpool', 'vgg19_avgpool', 'vgg16_big',
                                  'vgg19_big'])
def test_netdesc_roundtrip_and_reference_prototxt(name):
    from style_transfer_b200 import netdesc
    net = netdesc.from_model(name + '.prototxt')
    again = netdesc.parse_prototxt(netdesc.to_prototxt(net), name)
    assert [(l.kind, l.name, l.bottom, l.cin, l.cout, l.pool) for l in again.layers] == \
        [(l.kind, l.name, l.bottom, l.cin, l.cout, l.pool) for l in net.layers]
    assert net.layer_info('conv4_2') == (8, 512) if 'big' not in name else True
    ref = os.path.join('/root/reference', name + '.prototxt')
    if os.path.exists(ref):                       # only in the build container
        parsed = netdesc.parse_prototxt(open(ref).read(), name)
        assert [(l.kind, l.name, l.bottom, l.cin, l.cout, l.pool) for l in parsed.layers] == \
            [(l.kind, l.name, l.bottom, l.cin, l.cout, l.pool) for l in net.layers]
        assert parsed.shapes == net.shapes


def test_reference_shape_tables():
    """The shape tables the reference hard-codes (style_transfer.py:1030-1073)."""
    from style_transfer_b200 import netdesc
    net = netdesc.from_model('vgg19.prototxt')
    assert net.shapes['conv1_1'] == (64, 224, 224) and net.shapes['pool5'] == (512, 7, 7)
    assert net.shapes['conv3_4'] == (256, 56, 56) and net.shapes['conv5_1'] == (512, 14, 14)
    assert [net.layer_info(l)[0] for l in ('conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1')] == \
        [1, 2, 4, 8, 16]
    assert 'conv3_4' not in netdesc.from_model('vgg16.prototxt').shapes


# ---- world-size-2 dispatcher run over gloo --------------------------------------------------------------
def _dispatcher_worker(rank, world, port, H, W, tile, roll, out_path):
    """One rank: evaluates its round-robin tiles with the CPU oracle (standing in for the CUDA
    per-tile operator), exchanges them exactly as TileEngine.eval_sc_grad does."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import numeric as on
    from oracle.caffe_net import he_normal_weights, model_layers
    from oracle.tile_operator import OracleModel
    from oracle.transfer import to_params
    from style_transfer_b200 import sharding
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    model = 'vgg16.prototxt'
    ora = OracleModel(model, he_normal_weights(model_layers(model)))
    rs = np.random.RandomState(0)
    img, content, style = (to_params(rs.randint(0, 256, (H, W, 3))) for _ in range(3))
    c_layers, s_layers = ['conv2_2'], ['conv1_1', 'conv2_1']
    ora.img = style
    ora.styles = [{l: on.gram_lower(f) for l, f in ora.features_once(s_layers, tile).items()}]
    ora.img = content
    ora.contents = [ora.features_once(c_layers, tile)]
    ora.publish()
    lw = {l: 1.0 for l in ora.layers()}
    cw, sw = {'conv2_2': 0.05}, {'conv1_1': 0.5, 'conv2_1': 0.5}
    layers = ora.ordered_layers(c_layers, s_layers)
    rolled = on.roll2_(img.copy(), np.array(roll))
    ora.roll_features_all(ora.w_contents, np.array(roll), 1)
    # this rank's chunk of the exchange buffer: tiles + the loss in its tail, ONE all-gather
    chunk = torch.zeros(sharding.packed_floats(H, W, tile, world), dtype=torch.float32)
    tiles = sharding.tiles_view(chunk, H, W, tile, world)
    for slot, (sy, sx, ey, ex) in sharding.local_tiles(H, W, tile, rank, world):
        l, g = ora.sc_grad_tile(np.ascontiguousarray(rolled[:, sy:ey, sx:ex]), np.array([sy, sx]),
                                layers, c_layers, s_layers, [], lw, cw, sw, {})
        tiles[slot, :, :ey - sy, :ex - sx] = torch.from_numpy(g)
        sharding.loss_view(chunk)[0] += float(l)
    packed_all = sharding.exchange(chunk, world)
    grad, loss = sharding.unpack_numpy(packed_all.numpy(), H, W, tile, roll[1], roll[0])
    if rank == 0:
        np.savez(out_path, grad=grad, loss=np.float64([loss]))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_dispatch_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import numeric as on
    from oracle.caffe_net import he_normal_weights, model_layers
    from oracle.tile_operator import OracleModel
    from oracle.transfer import to_params
    H, W, tile, roll = 40, 56, 24, (8, -16)
    out = str(tmp_path / 'two_rank.npz')
    port = 29500 + os.getpid() % 400
    mp.spawn(_dispatcher_worker, args=(2, port, H, W, tile, roll, out), nprocs=2, join=True)
    got = np.load(out)
    # single-process oracle of the same evaluation (eval_sc_grad, style_transfer.py:614-645)
    model = 'vgg16.prototxt'
    ora = OracleModel(model, he_normal_weights(model_layers(model)))
    rs = np.random.RandomState(0)
    img, content, style = (to_params(rs.randint(0, 256, (H, W, 3))) for _ in range(3))
    c_layers, s_layers = ['conv2_2'], ['conv1_1', 'conv2_1']
    ora.img = style
    ora.styles = [{l: on.gram_lower(f) for l, f in ora.features_once(s_layers, tile).items()}]
    ora.img = content
    ora.contents = [ora.features_once(c_layers, tile)]
    ora.publish()
    lw = {l: 1.0 for l in ora.layers()}
    ora.img = on.roll2_(img.copy(), np.array(roll))
    loss, grad = ora.sc_grad(np.array(roll), c_layers, s_layers, [], lw, {'conv2_2': 0.05},
                             {'conv1_1': 0.5, 'conv2_1': 0.5}, {}, tile)
    grad = on.roll2_(grad.copy(), -np.array(roll))
    assert np.array_equal(got['grad'], grad)
    assert abs(float(got['loss'][0]) - loss) <= 1e-9 * abs(loss)


# ---- flag surface ----------------------------------------------------------------------------------------
def test_flag_table_matches_reference_defaults(tmp_path, monkeypatch):
    from style_transfer_b200 import config_system
    monkeypatch.setattr(config_system, 'CONFIG_PY', tmp_path / 'config.py')
    a = config_system.parse_args(['-ci', 'c.png', '-si', 's1.png', 's2.png'])
    assert a.display == 'browser'                     # the reference's default (config_system.py:56)
    assert (a.size, a.min_size, a.tile_size, a.optimizer) == (256, 182, 512, 'adam')
    assert (a.step_size, a.step_decay, a.avg_window) == (15, [0.05, 0.5], 20)
    assert (a.content_weight, a.dd_weight, a.tv_weight, a.tv_power) == (0.05, 0, 5, 2)
    assert (a.p_weight, a.p_power, a.aux_weight) == (2, 6, 10)
    assert a.iterations == [200, 100] and a.devices == [-1] and a.seed == 0 and a.div == 1
    assert tuple(a.mean) == (103.939, 116.779, 123.68)
    assert a.style_images == ['s1.png', 's2.png'] and a.model == 'vgg19.prototxt'
    # fractions (ffloat), short forms, layer lists with weights
    b = config_system.parse_args(['-ci', 'c', '-si', 's', '-s', '2048', '-cw', '1/20', '-o', 'lbfgs',
                                  '--devices', '0', '1', '2', '3', '--content-layers', 'conv4_2:2',
                                  'conv5_2', '-i', '300'])
    assert b.size == 2048 and b.content_weight == 0.05 and b.optimizer == 'lbfgs'
    assert b.devices == [0, 1, 2, 3] and b.content_layers == ['conv4_2:2', 'conv5_2']
    assert b.iterations == [300]
    if os.path.exists('/root/reference/config_system.py'):
        # every flag of the reference exists here with the same default
        src = open('/root/reference/config_system.py').read()
        for flag in re.findall(r"arg\('(--[\w-]+)'", src):
            name = flag[2:].replace('-', '_')
            assert hasattr(a, name), flag


def test_config_precedence(tmp_path, monkeypatch):
    """defaults < config.py beside the entry script < argv (values that differ from the default)
    < --config FILE (config_system.py:121-136)."""
    from style_transfer_b200 import config_system
    # the default location is beside style_transfer.py, as the reference's is (config_system.py:14)
    assert config_system.CONFIG_PY == __import__('pathlib').Path(ROOT) / 'config.py'
    monkeypatch.setattr(config_system, 'CONFIG_PY', tmp_path / 'config.py')
    (tmp_path / 'config.py').write_text('size = 512\ntile_size = 256\ntv_weight = 1\n')
    (tmp_path / 'extra.py').write_text('tv_weight = 7\n')
    a = config_system.parse_args(['-ci', 'c', '-si', 's', '--tile-size', '384'])
    assert (a.size, a.tile_size, a.tv_weight) == (512, 384, 1)
    b = config_system.parse_args(['-ci', 'c', '-si', 's', '--config', str(tmp_path / 'extra.py')])
    assert (b.size, b.tile_size, b.tv_weight) == (512, 256, 7)


def test_config_globals_and_callable_values(tmp_path, monkeypatch):
    """Config files see detect_devices / math / np (config_system.py:187) and may define callables
    that are evaluated against the STATE object on every read (AutocallNamespace, :151-184).  The
    reference's own docker/config.py (``devices = detect_devices()``) is the first case."""
    import argparse
    from style_transfer_b200 import config_system
    monkeypatch.setattr(config_system, 'CONFIG_PY', tmp_path / 'config.py')
    docker_config = ("caffe_path = '/root/caffe'\ndevices = detect_devices()\n"
                     "display = 'none'\ndiv = 8\n")                     # docker/config.py verbatim
    ref = '/root/reference/docker/config.py'
    if os.path.exists(ref):
        assert open(ref).read() == docker_config
    (tmp_path / 'config.py').write_text(docker_config)
    a = config_system.parse_args(['-ci', 'c', '-si', 's'])
    assert a.div == 8 and a.display == 'none' and a.caffe_path == '/root/caffe'
    assert isinstance(a.devices, list) and all(isinstance(d, int) for d in a.devices)
    (tmp_path / 'dyn.py').write_text(
        'size = int(np.sqrt(512 * 512))\n'
        'tv_weight = lambda st: 5 if st.scale < 2 else 2 * math.pi\n')
    state = argparse.Namespace()
    b = config_system.parse_args(['-ci', 'c', '-si', 's', '--config', str(tmp_path / 'dyn.py')], state)
    assert b.size == 512
    assert isinstance(b.tv_weight, config_system.ValuePlaceholder)     # STATE.scale does not exist yet
    state.scale = 0
    assert b.tv_weight == 5
    state.scale = 3
    assert abs(b.tv_weight - 2 * np.pi) < 1e-12
    assert 'tv_weight' in b and 'nonsense' not in b
    assert getattr(b, 'jitter', None) is False and getattr(b, 'not_a_flag', 7) == 7


def test_cli_style_multiscale_scales_the_original_style_image(monkeypatch):
    """--style-multiscale hands the ORIGINAL style image to the variant loop (style_transfer.py:
    857-858 + :501-524), not the copy resized for the current scale: with -sm 64 256 on a 300x200
    style image at a 96-px first scale the Grams average the 64..256-px copies of the original."""
    from types import SimpleNamespace
    from PIL import Image
    from style_transfer_b200 import cli
    rs = np.random.RandomState(3)
    style = Image.fromarray(rs.randint(0, 256, (200, 300, 3)).astype(np.uint8))
    content = Image.fromarray(rs.randint(0, 256, (80, 120, 3)).astype(np.uint8))
    seen = []

    class FakeModel:
        rank, styles = 0, []

        def pil_to_image(self, im):
            return im.size                                # (w, h) stands for the array

    class FakeTransfer:
        model = FakeModel()
        current_raw, optimizer = None, None

        def init_first_scale(self, h, w, init):
            pass

        def transfer(self, iters, contents, styles, callback=None):
            seen.append(styles)
            return 'out'

    args = SimpleNamespace(size=96, min_size=96, div=1, style_scale=1.0, max_style_size=None,
                           style_scale_up=False, style_multiscale=[64, 256], iterations=[1])
    cli.transfer_multiscale(FakeTransfer(), args, [content], [style])
    assert len(seen) == 1 and len(seen[0]) == 1
    sizes = [max(wh) for wh in seen[0][0]]
    assert sizes == [64, 91, 128, 181, 256], sizes        # reference loop: 256 down by sqrt(2) to >= 64
    # without the flag the style is resized for the scale (96 * style_scale)
    seen.clear()
    args.style_multiscale = None
    cli.transfer_multiscale(FakeTransfer(), args, [content], [style])
    assert seen[0] == [(96, 64)]


def test_oracle_output_step_restates_reference_lines():
    """oracle.transfer.iter_stats / get_image_array against the literal numpy expressions of
    style_transfer.py:808-815 and :378-386 on a hand-made case (incl. clipping and BGR reversal)."""
    from oracle.transfer import get_image_array, iter_stats
    avg = np.float32(np.arange(3 * 2 * 3).reshape(3, 2, 3)) * np.float32(20) - np.float32(150)
    old = avg + np.float32(1.5)
    us, tv = iter_stats(avg, old)
    assert abs(us - 1.5) < 1e-6 and np.array_equal(old, avg)
    xd = avg - np.roll(avg, -1, axis=-1)
    yd = avg - np.roll(avg, -1, axis=-2)
    assert abs(tv - float(np.sqrt(np.mean(xd ** 2 + yd ** 2)))) < 1e-6
    pic = get_image_array(avg)
    mean = np.float32((103.939, 116.779, 123.68))
    assert pic.shape == (2, 3, 3) and pic.dtype == np.uint8
    # picture channel 0 (R) is model plane 2 (the model is BGR); values clip to [0, 255] and truncate
    assert pic[0, 0, 0] == np.uint8(np.clip(avg[2, 0, 0] + mean[2], 0, 255))
    assert pic[1, 2, 2] == np.uint8(np.clip(avg[0, 1, 2] + mean[0], 0, 255))
    assert pic.min() == 0          # plane 0 starts at -150 + 103.9 < 0


@pytest.mark.parametrize('method,name', [(0, 'lanczos'), (1, 'bilinear')])
def test_resample_coefficient_table_matches_oracle(method, name):
    """st_resample_coeffs (host half of st_resize_f32; no device needed) against the oracle's
    restatement of Pillow's precompute_coeffs: same bounds, the float64 weights bit for bit -- up-
    and down-scaling, including the ragged supports at the image borders."""
    import ctypes as C
    from oracle import numeric as on
    from style_transfer_b200 import _lib
    for in_size, out_size in ((53, 75), (64, 91), (70, 49), (181, 256), (512, 362), (7, 3)):
        ksize = C.c_int()
        _lib.call('st_resample_coeffs', in_size, out_size, method, C.byref(ksize), None, None)
        bounds = np.zeros((out_size, 2), np.int32)
        kk = np.zeros((out_size, ksize.value), np.float64)
        _lib.call('st_resample_coeffs', in_size, out_size, method, C.byref(ksize),
                  bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p))
        b_o, k_o = on.resample_coeffs(in_size, out_size, name)
        assert k_o.shape == kk.shape and np.array_equal(bounds, b_o)
        assert np.array_equal(kk, k_o), (in_size, out_size, np.abs(kk - k_o).max())


def test_style_multiscale_variants_follow_reference_loop():
    """cli.style_multiscale_variants against the loop of preprocess_images (:501-524) worked by
    hand: sizes 512, 362, 256, 181, 128, 91, 64 visited smallest first; the 300x200 image stops
    shrinking at 362 (returned unscaled, processed, then the loop ends); a 40-pixel short side is
    kept, a 28-pixel one skipped."""
    from PIL import Image
    from style_transfer_b200.cli import style_multiscale_variants
    im = Image.new('RGB', (300, 200))
    got = [v.size for v in style_multiscale_variants(im, 64, 512)]
    assert got == [(64, 43), (91, 61), (128, 85), (181, 121), (256, 171), (300, 200)]
    # MIN below 32 is clamped to 32; copies whose short side is < 32 are skipped, larger ones kept
    im = Image.new('RGB', (400, 100))
    got = [v.size for v in style_multiscale_variants(im, 1, 256)]
    assert got == [(128, 32), (181, 45), (256, 64)]
    # div quantises both sides like resize_to_fit (:965-975)
    im = Image.new('RGB', (300, 200))
    assert [v.size for v in style_multiscale_variants(im, 128, 181, div=8)] == [(128, 80), (176, 112)]


def test_jitter_copy_formulation_equals_in_place_rolls():
    """--jitter (style_transfer.py:757-759, 778-797).  The reference rolls the iterate and the
    optimizer state in place, evaluates, and rolls back; the engine's host loop
    (StyleTransfer._step_jitter) evaluates a rolled COPY and rolls the gradient back.  Both
    formulations are run here on the CPU oracle (restated reference semantics vs the engine's
    formulation built from the same oracle pieces): identical averaged iterates, bit for bit."""
    from oracle.caffe_net import he_normal_weights, model_layers
    from oracle import numeric as on
    from oracle.optimizers import Adam
    from oracle.tile_operator import OracleModel
    from oracle.transfer import OracleTransfer, default_args, parse_weights, to_params
    model = 'vgg16.prototxt'
    params_w = he_normal_weights(model_layers(model))
    rs = np.random.RandomState(21)
    H, W = 24, 32
    content, style = (to_params(rs.randint(0, 256, (H, W, 3))) for _ in range(2))
    args = default_args(size=W, min_size=W, tile_size=512, content_layers=['conv2_1'],
                        style_layers=['conv1_1'], model=model)
    args.jitter = True

    # (A) the oracle's restatement of the reference loop
    ora = OracleModel(model, params_w)
    tr = OracleTransfer(ora, args)
    np.random.seed(0)
    tr.init_first_scale(H, W)
    avg_a = tr.run(3, [content], [style]).copy()

    # (B) the engine's formulation on the same oracle pieces
    ora = OracleModel(model, params_w)
    tr = OracleTransfer(ora, args)
    np.random.seed(0)
    tr.init_first_scale(H, W)
    p = ora.img
    c_layers, c_weight = parse_weights(args.content_layers, args.content_weight)
    s_layers, s_weight = parse_weights(args.style_layers, 1)
    ora.contents, ora.styles = [], []
    ora.preprocess([], [style], [], s_layers, args.tile_size)
    ora.img = p
    avg_b = None
    for _ in range(3):
        xy = np.int32(np.random.uniform(-0.5, 0.5, size=2) * np.array([H, W])) // 1
        ora.contents = []
        ora.preprocess([content], [], c_layers, [], args.tile_size, roll=xy)
        ora.img = p
        ora.publish()
        sc_args = (np.zeros(2, dtype=np.int64), c_layers, s_layers, [], tr.layer_weights, c_weight,
                   s_weight, {}, args.tile_size)

        def opfunc(x):
            loss, grad = tr.loss_and_grad(on.roll2_(x.copy(), xy), sc_args)
            return loss, on.roll2_(grad.copy(), -xy)
        avg_b, _ = tr.optimizer.update(opfunc)
    assert np.array_equal(avg_a, avg_b)


def test_split_operand_arithmetic_model():
    """The arithmetic ST_PREC_TC32 relies on, modelled in numpy: x = hi + lo with hi = fp16(x),
    lo = fp16(x - hi) carries ~22 significant bits, and a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
    (three fp16 x fp16 products, exact in fp32, summed in fp32) is within 2^-20 of the fp32 product --
    against 2^-10 for single fp16 operands.  The weights are pre-scaled by a power of two so that their
    lo parts stay out of the fp16 subnormals (tc_pack_split); the same here."""
    rs = np.random.RandomState(0)
    a = np.float32(rs.uniform(0, 300, 20000))                  # post-ReLU activations
    w = np.float32(rs.randn(20000) * 0.02)                     # He-normal-sized weights
    scale = np.float32(2.0 ** (13 - np.frexp(np.abs(w).max())[1]))
    ws = w * scale

    def split(x):
        hi = np.float16(x)
        return hi, np.float16(x - np.float32(hi))
    a_hi, a_lo = split(a)
    w_hi, w_lo = split(ws)
    f = np.float32
    three = (f(a_hi) * f(w_hi) + f(a_lo) * f(w_hi) + f(a_hi) * f(w_lo)) / scale
    exact = np.float64(a) * np.float64(w)
    single = np.float64(f(a_hi)) * np.float64(f(np.float16(ws))) / np.float64(scale)
    err3 = np.abs(three - exact).max() / np.abs(exact).max()
    err1 = np.abs(single - exact).max() / np.abs(exact).max()
    assert err3 < 2.0 ** -20, err3
    assert err1 > 2.0 ** -13 and err3 < err1 / 200
    # the split itself: 2^-21 relative where lo is a normal fp16 number (|x| >= 1 here), and never
    # worse than one subnormal step (2^-24) in absolute terms below that
    err = np.abs((f(a_hi) + f(a_lo)) - a)
    big = np.abs(a) >= 1.0
    assert (err[big] / np.abs(a)[big]).max() < 2.0 ** -21
    assert err[~big].max() <= 2.0 ** -24


@pytest.mark.parametrize('H,tile,ry', [(2048, 512, 0), (2048, 512, -424), (2048, 512, 808), (1448, 512, 96),
                                      (1024, 512, -312), (512, 512, 40)])
def test_host_image_upload_phases_cover_every_row_once(H, tile, ry):
    """TileEngine.stage_host_image on one GPU uploads the image in two phases -- first the rows the
    first tile row reads (a circular range in the un-rolled frame: rolled row y is image row
    (y - roll) mod H), then the rest.  Index logic only (CPU tensors): the two phases together
    write every row exactly once, and phase A is exactly what the first batch of tiles reads."""
    import torch
    from style_transfer_b200 import sharding
    from style_transfer_b200.engine import TileEngine
    eng = object.__new__(TileEngine)                       # no device needed for _copy_rows
    W = 8
    host = torch.arange(3 * H * W, dtype=torch.float32).reshape(3, H, W)
    nty, _, th, _, _, _ = sharding.tile_grid(H, H, tile)
    rows_a = 1 if nty >= 3 else nty // 2                   # engine._eval_tiles_from_host
    img = torch.full((3, H, W), -1.0)
    count = torch.zeros(H, dtype=torch.int64)

    def copy(start, n):
        before = img.clone()
        TileEngine._copy_rows(eng, host, img, start, n)
        changed = (img != before).any(dim=2).any(dim=0)
        count.add_(changed.to(torch.int64))
        return changed
    if rows_a > 0:
        a = copy((-ry) % H, rows_a * th)
        # rolled rows [0, rows_a * th) are image rows (y - ry) mod H
        want = torch.zeros(H, dtype=torch.bool)
        want[[(y - ry) % H for y in range(rows_a * th)]] = True
        assert torch.equal(a, want)
        copy((rows_a * th - ry) % H, H - rows_a * th)
    else:
        copy(0, H)
    assert torch.equal(img, host) and int(count.min()) == 1 and int(count.max()) == 1


@pytest.mark.parametrize('H,W,tile,world', [(75, 52, 40, 3), (2048, 2048, 512, 8), (724, 724, 512, 3),
                                            (96, 128, 32, 5)])
def test_exchange_chunks_unpack_for_any_world(H, W, tile, world):
    """The chunk layout of the exchange step (tiles padded to 4 floats + the loss tail) for world
    sizes that do not divide the tile count and ragged grids: every rank fills its slots, the
    rank-major concatenation (= the all-gather result) un-packs to the directly assembled gradient
    and the losses add up."""
    from style_transfer_b200 import sharding
    rs = np.random.RandomState(H + world)
    full = rs.randn(3, H, W).astype(np.float32)            # the gradient in the rolled frame
    roll_y, roll_x = 8, -16
    nfl = sharding.packed_floats(H, W, tile, world)
    chunks = np.zeros((world, nfl), np.float32)
    losses = []
    for rank in range(world):
        tiles = sharding.tiles_view(chunks[rank], H, W, tile, world)
        for slot, (sy, sx, ey, ex) in sharding.local_tiles(H, W, tile, rank, world):
            tiles[slot, :, :ey - sy, :ex - sx] = full[:, sy:ey, sx:ex]
        losses.append(float(rank) + 0.25)
        sharding.loss_view(chunks[rank])[0] = losses[-1]
    grad, loss = sharding.unpack_numpy(chunks, H, W, tile, roll_y, roll_x)
    assert np.array_equal(grad, np.roll(full, (-roll_y, -roll_x), axis=(1, 2)))
    assert loss == sum(losses)
