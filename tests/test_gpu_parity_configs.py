"""GPU parity at the shapes and precisions the benchmark is quoted on (BASELINE.json configs 1-5):
the CUDA engine through the C ABI against the CPU oracle on identical seeded inputs, at full size.

What the small cases of test_gpu_parity.py cannot reach and these do: the persistent multi-wave
schedule of ``conv_tc2_kernel<256,9,*>``, the 4-D TMA boxes with a batch of 4 / 16 tiles, resident
weights on conv1_2 / conv2_1, the 64 / 32 / 8 / 4-way Gram splits, the fused pooling epilogues at
full tile width, and the tile grid + virtual roll at 1024^2 / 2048^2.

Stated tolerances (measured on B200 in brackets, see profiles/r02_parity_configs.md):

  single evaluation (loss relative, gradient relative L2 / max-norm):
    fp32 mode  : loss 1e-4, gradient max|d| <= 5e-4 * max|ref|
    tc32 mode  : loss 1e-4, gradient relative L2 <= 2e-3 (split fp16 hi+lo operands on tcgen05,
                 fp32-class; the residual is arg-max / ReLU flips at near ties)
    fp16 mode  : loss 5e-3, gradient relative L2 <= 6e-2 (max-pool nets), 2e-2 (average-pool)
    bf16 mode  : loss 2e-2, gradient relative L2 <= 1.5e-1 (max-pool nets)
  after N iterations at cfg1's real size (grey levels, pixel range 0..255): see the tests.

The oracle costs ~3 s per 512x512 VGG-19 tile evaluation on the box's host cores; its results are
cached per case so that every precision of a case shares one oracle run.
"""

import json
import os

import numpy as np
import pytest
import torch

from oracle.caffe_net import he_normal_weights, model_layers
from oracle import numeric as on
from oracle.tile_operator import OracleModel
from oracle.transfer import OracleTransfer, default_args, parse_weights, to_params

pytestmark = pytest.mark.gpu

MEAN = (103.939, 116.779, 123.68)
STYLE5 = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']
_REPORT = os.environ.get('ST_PARITY_REPORT')


def report(case, **vals):
    """One JSON line per measurement (ST_PARITY_REPORT=path): how the stated tolerances were set."""
    print(case, vals)
    if _REPORT:
        with open(_REPORT, 'a') as f:
            f.write(json.dumps(dict(case=case, **vals)) + '\n')


def l2rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.linalg.norm((a - b).ravel().astype(np.float64)) /
                 max(np.linalg.norm(b.ravel().astype(np.float64)), 1e-30))


def maxrel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rand_img(rs, h, w):
    return to_params(rs.randint(0, 256, (h, w, 3)))


def make_engine(model, params, precision, **kw):
    from style_transfer_b200 import netdesc
    from style_transfer_b200.engine import PRECISIONS, TileEngine
    if precision not in PRECISIONS:
        pytest.skip('precision mode %s is not built into this library' % precision)
    return TileEngine(netdesc.from_model(model), params, precision=precision, **kw)


def hand_targets(eng, ora):
    from style_transfer_b200.engine import ContentData, StyleData
    eng.set_contents_and_styles([ContentData(c) for c in ora.contents],
                                [StyleData(g) for g in ora.styles])


# (loss relative, gradient relative L2, fraction of pixels with |d| > 1e-3 * max|ref|); None = not bounded
EVAL_BOUNDS = {
    # fp32-class modes: the bulk of the pixels agrees to round-off; what remains is a sparse set of
    # pixels behind a max-pool arg-max / ReLU decision that fp32 round-off flips between two correct
    # implementations (one flipped window re-routes that window's whole gradient), see DESIGN.md
    # measured: fp32 L2 1.2e-4 .. 1.0e-3, 0.0005 .. 0.85 % of the pixels beyond 1e-3; tc32 L2 3.0e-4 ..
    # 1.3e-3, 0 .. 0.41 % (its 12-step accumulation chains round less than the SIMT kernel's one
    # sequential FFMA chain over K = 9 Cin)
    ('fp32', 'oracle'): (1e-4, 3e-3, 2e-2), ('tc32', 'oracle'): (1e-4, 3e-3, 1e-2),
    # 16-bit operand modes with the engine's OWN targets (the operating condition: style Grams and
    # content features come from the same kernels, so the systematic part of the weight / activation
    # rounding cancels in G - G_style)
    # measured: fp16 L2 2.3e-2 .. 4.7e-2, bf16 1.1e-1 .. 1.4e-1
    ('fp16', 'engine'): (5e-3, 7e-2, None), ('bf16', 'engine'): (2e-2, 2.5e-1, None),
    # ... and against the ORACLE's fp32 targets on synthetic noise images (style statistics == image
    # statistics, G - G_style is the difference of two nearly equal matrices: the worst case)
    # measured: fp16 L2 3.1e-2 .. 9.0e-2 (bf16: 0.5 .. 0.9, not a usable mode in this corner)
    ('fp16', 'oracle'): (5e-3, 1.5e-1, None),
}


def check_eval(case, precision, model, loss_g, grad_g, loss_o, grad_o, targets='oracle'):
    e_loss = abs(float(loss_g) - loss_o) / abs(loss_o)
    e_l2, e_max = l2rel(grad_g, grad_o), maxrel(grad_g, grad_o)
    g = grad_g.detach().cpu().numpy() if isinstance(grad_g, torch.Tensor) else np.asarray(grad_g)
    err = np.abs(g - grad_o) / np.abs(grad_o).max()
    frac = float((err > 1e-3).mean())
    q = [float(v) for v in np.quantile(err, [0.5, 0.99, 0.999])]
    report(case, precision=precision, targets=targets, loss_rel=e_loss, grad_l2rel=e_l2,
           grad_maxrel=e_max, frac_gt_1e3=frac, q50=q[0], q99=q[1], q999=q[2])
    b_loss, b_l2, b_frac = EVAL_BOUNDS[(precision, targets)]
    assert e_loss <= b_loss, e_loss
    assert e_l2 <= b_l2, e_l2
    if b_frac is not None:
        assert frac <= b_frac, frac


# ---- one tile: cfg1 (256^2 VGG-16), cfg2 (512^2 VGG-19), cfg4 (1024^2 VGG-19 average-pool) ---------------
_TILE_CASES = {
    'cfg1': ('vgg16.prototxt', 256, ['conv4_2'], ['conv3_1']),
    'cfg2': ('vgg19.prototxt', 512, ['conv4_2'], STYLE5),
    'cfg4': ('vgg19_avgpool.prototxt', 1024, ['conv4_2'], STYLE5),
}
_tile_cache = {}


def tile_case(name):
    if name in _tile_cache:
        return _tile_cache[name]
    model, size, c_layers, s_layers = _TILE_CASES[name]
    params = he_normal_weights(model_layers(model))
    ora = OracleModel(model, params)
    rs = np.random.RandomState(101)
    content, style, img = (rand_img(rs, size, size) for _ in range(3))
    ora.contents, ora.styles = [], []
    ora.preprocess([content], [style], c_layers, s_layers, size)
    ora.publish()
    lw = {l: 1.0 for l in ora.layers()}
    _, cw = parse_weights(c_layers, 0.05)
    _, sw = parse_weights(s_layers, 1)
    layers = ora.ordered_layers(c_layers, s_layers)
    loss_o, grad_o = ora.sc_grad_tile(img, np.array([0, 0]), layers, c_layers, s_layers, [], lw, cw,
                                      sw, {})
    case = dict(model=model, params=params, ora=ora, img=img, layers=layers, c_layers=c_layers,
                s_layers=s_layers, lw=lw, cw=cw, sw=sw, loss_o=float(loss_o), grad_o=grad_o.copy(),
                content=content, style=style, size=size)
    ora.net.data.clear(), ora.net.diff.clear()          # activations of the last run: not needed
    _tile_cache[name] = case
    return case


@pytest.mark.parametrize('precision,targets', [('fp32', 'oracle'), ('tc32', 'oracle'),
                                               ('fp16', 'oracle'), ('fp16', 'engine'),
                                               ('bf16', 'engine')])
@pytest.mark.parametrize('name', ['cfg1', 'cfg2', 'cfg4'])
def test_single_tile_evaluation_at_config_size(name, precision, targets):
    """eval_sc_grad_tile (style_transfer.py:556-612) of ONE full-size tile of BASELINE configs 1, 2
    and 4.  ``targets``: the style Grams / content features come from the oracle (fp32) or from the
    engine's own preprocessing in the mode under test (what a run does, :488-554).  cfg4's 1024^2
    tile is run in the tensor-core modes only."""
    if name == 'cfg4' and precision in ('fp32', 'bf16'):
        pytest.skip('cfg4 is checked in the tc32 / fp16 modes')
    c = tile_case(name)
    eng = make_engine(c['model'], c['params'], precision)
    if targets == 'oracle':
        hand_targets(eng, c['ora'])
    else:
        eng.contents, eng.styles = [], []
        eng.preprocess_images([c['content']], [c['style']], c['c_layers'], c['s_layers'], c['size'])
        eng.set_contents_and_styles()
    loss_g, grad_g = eng.eval_sc_grad_tile(c['img'], (0, 0), c['layers'], c['c_layers'],
                                           c['s_layers'], [], c['lw'], c['cw'], c['sw'], {})
    check_eval(name + '_tile', precision, c['model'], loss_g, grad_g, c['loss_o'], c['grad_o'], targets)


# ---- tile grids: 724 (ragged ladder size, 2x2 of 362), 1024 (nb = 4), 2048 (nb = 16: cfg3) --------------
_GRID_CASES = {
    'ladder724': (724, (-88, 152)),
    'grid1024': (1024, (200, -312)),
    'cfg3_2048': (2048, (-424, 808)),
}
_grid_cache = {}


def grid_case(name):
    if name in _grid_cache:
        return _grid_cache[name]
    size, roll = _GRID_CASES[name]
    model, tile = 'vgg19.prototxt', 512
    params = he_normal_weights(model_layers(model))
    ora = OracleModel(model, params)
    rs = np.random.RandomState(202)
    content, style, img = (rand_img(rs, size, size) for _ in range(3))
    c_layers, s_layers = ['conv4_2'], STYLE5
    ora.contents, ora.styles = [], []
    # one content pass (the 10-pass averaging is preprocessing, exercised elsewhere)
    ora.img = style.copy()
    feats = ora.prepare_features(s_layers, tile, passes=1)
    ora.styles.append({l: on.gram_lower(feats[l]) for l in feats})
    ora.img = content.copy()
    ora.contents.append(ora.prepare_features(c_layers, tile, passes=1))
    ora.publish()
    lw = {l: 1.0 for l in ora.layers()}
    _, cw = parse_weights(c_layers, 0.05)
    _, sw = parse_weights(s_layers, 1)
    roll = np.array(roll)                           # (x, y) pixels, multiples of 8 (conv4_2's scale)
    ora.img = on.roll2_(img.copy(), roll)
    loss_o, grad_o = ora.sc_grad(roll, c_layers, s_layers, [], lw, cw, sw, {}, tile)
    grad_o = on.roll2_(grad_o.copy(), -roll)
    case = dict(model=model, params=params, ora=ora, img=img, roll=roll, tile=tile,
                c_layers=c_layers, s_layers=s_layers, lw=lw, cw=cw, sw=sw, loss_o=float(loss_o),
                grad_o=grad_o, content=content, style=style)
    ora.net.data.clear(), ora.net.diff.clear()
    _grid_cache[name] = case
    return case


@pytest.mark.parametrize('precision,targets', [('fp32', 'oracle'), ('tc32', 'oracle'),
                                               ('fp16', 'oracle'), ('fp16', 'engine')])
@pytest.mark.parametrize('name', ['ladder724', 'grid1024', 'cfg3_2048'])
def test_tile_grid_evaluation_with_virtual_roll(name, precision, targets):
    """eval_sc_grad (style_transfer.py:614-645) on the un-rolled image + virtual roll against the
    oracle on the physically rolled image, rolled back (:784-806): the batched tensor-core path
    (4 and 16 tiles of 512^2 per launch, the benchmark's shape) and the ragged 2x2 grid of 362^2 the
    reference's scale ladder produces at 724.  The 2048^2 case runs in the benchmark's fp16 mode and
    the tc32 mode."""
    if name == 'cfg3_2048' and precision == 'fp32':
        pytest.skip('the 16-tile grid is checked in the tc32 / fp16 modes')
    c = grid_case(name)
    eng = make_engine(c['model'], c['params'], precision)
    if targets == 'oracle':
        hand_targets(eng, c['ora'])
    else:
        eng.contents, eng.styles = [], []
        eng.preprocess_images([c['content']], [c['style']], c['c_layers'], c['s_layers'], c['tile'],
                              content_passes=1)
        eng.set_contents_and_styles()
    eng.img = eng.to_device(c['img'])
    loss_g, grad_g = eng.eval_sc_grad(c['roll'], c['c_layers'], c['s_layers'], [], c['lw'], c['cw'],
                                      c['sw'], {}, c['tile'])
    check_eval(name, precision, c['model'], loss_g, grad_g, c['loss_o'], c['grad_o'], targets)


# ---- N iterations at cfg1's real size ---------------------------------------------------------------------
_iter_cache = {}


def cfg1_oracle_run(optimizer, iters):
    key = (optimizer, iters)
    if key not in _iter_cache:
        model = 'vgg16.prototxt'
        params = he_normal_weights(model_layers(model))
        rs = np.random.RandomState(303)
        content, style = rand_img(rs, 256, 256), rand_img(rs, 256, 256)
        args = default_args(tile_size=512, optimizer=optimizer, content_layers=['conv4_2'],
                            style_layers=['conv3_1'])
        ot = OracleTransfer(OracleModel(model, params), args)
        np.random.seed(0)
        ot.init_first_scale(256, 256)
        want = ot.run(iters, [content], [style]).copy()
        _iter_cache[key] = (model, params, content, style, args, want)
    return _iter_cache[key]


N_ITER_BOUNDS = {
    # (precision, optimizer): (max |d|, RMS |d|, fraction of pixels with |d| > 1) in grey levels.
    # Measured on B200 (profiles/r02_parity_configs.md): fp32 Adam max 5.8 / RMS 0.09 / 0.14 % of the
    # pixels beyond one grey level, fp32 L-BFGS 2.3 / 0.05 / 0.05 %; fp16 Adam RMS 4.5, L-BFGS RMS 1.4.
    # The fp32-class maxima are single pixels behind a flipped arg-max / ReLU decision whose +-15 grey
    # level first Adam steps (step_size * sign(g)) went opposite ways; hence RMS + fraction bounds.
    ('fp32', 'adam'): (None, 0.25, 1e-2), ('fp32', 'lbfgs'): (None, 0.15, 5e-3),
    ('tc32', 'adam'): (None, 0.25, 1e-2), ('tc32', 'lbfgs'): (None, 0.15, 5e-3),
    ('fp16', 'adam'): (None, 8.0, None), ('fp16', 'lbfgs'): (None, 3.0, None),
}


@pytest.mark.parametrize('precision', ['fp32', 'tc32', 'fp16'])
@pytest.mark.parametrize('optimizer,iters', [('adam', 6), ('lbfgs', 5)])
def test_cfg1_n_iterations_at_real_size(optimizer, iters, precision):
    """BASELINE config 1 as specified: 256x256 single-tile VGG-16, conv4_2 content + conv3_1 style,
    TV + p-norm regularisers, the loop of style_transfer.py:771-828 for N iterations, image after N
    iterations against the oracle's (per-pixel, grey levels of 0..255).  Bounds: N_ITER_BOUNDS."""
    from style_transfer_b200.transfer import StyleTransfer
    model, params, content, style, args, want = cfg1_oracle_run(optimizer, iters)
    eng = make_engine(model, params, precision, mean=MEAN)
    st = StyleTransfer(eng, args)
    np.random.seed(0)
    st.init_first_scale(256, 256)
    got = st.transfer(iters, [content], [style]).cpu().numpy()
    err = np.abs(got - want)
    rms = float(np.sqrt((err.astype(np.float64) ** 2).mean()))
    frac1 = float((err > 1.0).mean())
    report('cfg1_%s_%dit' % (optimizer, iters), precision=precision, max=float(err.max()), rms=rms,
           frac_gt1=frac1, q99=float(np.quantile(err, 0.99)))
    bmax, brms, bfrac = N_ITER_BOUNDS[(precision, optimizer)]
    assert rms <= brms, rms
    if bmax is not None:
        assert err.max() <= bmax, float(err.max())
    if bfrac is not None:
        assert frac1 <= bfrac, frac1


# ---- two scales: scale change of the iterate and of Adam's state ----------------------------------------
@pytest.mark.parametrize('optimizer', ['adam', 'lbfgs'])
def test_two_scale_run_matches_oracle(optimizer):
    """Two rungs of the scale ladder (style_transfer.py:840-881): N iterations at 91x64, the averaged
    iterate resampled to 128x90 (num_utils.resize, Lanczos), the optimizer restarted by
    ``set_params`` (optimizers.py:53-61: i = 1, g1 / p1 Lanczos, g2 bilinear + clamp, beta_accum kept),
    N more iterations; fp32 mode, device-side resize.  Bounds as for one scale (grey levels)."""
    from style_transfer_b200.cli import resize_f32_device
    from style_transfer_b200.transfer import StyleTransfer
    model = 'vgg16.prototxt'
    params = he_normal_weights(model_layers(model))
    rs = np.random.RandomState(404)
    sizes = [(64, 91), (90, 128)]
    contents = [rand_img(rs, *hw) for hw in sizes]
    styles = [rand_img(rs, *hw) for hw in sizes]
    args = default_args(tile_size=64, optimizer=optimizer, content_layers=['conv4_2'],
                        style_layers=['conv2_1', 'conv3_1'])
    iters = (4, 3)
    # oracle
    ot = OracleTransfer(OracleModel(model, params), args)
    np.random.seed(0)
    ot.init_first_scale(*sizes[0])
    raw = ot.run(iters[0], [contents[0]], [styles[0]]).copy()
    ot.model.img = on.resize(raw, sizes[1])
    ot.optimizer.set_params(ot.model.img)
    want = ot.run(iters[1], [contents[1]], [styles[1]]).copy()
    # engine
    eng = make_engine(model, params, 'fp32', mean=MEAN)
    st = StyleTransfer(eng, args)
    np.random.seed(0)
    st.init_first_scale(*sizes[0])
    raw_g = st.transfer(iters[0], [contents[0]], [styles[0]])
    assert np.sqrt(((raw_g.cpu().numpy() - raw) ** 2).mean()) <= 0.25
    eng.img = resize_f32_device(raw_g, sizes[1])
    st.optimizer.set_params(eng.img, resize=resize_f32_device)
    eng.styles = []
    got = st.transfer(iters[1], [contents[1]], [styles[1]]).cpu().numpy()
    err = np.abs(got - want)
    rms = float(np.sqrt((err.astype(np.float64) ** 2).mean()))
    report('two_scale_' + optimizer, precision='fp32', max=float(err.max()), rms=rms)
    # measured on B200: Adam max 0.79 / RMS 0.024; L-BFGS max 5.9 / RMS 0.2 (its fixed-size steps
    # amplify the round-off of the resampled start; the bulk stays within a tenth of a grey level)
    frac1 = float((err > 1.0).mean())
    if optimizer == 'adam':
        assert rms <= 0.25 and frac1 <= 1e-2, (rms, float(err.max()))
    else:
        assert rms <= 0.5 and frac1 <= 2e-2, (rms, frac1, float(err.max()))
